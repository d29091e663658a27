"""Pulay DIIS with the call surface of PySCF's `lib.diis.DIIS` (`space`, `min_space`, `update(x)`), restated from the
published algorithm (PySCF 2.x, `pyscf/lib/diis.py`; the package is not installed here and its source is not part of
the reference tree, so results of DIIS-accelerated loops are "parity unpinned" — SURVEY §8c).  TEST INFRASTRUCTURE ONLY.

What the published algorithm does, and this restates:
  * the first vector handed to `update` only becomes `xprev` (no error vector yet);
  * afterwards each call stores x_k and e_k = x_k - xprev in a ring of `space` slots, where `xprev` is the vector the
    previous call RETURNED after an extrapolation — while fewer than `min_space` vectors are stored, `update` returns
    x unchanged and `xprev` keeps its old value;
  * B_ij = <e_i, e_j> bordered by ones; solve B c = (1,0,..,0) — through the eigen-decomposition with eigenvalues
    |w| < 1e-14 dropped when there are any, else a direct solve;  x_new = sum_i c_i x_i, which also becomes `xprev`.
"""
import numpy as np


class DIIS(object):
    def __init__(self, dev=None, filename=None, incore=True):
        self.space = 6
        self.min_space = 1
        self._head = 0
        self._book = []
        self._x = {}
        self._e = {}
        self._H = None
        self._xprev = None

    def get_num_vec(self):
        return len(self._book)

    def update(self, x, xerr=None):
        if xerr is not None:
            raise NotImplementedError("pyscf stub: DIIS.update with an explicit error vector")
        shape = np.shape(x)
        x = np.array(x, dtype=np.float64).ravel()
        while len(self._book) >= self.space:
            self._book.pop(0)
        if self._xprev is None:
            self._xprev = x
        else:
            if self._head >= self.space:
                self._head = 0
            self._book.append(self._head)
            self._x[self._head] = x
            self._e[self._head] = x - self._xprev
            self._head += 1
        nd = self.get_num_vec()
        if nd < self.min_space:
            return x.reshape(shape)
        if self._H is None:
            self._H = np.zeros((self.space + 1, self.space + 1))
            self._H[0, 1:] = self._H[1:, 0] = 1.0
        dt = self._e[self._head - 1]
        for i in range(nd):
            tmp = np.dot(dt, self._e[i])
            self._H[self._head, i + 1] = tmp
            self._H[i + 1, self._head] = tmp
        c = solve_coefficients(self._H[:nd + 1, :nd + 1])
        xnew = np.zeros_like(x)
        for i, ci in enumerate(c[1:]):
            xnew += self._x[i] * ci
        self._xprev = xnew
        return xnew.reshape(shape)


def solve_coefficients(h):
    g = np.zeros(h.shape[0])
    g[0] = 1.0
    w, v = np.linalg.eigh(h)
    if np.any(abs(w) < 1e-14):
        idx = abs(w) > 1e-14
        return np.dot(v[:, idx] * (1.0 / w[idx]), np.dot(v[:, idx].T, g))
    return np.linalg.solve(h, g)
