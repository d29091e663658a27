"""Pulay DIIS with PySCF's lib.diis.DIIS call surface (`space`, `min_space`,
`update(x)`); error vector = x_k - x_{k-1}.  Test infrastructure only; results
of DIIS-accelerated loops are "parity unpinned" (no reference test pins them).
"""
import numpy as np


class DIIS:
    def __init__(self, dev=None, filename=None, incore=True):
        self.space = 6
        self.min_space = 1
        self._xs = []
        self._es = []
        self._xprev = None

    def update(self, x, xerr=None):
        x = np.asarray(x, dtype=float).ravel().copy()
        if xerr is None:
            if self._xprev is None:
                self._xprev = x
                if self.min_space > 0:
                    return x
                err = x.copy()
            else:
                err = x - self._xprev
        else:
            err = np.asarray(xerr).ravel().copy()
        self._xs.append(x)
        self._es.append(err)
        if len(self._xs) > self.space:
            self._xs.pop(0)
            self._es.pop(0)
        nd = len(self._xs)
        if nd < self.min_space:
            self._xprev = x
            return x
        H = np.zeros((nd + 1, nd + 1))
        H[0, 1:] = H[1:, 0] = 1.0
        for i in range(nd):
            for j in range(i + 1):
                H[i + 1, j + 1] = H[j + 1, i + 1] = np.dot(self._es[i], self._es[j])
        g = np.zeros(nd + 1)
        g[0] = 1.0
        try:
            c = np.linalg.solve(H, g)
        except np.linalg.LinAlgError:
            c = np.linalg.lstsq(H, g, rcond=None)[0]
        xnew = np.zeros_like(x)
        for i in range(nd):
            xnew += c[i + 1] * self._xs[i]
        self._xprev = xnew
        return xnew
