"""HARNESS, not part of the drop-in: mirror of the reference's ECW-CCS ground-state solver (`Solver_GS.Solver_CCS`,
Solver_GS.py:22-243).  SURVEY §2 marks the solvers as callers of the path (kept unchanged by a real integration, which
only swaps `CCS.Gccs` for `ecw_cc_b200.Gccs`); this copy of the loop exists so that BASELINE.json's config 3 and the
CCS ground state can be driven end to end in tests against runs of the unmodified reference solver."""
import numpy as np

from ..diis import DIIS


class Solver_CCS(object):
    """Mirror of the reference's ground-state ECW-CCS solver (`Solver_GS.Solver_CCS`, Solver_GS.py:22-243; the `SCF`
    method — `Gradient` / `L1_grad` drive the reference's dead gradient code and are not provided).  All tensors of
    this loop are o x v or n x n and every step is one call into `ecw_cc_b200.Gccs` (replicas only, latency bound), so
    the amplitudes travel through the numpy API of `Gccs`; DIIS runs on the host store of `ecw_cc_b200.diis`."""

    def __init__(self, mycc, VX_exp, conv='tl', conv_thres=10 ** -6, tsini=None, lsini=None, diis='', maxiter=40,
                 maxdiis=15, CCS_grad=None):
        self.nocc, self.nvir = mycc.nocc, mycc.nvir
        self.tsini = np.zeros((self.nocc, self.nvir)) if tsini is None else tsini
        self.lsini = np.zeros((self.nocc, self.nvir)) if lsini is None else lsini
        self.diis, self.maxdiis = diis, maxdiis
        self.Grad = CCS_grad
        self.mycc, self.myVexp = mycc, VX_exp
        self.maxiter, self.conv_thres = maxiter, conv_thres
        if conv not in ('Ep', 'l', 'tl'):
            raise ValueError('Accepted convergence parameter is Ep, l or tl')
        self.conv = conv
        self.fock = mycc.fock

    def _conv_vector(self, ts, ls, fsp):                   # Solver_GS.py:79-95
        if self.conv == 'Ep':
            return self.mycc.energy_ccs(ts, fsp)
        return ls if self.conv == 'l' else ls + ts

    def SCF(self, L, ts=None, ls=None, diis='', alpha=None, store_ite=False):
        """Solver_GS.py:101-239.  Returns (text, Ep(it), (Delta, vmax)(it), conv(it), last rdm1, (ts, ls)) — or, with
        store_ite, the per-iteration amplitude stacks in place of the last tuple (the reference stores ts twice)."""
        if ts is None:
            ts, ls = self.tsini, self.lsini
        if not diis:
            diis = self.diis
        mycc, VXexp = self.mycc, self.myVexp
        o, v = self.nocc, self.nvir
        rdm1 = mycc.gamma(ts, ls)
        conv, Dconv, ite = 0., 1., 0
        Delta_ite, Ep_ite, conv_ite, ts_ite, ls_ite = [], [], [], [], []
        cl_diis = None
        if diis:
            cl_diis = DIIS()
            cl_diis.space, cl_diis.min_space = self.maxdiis, 2
        while Dconv > self.conv_thres:
            conv_old = conv
            Delta, vmax = VXexp.Vexp_update(rdm1, rdm1, (0, 0), L=L)
            fsp = np.subtract(self.fock, VXexp.Vexp[0, 0])
            Delta_ite.append((Delta, vmax))
            inter = mycc.T1inter(ts, fsp)
            ts = mycc.tsupdate(ts, inter) if alpha is None else mycc.tsupdate_L1(ts, inter, alpha)
            inter = mycc.L1inter(ts, fsp)
            ls = mycc.lsupdate(ts, ls, inter) if alpha is None else mycc.lsupdate_L1(ls, inter, alpha)
            if diis == 'tl':
                vec = cl_diis.update(np.concatenate((np.ravel(ls), np.ravel(ts))))
                ls, ts = (x.reshape((o, v)) for x in np.split(vec, 2))
            rdm1 = mycc.gamma(ts, ls)
            if diis == 'rdm1':
                rdm1 = cl_diis.update(rdm1)
            Ep_ite.append(mycc.energy_ccs(ts, fsp))
            conv = self._conv_vector(ts, ls, fsp)
            if ite > 0:
                Dconv = np.linalg.norm(conv - conv_old)
            conv_ite.append(Dconv)
            if ite >= self.maxiter:
                Conv_text = 'Max iteration reached'
                break
            if Dconv > 10.:
                Conv_text = 'Diverges for lambda = {} after {} iterations'.format(L, ite)
                break
            ite += 1
            if store_ite:
                ts_ite.append(ts)
                ls_ite.append(ts)                          # sic (Solver_GS.py:227)
        else:
            Conv_text = 'Convergence reached for lambda= {}, after {} iteration'.format(L, ite)
        head = (Conv_text, np.asarray(Ep_ite), np.asarray(Delta_ite), np.asarray(conv_ite), rdm1)
        if store_ite:
            return head + (np.asarray(ts_ite), np.asarray(ls_ite))
        return head + ((ts, ls),)
