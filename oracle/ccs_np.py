"""CPU ORACLE (numpy) for the CCS ground- and excited-state path — TEST INFRASTRUCTURE ONLY.

Restates `/root/reference/ECW_CC/CCS.py:23-1518` (class `Gccs` and the module-level
rdm1 builders): T1/Lambda1 intermediates and updates incl. the ES-coupling terms, the
EOM-like right (R1/R0) and left (L1/L0) residuals for H + Vexp, their `Em` extraction,
normalisation helpers and the four rdm1 variants.  Each function cites the lines it follows.
Quirks kept on purpose: the update functions shift the passed `Fab/Fji/Fba/Fij`
IN PLACE (Q6); `force_alpha` zeroes even rows (Q9); `R1inter`'s `'ii,ja,ib->ai'` term is taken
literally; the sign conventions of `vm` differ between `R0inter` and `r0_fromE` (Q9).

Pinned by tests/test_oracle_pins_ccs.py against the live reference (build container) and the
reference-generated vectors in tests/golden/ccs_*.npz.  Only tests/, smoke() and bench.py's
cpu_baseline leg may import this module.
"""
import numpy as np

from .ccsd_np import soft_threshold


def _e(spec, *ops):
    return np.einsum(spec, *ops, optimize=True)


# ----------------------------------------------------------------------------- rdm1 (CCS.py:23-190)
def _assemble(oo, ov, vo, vv, unit_occ):
    no, nv = ov.shape
    dm = np.empty((no + nv, no + nv))
    dm[:no, :no], dm[:no, no:], dm[no:, :no], dm[no:, no:] = oo, ov, vo, vv
    if unit_occ:
        dm[np.arange(no), np.arange(no)] += 1.0
    return dm


def gamma_unsym_CCS(ts, ls):                                   # CCS.py:23-48
    oo = -_e('ie,je->ij', ts, ls)
    vv = _e('ib,ia->ab', ts, ls)
    ov = ts - _e('ja,ib,jb->ia', ts, ts, ls)
    return _assemble(oo, ov, ls.T, vv, True)


def _gamma_es_blocks(ts, ln, rk, r0k, l0n):                    # shared body of CCS.py:75-91 / :130-146
    oo = -r0k * _e('ie,je->ij', ts, ln) - _e('ie,je->ij', rk, ln)
    vo = r0k * ln.T
    vv = r0k * _e('mb,ma->ab', ts, ln) + _e('mb,ma->ab', rk, ln)
    x = _e('ja,jb->ab', ts, ln)
    ov = -r0k * _e('ib,ab->ia', ts, x)
    ov = ov - _e('ma,ie,me->ia', ts, rk, ln) - _e('ie,ma,me->ia', ts, rk, ln)
    ov = ov + ts + l0n * rk
    return oo, ov, vo, vv


def gamma_es_CCS(ts, ln, rk, r0k, l0n):                        # CCS.py:51-102
    if rk is None or isinstance(rk, (float, int)):
        rk, r0k, l0n = np.zeros_like(ts), 1.0, 0.0
    return _assemble(*_gamma_es_blocks(ts, ln, rk, r0k, l0n), unit_occ=True)


def gamma_tr_CCS(ts, ln, rk, r0k, l0n):                        # CCS.py:105-154
    if rk is None or isinstance(rk, (float, int)) or r0k is None:
        rk, r0k = np.zeros_like(ts), 1.0
    return _assemble(*_gamma_es_blocks(ts, ln, rk, r0k, l0n), unit_occ=False)


def gamma_CCS(ts, ls):                                         # CCS.py:157-190
    oo = -_e('ja,ia->ij', ts, ls)
    vv = _e('ia,ib->ab', ts, ls)
    x = _e('ie,me->im', ts, ls)
    vo = ts.T - _e('im,ma->ai', x, ts)
    dm = _assemble(oo + oo.T, ls + vo.T, (ls + vo.T).T, vv + vv.T, False)
    dm *= 0.5
    no = ts.shape[0]
    dm[np.arange(no), np.arange(no)] += 1.0
    return dm


class OracleGccs(object):
    """numpy restatement of `CCS.Gccs` (CCS.py:197-1518)."""

    def __init__(self, eris, fock=None, M_tot=None):
        self.M_tot = 1 if M_tot is None else M_tot             # CCS.py:207-210
        self.fock = np.asarray(eris.fock) if fock is None else fock
        self.eris = eris
        self.nocc = eris.nocc
        self.nvir = self.fock.shape[0] - self.nocc

    def _blocks(self, fsp):
        no = self.nocc
        f = self.fock if fsp is None else fsp
        return f[:no, :no].copy(), f[:no, no:].copy(), f[no:, :no].copy(), f[no:, no:].copy()

    def _eps(self):
        no = self.nocc
        d = np.diagonal(self.fock)
        return d[:no], d[no:]

    # -- energy (CCS.py:226-249) -------------------------------------------------------
    def energy_ccs(self, ts, fsp, rsn=None, r0n=None, vn=None):
        no = ts.shape[0]
        if fsp is None:
            fsp = self.fock.copy()
        e = _e('ia,ia', fsp[:no, no:], ts) + 0.5 * _e('ia,jb,ijab', ts, ts, self.eris.oovv)
        if rsn is not None:
            for rs, v, r0 in zip(rsn, vn, r0n):
                if v is not None:
                    v_ov = -v[:no, no:]
                    e += _e('ia,ia', v_ov, rs) + r0 * _e('ia,ia', v_ov, ts) + r0 * np.trace(-v[:no, :no])
        return e

    gamma = staticmethod(gamma_CCS)
    gamma_unsym = staticmethod(gamma_unsym_CCS)
    gamma_es = staticmethod(gamma_es_CCS)
    gamma_tr = staticmethod(gamma_tr_CCS)

    # -- T1 (CCS.py:271-440) --------------------------------------------------------------
    def T1inter(self, ts, fsp):
        er = self.eris
        foo, fov, fvo, fvv = self._blocks(fsp)
        Fai = fvo + _e('jb,jabi->ai', ts, er.ovvo)
        Fab = fvv - _e('jb,ja->ab', fov, ts) + _e('jc,jacb->ab', ts, er.ovvv)
        Fji = foo + _e('kb,kjbi->ji', ts, er.oovo)
        Fji = Fji - _e('ib,jb->ji', ts, _e('kc,jkcb->jb', ts, er.oovv))
        return Fab, Fji, Fai

    @staticmethod
    def _t1(ts, Fab, Fji, Fai):
        return Fai.T + _e('ib,ab->ia', ts, Fab) - _e('ja,ji->ia', ts, Fji)

    def T1eq(self, ts, fsp):
        return self._t1(ts, *self.T1inter(ts, fsp))

    def tsupdate(self, ts, T1inter, rsn=None, r0n=None, vn=None):
        Fab, Fji, Fai = T1inter
        no, nv = ts.shape
        e_o, e_v = self._eps()
        Fab[np.arange(nv), np.arange(nv)] -= e_v               # in place, CCS.py:307-308
        Fji[np.arange(no), np.arange(no)] -= e_o
        new = self._t1(ts, Fab, Fji, Fai)
        if rsn is not None:                                    # ES coupling, CCS.py:316-347
            if r0n is None:
                raise ValueError('if Vexp are to be calculated, list of r0 amp must be given')
            if len(vn) != len(rsn):
                raise ValueError('Number of experimental potentials must be equal to number of r amplitudes')
            for r, v, r0 in zip(rsn, vn, r0n):
                if v is None:
                    continue
                v_oo, v_vv, v_ov = -v[:no, :no], -v[no:, no:], -v[:no, no:]
                Z = np.trace(v_oo) + _e('jb,jb', v_ov, ts)
                Z0 = v_ov + _e('ib,ab->ia', ts, v_vv) - _e('ja,ji->ia', ts, v_oo)
                Z0 = Z0 - _e('ab,ib->ia', _e('ja,jb->ab', ts, v_ov), ts)
                Zab = v_vv - _e('ja,jb->ab', ts, v_ov)
                Zji = -v_oo - _e('ib,jb->ji', ts, v_ov)
                new = new + r * Z + r0 * Z0 + _e('ab,ib->ia', Zab, r) + _e('ji,ja->ia', Zji, r)
        return new / (e_o[:, None] - e_v)

    def tsupdate_L1(self, ts, T1inter, alpha):                 # CCS.py:353-384
        Fab, Fji, Fai = T1inter
        e_o, e_v = self._eps()
        d = e_o[:, None] - e_v
        w = soft_threshold(self._t1(ts, Fab, Fji, Fai), ts, alpha)
        return (w + ts * d) / d

    # -- Lambda1 (CCS.py:490-698) ----------------------------------------------------------
    def L1inter(self, ts, fsp, E_term=True):
        er = self.eris
        foo, fov, fvo, fvv = self._blocks(fsp)
        Fba = fvv - _e('ja,jb->ba', fov, ts) + _e('jbca,jc->ba', er.ovvv, ts)
        Fba = Fba - _e('ka,kb->ba', _e('jkca,jc->ka', er.oovv, ts), ts)
        Fij = foo + _e('ib,jb->ij', fov, ts) + _e('kibj,kb->ij', er.oovo, ts)
        Fij = Fij + _e('ic,jc->ij', _e('kibc,kb->ic', er.oovv, ts), ts)
        W = np.array(er.voov, copy=True)
        W -= _e('kija,kb->bija', er.ooov, ts)
        W -= _e('icab,jc->bija', _e('kica,kb->icab', er.oovv, ts), ts)
        W += _e('bica,jc->bija', er.vovv, ts)
        Fia = fov + _e('jiba,jb->ia', er.oovv, ts)
        E = (-_e('jb,jb', ts, fov) - 0.5 * _e('jb,kc,jkbc', ts, ts, er.oovv)) if E_term else 0.0
        return Fia, Fba, Fij, W, E

    @staticmethod
    def _l1(ls, Fia, Fba, Fij, W, E):
        return Fia + _e('ib,ba->ia', ls, Fba) - _e('ja,ij->ia', ls, Fij) + _e('jb,bija->ia', ls, W) + ls * E

    def L1eq(self, ts, ls, fsp, E_term=True):
        return self._l1(ls, *self.L1inter(ts, fsp, E_term=E_term))

    def lsupdate(self, ts, ls, L1inter, rsn=None, lsn=None, r0n=None, l0n=None, vn=None):
        Fia, Fba, Fij, W, E = L1inter
        no, nv = ls.shape
        e_o, e_v = self._eps()
        Fba[np.arange(nv), np.arange(nv)] -= e_v               # in place, CCS.py:529-530
        Fij[np.arange(no), np.arange(no)] -= e_o
        new = self._l1(ls, Fia, Fba, Fij, W, E)
        if rsn is not None:                                    # CCS.py:539-579
            if len(lsn) != len(rsn) or len(vn) != len(rsn):
                raise ValueError('v0n, l and r list must be of same length')
            if r0n is None or l0n is None:
                raise ValueError('r0 and l0 values must be given')
            for r, l, v, r0, l0 in zip(rsn, lsn, vn, r0n, l0n):
                if v is None:
                    continue
                v_oo, v_vv, v_ov = -v[:no, :no], -v[no:, no:], -v[:no, no:]
                Pl = _e('jb,jb', r, v_ov) + r0 * _e('jb,jb', ts, v_ov) + r0 * np.trace(v_oo)
                P = np.trace(v_oo) + _e('jb,jb', ts, v_ov)
                Pba = v_vv - _e('jb,ja->ba', ts, v_ov)
                Pij = -v_oo - _e('jb,ib->ij', ts, v_ov)
                new = new + ls * Pl + l0 * v_ov + l * P + _e('ib,ba->ia', l, Pba) + _e('ja,ij->ia', l, Pij)
        return new / (e_o[:, None] - e_v)

    def lsupdate_L1(self, ls, L1inter, alpha):                 # CCS.py:585-617
        e_o, e_v = self._eps()
        d = e_o[:, None] - e_v
        w = soft_threshold(self._l1(ls, *L1inter), ls, alpha)
        return (w + ls * d) / d

    # -- ES right (CCS.py:774-1158) -----------------------------------------------------------
    def R1inter(self, ts, fsp, vm):
        er = self.eris
        no = ts.shape[0]
        foo, fov, fvo, fvv = self._blocks(fsp)
        Fab = fvv - _e('ja,jb->ab', ts, fov) + _e('jc,jacb->ab', ts, er.ovvv)
        Fab = Fab - _e('jc,ka,jkcb->ab', ts, ts, er.oovv)
        Fji = foo + _e('ib,jb->ji', ts, fov) + _e('kb,kjbi->ji', ts, er.oovo)
        Fji = Fji + _e('kb,ic,kjbc->ji', ts, ts, er.oovv)
        W = np.array(er.voov, copy=True)
        W += _e('ib,akbc->akic', ts, er.vovv)
        W -= _e('ib,ja,jkbc->akic', ts, ts, er.oovv)
        W -= _e('ja,jkic->akic', ts, er.ooov)
        Er = _e('jb,jb', ts, fov + 0.5 * _e('kc,jkbc->jb', ts, er.oovv))
        Zab = fvv - _e('ja,jb->ab', ts, fov)
        Zji = foo + _e('kb,kjbi->ji', ts, er.oovo)
        Zji = Zji - _e('kb,ijkb->ji', ts, _e('ic,jkbc->ijkb', ts, er.oovv))
        Zai = fvo + _e('jb,jabi->ai', ts, er.ovvo) + _e('jb,ic,jabc->ai', ts, ts, er.ovvv)
        Tia = Zai.T + _e('ib,ab->ia', ts, Zab) - _e('ja,ji->ia', ts, Zji)
        if vm is None:
            Pia = np.zeros_like(Tia)
        else:
            v_vo, v_vv, v_oo = -vm[no:, :no], -vm[no:, no:], -vm[:no, :no]
            P = v_vo + _e('ab,ib->ai', v_vv, ts) - _e('ii,ja,ib->ai', v_oo, ts, ts)   # literal, CCS.py:869
            Pia = np.ascontiguousarray(P.T)
        return Fab, Fji, W, Er, Tia, Pia

    @staticmethod
    def _r1core(rs, Fab, Fji, W):
        return _e('ab,ib->ia', Fab, rs) - _e('ji,ja->ia', Fji, rs) + _e('akic,kc->ia', W, rs)

    def Extract_Em_r(self, rs, r0, Rinter, ov=None):           # CCS.py:874-906
        Fab, Fji, W, F, Zia, Pia = Rinter
        R = self._r1core(rs, Fab, Fji, W)
        if ov is None:
            o, v = np.unravel_index(np.argmax(abs(rs), axis=None), rs.shape)
        else:
            o, v = ov
        Rov = R[o, v] + rs[o, v] * F + r0 * Zia[o, v] + Pia[o, v]
        return Rov / rs[o, v], o, v

    def rsupdate(self, rs, r0, Rinter, Em, force_alpha=True):  # CCS.py:908-943
        Fab, Fji, W, F, Zia, Pia = Rinter
        no, nv = rs.shape
        e_o, e_v = self._eps()
        Fab[np.arange(nv), np.arange(nv)] -= e_v
        Fji[np.arange(no), np.arange(no)] -= e_o
        new = self._r1core(rs, Fab, Fji, W) + rs * F + r0 * Zia + Pia
        new = new / (Em + e_o[:, None] - e_v)
        if force_alpha:
            new[0::2, :] = 0.0
        return new

    def get_ov(self, ls, l0, rs, r0, ind):                     # CCS.py:945-963
        o, v = ind
        r = rs.copy()
        r[o, v] = 0.0
        return (1.0 - r0 * l0 - _e('ia,ia', r, ls)) / ls[o, v]

    def R1eq(self, rs, r0, Rinter):                            # CCS.py:965-985
        Fab, Fji, W, F, Tia, Pia = Rinter
        return self._r1core(rs, Fab, Fji, W) + rs * F + r0 * Tia + Pia

    def R0inter(self, ts, fsp, vm):                            # CCS.py:987-1034
        no = ts.shape[0]
        if fsp is None:
            fsp = self.fock.copy()
        fov = fsp[:no, no:]
        Fjb = fov + _e('kc,kjcb->jb', ts, self.eris.oovv)
        E = _e('jb,jb', ts, fov + 0.5 * _e('kc,jkbc->jb', ts, self.eris.oovv))
        P = np.trace(vm[:no, :no]) + _e('jb,jb', ts, vm[:no, no:])
        return Fjb, E, P

    def Extract_r0(self, r1, ts, fsp, vm):                     # CCS.py:1036-1079
        """r0 from the R1 and R0 equations for a given r1.  Kept as the reference has it: the roots are divided by c
        (not 2a), `return 0` when c == 0., ValueError when both roots are negative."""
        f = self.fock if fsp is None else fsp.copy()
        Fab, Fji, W, F, Zia, Pia = self.R1inter(ts, f, vm)
        Fjb, Z, P = self.R0inter(ts, f, vm)
        R1 = self._r1core(r1, Fab, Fji, W) + r1 * F + Pia
        c = -_e('jb,jb', r1, Fjb) - P
        if c == 0.:
            return 0
        i, j = np.unravel_index(np.argmax(abs(r1), axis=None), r1.shape)
        a = Zia[i, j] / r1[i, j]
        b = R1[i, j] / r1[i, j] - Z
        r0_1 = (-b + np.sqrt((b ** 2) - (4 * a * c))) / c
        r0_2 = (-b - np.sqrt((b ** 2) - (4 * a * c))) / c
        if r0_1 > 0:
            return r0_1
        elif r0_2 > 0:
            return r0_2
        raise ValueError('Both solution for r0 are negative')

    def r0update(self, rs, r0, Em, R0inter):                   # CCS.py:1081-1096
        Fjb, E, P = R0inter
        return (_e('jb,jb', rs, Fjb) + P + r0 * E) / Em

    def R0eq(self, rs, r0, R0inter):                           # CCS.py:1098-1114
        Fjb, E, P = R0inter
        return _e('jb,jb', rs, Fjb) + r0 * E + P

    def r0_fromE(self, En, t1, r1, vm0, fsp=None):             # CCS.py:1116-1158
        if fsp is None:
            fsp = self.fock.copy()
        no, nv = r1.shape
        if vm0 is not None:
            vov, voo = -vm0[:no, no:], -vm0[:no, :no]
        else:
            vov, voo = np.zeros((no, nv)), np.zeros((no, no))
        fov = fsp[:no, no:]
        d = En - _e('jb,jb', t1, fov) - 0.5 * _e('jb,kc,jkbc', t1, t1, self.eris.oovv)
        r0 = _e('jb,jb', r1, fov) + _e('kc,jb,jkbc', r1, t1, self.eris.oovv)
        r0 += _e('jb,jb', t1, vov) + np.trace(voo)
        return r0 / d

    # -- ES left (CCS.py:1164-1518) ---------------------------------------------------------------
    def es_L1inter(self, ts, fsp, vm):
        er = self.eris
        no, nv = ts.shape
        foo, fov, fvo, fvv = self._blocks(fsp)
        Fba = fvv - _e('jb,ja->ba', ts, fov) + _e('jc,jbca->ba', ts, er.ovvv)
        Fba = Fba - _e('jc,kb,jkca->ba', ts, ts, er.oovv)
        Fij = foo + _e('jb,ib->ij', ts, fov) + _e('kb,kibj->ij', ts, er.oovo)
        Fij = Fij + _e('kb,jc,kibc->ij', ts, ts, er.oovv)
        W = np.array(er.voov, copy=True)
        W -= _e('kb,kija->bija', ts, er.ooov)
        W += _e('jc,bica->bija', ts, er.vovv)
        W -= _e('jc,kb,kica->bija', ts, ts, er.oovv)
        El = _e('jb,jb', ts, fov + 0.5 * _e('kc,jkbc->jb', ts, er.oovv))
        Zia = fov + _e('jb,jiba->ia', ts, er.oovv)
        P = np.zeros((no, nv)) if vm is None else -vm[:no, no:].copy()
        return Fba, Fij, W, El, Zia, P

    def L0inter(self, ts, fsp, vm):                            # CCS.py:1236-1286
        er = self.eris
        no = ts.shape[0]
        if fsp is None:
            fsp = self.fock.copy()
        foo, fov, fvo, fvv = fsp[:no, :no], fsp[:no, no:], fsp[no:, :no], fsp[no:, no:]
        Fbj = fvo - _e('kb,kj->bj', ts, foo) + _e('ja,ba->bj', ts, fvv) - _e('jc,kb,kc->bj', ts, ts, fov)
        x = np.array(er.ovvo, copy=True)
        x += _e('lb,jd,lkcd->kbcj', ts, ts, er.oovv)
        x -= _e('lb,klcj->kbcj', ts, er.oovo)
        x += _e('jd,kbcd->kbcj', ts, er.ovvv)
        Wjb = _e('kc,kbcj->jb', ts, x)
        Z = _e('jb,jb', ts, fov + 0.5 * _e('kc,jkbc->jb', ts, er.oovv))
        P = _e('ia,ia', ts, vm[:no, no:]) + np.sum(np.diagonal(vm[:no, :no]))
        return Fbj, Wjb, Z, P

    @staticmethod
    def _l1core(ls, Fba, Fij, W):
        return _e('ib,ba->ia', ls, Fba) - _e('ja,ij->ia', ls, Fij) + _e('jb,bija->ia', ls, W)

    def Extract_Em_l(self, ls, l0, L1inter, ov=None):          # CCS.py:1288-1319
        Fba, Fij, W, F, Zia, P = L1inter
        if ov is None:
            o, v = np.unravel_index(np.argmax(abs(ls), axis=None), ls.shape)
        else:
            o, v = ov
        L = self._l1core(ls, Fba, Fij, W)
        Lov = L[o, v] + ls[o, v] * F + l0 * Zia[o, v] + P[o, v]
        return Lov / ls[o, v], o, v

    def es_lsupdate(self, ls, l0, Em, L1inter, force_alpha=True):   # CCS.py:1366-1399
        Fba, Fij, W, F, Zia, P = L1inter
        no, nv = ls.shape
        e_o, e_v = self._eps()
        Fba[np.arange(nv), np.arange(nv)] -= e_v
        Fij[np.arange(no), np.arange(no)] -= e_o
        new = self._l1core(ls, Fba, Fij, W) + ls * F + l0 * Zia + P
        new = new / (Em + e_o[:, None] - e_v)
        if force_alpha:
            new[0::2, :] = 0.0
        return new

    def es_L1eq(self, ls, l0, es_L1inter):                     # CCS.py:1401-1421
        Fba, Fij, W, El, Zia, P = es_L1inter
        return self._l1core(ls, Fba, Fij, W) + ls * El + l0 * Zia + P

    def l0update(self, ls, l0, Em, L0inter):                   # CCS.py:1423-1439
        Fbj, Wjb, Z, P = L0inter
        return (_e('jb,bj', ls, Fbj) + _e('jb,jb', ls, Wjb) + P + l0 * Z) / Em

    def L0eq(self, ls, l0, L0inter):                           # CCS.py:1441-1457
        Fbj, Wjb, El, P = L0inter
        return _e('jb,bj', ls, Fbj) + _e('jb,jb', ls, Wjb) + l0 * El + P

    def l0_fromE(self, En, t1, l1, v0m, fsp=None):             # CCS.py:1459-1518
        er = self.eris
        no, nv = t1.shape
        if fsp is None:
            fsp = self.fock.copy()
        fov, fvv, foo = fsp[:no, no:], fsp[no:, no:], fsp[:no, :no]
        if v0m is not None:
            vov, voo = v0m[:no, no:], v0m[:no, :no]
        else:
            vov, voo = np.zeros((no, nv)), np.zeros((no, no))
        d = En                                                 # Q12: in place when En is an array (CCS.py:1488-1490)
        d -= 0.5 * _e('jb,kc,jkbc', t1, t1, er.oovv)
        l0 = _e('jb,jb', l1, fov) + _e('jb,ab,ja', t1, fvv, l1) - _e('jb,kb,kj', l1, t1, foo)
        l0 -= _e('jc,kb,kc,jb', t1, t1, fov, l1)
        l0 += _e('jb,kc,kbcj', l1, t1, er.ovvo)
        x = _e('jb,jd->bd', l1, t1)
        l0 += _e('bd,kb,lc,klcd', x, t1, t1, er.oovv)
        l0 -= _e('jl,kc,klcj', _e('jb,lb->jl', l1, t1), t1, er.oovo)
        l0 += _e('bd,kc,kbcd', x, t1, er.ovvv)
        l0 += _e('ia,ia', t1, vov) + np.trace(voo)
        return l0 / d
