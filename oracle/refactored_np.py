"""numpy statement of the REFACTORED CCSD algorithm the CUDA plan implements —
TEST INFRASTRUCTURE ONLY (never imported by the product).

Same mathematics as `oracle/ccsd_np.py` (i.e. the reference `CCSD.py:248-623`)
but in the form the device executes (DESIGN.md "Algorithm"):
  * no Wvvvv (CCSD.py:396-402) and no wvvvo (CCSD.py:602-605) are ever formed;
  * both particle-particle ladders run on antisymmetry-packed operands
    (pair index p(a<b) = b(b-1)/2 + a);
  * ring terms run in the particle-hole layout X[(ia),(jb)];
  * only the canonical integral blocks oooo, ooov, oovv, ovov, ovvv, vvvv are
    used (Eris.py:128 symmetries give the rest).
`tests/test_refactored_spec.py` proves it equal to the reference factorisation
at small sizes; `ecw_cc_b200/csrc/ccsd_plan.cpp` is a transliteration of the
contraction list below (same index strings, same order).
"""
import numpy as np

from .ccsd_np import soft_threshold


def es(spec, *ops):
    return np.einsum(spec, *ops, optimize=True)


def pair_index(n):
    """(lo, hi) index arrays of the packed antisymmetric pair list, p = hi(hi-1)/2 + lo."""
    hi, lo = np.tril_indices(n, -1)   # hi > lo, ordered by hi then lo
    return lo, hi


def pack_last(x):
    """x[..., a, b] (antisymmetric) -> x[..., p(a<b)]."""
    lo, hi = pair_index(x.shape[-1])
    return np.ascontiguousarray(x[..., lo, hi])


def pack_first(x):
    lo, hi = pair_index(x.shape[0])
    return np.ascontiguousarray(x[lo, hi])


def unpack_last(xp, n):
    lo, hi = pair_index(n)
    out = np.zeros(xp.shape[:-1] + (n, n))
    out[..., lo, hi] = xp
    out[..., hi, lo] = -xp
    return out


def unpack_first(xp, n):
    lo, hi = pair_index(n)
    out = np.zeros((n, n) + xp.shape[1:])
    out[lo, hi] = xp
    out[hi, lo] = -xp
    return out


class DeviceErisSpec(object):
    """The constant layouts the device keeps (built once per integral set)."""

    def __init__(self, eris):
        o = eris.nocc
        self.o = o
        self.oooo = np.asarray(eris.oooo)
        self.ooov = np.asarray(eris.ooov)
        self.oovv = np.asarray(eris.oovv)
        self.ovvv = np.asarray(eris.ovvv)
        ovov = np.asarray(eris.ovov)
        v = self.oovv.shape[2]
        self.v = v
        self.oovv_ph = np.ascontiguousarray(self.oovv.transpose(0, 2, 1, 3))     # [(me),(nf)] = oovv[mnef]
        self.ovov_ph = np.ascontiguousarray(ovov.transpose(2, 1, 0, 3))          # [(ia),(nf)] = ovov[naif]
        self.oooo_p = pack_last(pack_first(self.oooo))                           # [ij_p, kl_p]
        self.oovv_p = pack_last(pack_first(self.oovv))                           # [ij_p, ab_p]
        self.ovvv_p = pack_last(self.ovvv)                                       # [m, a, ef_p]
        vvvv = np.asarray(eris.vvvv)
        self.vvvv_p = pack_last(pack_first(vvvv))                                # [ab_p, cd_p]


def make_tau(t1, t2, c=1.0):
    x = es('ia,jb->ijab', t1, t1)
    return t2 + c * (x - x.transpose(0, 1, 3, 2))


def finish(r1, r2, a1, a2, e_o, e_v, alpha, equation):
    d1 = e_o[:, None] - e_v[None, :]
    d2 = d1[:, None, :, None] + d1[None, :, None, :]
    if alpha is not None:
        w2 = soft_threshold(r2, a2, alpha)
        if equation:
            return r1, w2
        return (r1 + a1 * d1) / d1, (w2 + a2 * d2) / d2
    if not equation:
        return r1 / d1, r2 / d2
    return r1, r2


def antisym_ph(x_ph):
    """x_ph[i,a,j,b] -> P(ij)P(ab) x as [i,j,a,b]."""
    x = x_ph.transpose(0, 2, 1, 3)
    x = x - x.transpose(1, 0, 2, 3)
    return x - x.transpose(0, 1, 3, 2)


def tupdate(E, fock, t1, t2, fsp=None, alpha=None, equation=False):
    o, v = t1.shape
    if fsp is None:
        fsp = fock
    foo, fov, fvv = fsp[:o, :o], fsp[:o, o:], fsp[o:, o:]
    e_o, e_v = np.diagonal(fock)[:o], np.diagonal(fock)[o:]
    shift = (not equation) and alpha is None

    tau = make_tau(t1, t2, 1.0)
    ttl = make_tau(t1, t2, 0.5)
    tau_p = pack_last(pack_first(tau))                       # [ij_p, ef_p]
    t2ph = np.ascontiguousarray(t2.transpose(0, 2, 1, 3))   # [(ia),(jb)]

    # one-body intermediates (CCSD.py:355-387)
    G = es('menf,nf->me', E.oovv_ph, t1)
    Fov = fov + G
    Fvv = fvv - 0.5 * es('me,ma->ae', fov, t1)
    Fvv = Fvv - es('maef,mf->ae', E.ovvv, t1)                # vovv[amef] = -ovvv[maef]
    Fvv = Fvv - 0.5 * es('mnfa,mnfe->ae', ttl, E.oovv)
    Foo = foo + 0.5 * es('me,ie->mi', fov, t1)
    Foo = Foo + es('mnie,ne->mi', E.ooov, t1)
    Foo = Foo + 0.5 * es('mnef,inef->mi', E.oovv, ttl)
    if shift:
        Fvv = Fvv - np.diag(e_v)
        Foo = Foo - np.diag(e_o)

    # T1 (CCSD.py:288-294)
    r1 = fov + es('ie,ae->ia', t1, Fvv) - es('mi,ma->ia', Foo, t1)
    r1 += es('iame,me->ia', t2ph, Fov)
    r1 -= es('ianf,nf->ia', E.ovov_ph, t1)
    r1 -= 0.5 * es('imef,maef->ia', t2, E.ovvv)
    r1 += 0.5 * es('mnea,mnie->ia', t2, E.ooov)

    # T2 (CCSD.py:297-314)
    F1 = Fvv - 0.5 * es('mb,me->be', t1, Fov)
    x = es('ijae,be->ijab', t2, F1)
    r2 = E.oovv + x - x.transpose(0, 1, 3, 2)
    F2 = Foo + 0.5 * es('je,me->mj', t1, Fov)
    x = es('mj,imab->ijab', F2, t2)
    r2 -= x - x.transpose(1, 0, 2, 3)

    # packed accumulator: hole-hole ladder + particle-particle ladder + Wvvvv's t1 part
    x = es('mnie,je->mnij', E.ooov, t1)
    Woo_p = E.oooo_p + pack_last(pack_first(x - x.transpose(0, 1, 3, 2)))
    Woo_p = Woo_p + es('mf,if->mi', E.oovv_p, tau_p)          # 1/2 tau.oovv (K3 folded in)
    acc_p = es('mi,ma->ia', Woo_p, tau_p)                     # 1/2 tau[mnab] W[mnij]
    acc_p += es('if,af->ia', tau_p, E.vvvv_p)                 # 1/2 tau[ijef] vvvv[abef]
    ovvv_p2 = E.ovvv_p.reshape(o * v, -1)
    Y_p = -2.0 * es('if,qf->iq', tau_p, ovvv_p2).reshape(-1, o, v)   # Y[ij_p, m, a]
    Z = es('pma,mb->pab', Y_p, t1)                            # [ij_p, a, b]
    acc_p += -0.5 * pack_last(Z - Z.transpose(0, 2, 1))
    r2 += unpack_first(unpack_last(acc_p, v), o)

    # ring (CCSD.py:306-310, 404-413) in ph layout W'[(me),(jb)] = Wovvo[m,b,e,j]
    Wph = 0.5 * es('menf,nfjb->mejb', E.oovv_ph, t2ph)
    Wph += es('mbef,jf->mejb', E.ovvv, t1)
    Wph += es('nb,mnje->mejb', t1, E.ooov)
    U = es('mnef,jf->mnej', E.oovv, t1)
    Wph -= es('nb,mnej->mejb', t1, U)
    Wph -= E.ovov_ph
    ring = es('iame,mejb->iajb', t2ph, Wph)
    Q = es('jbme,ie->jbmi', E.ovov_ph, t1)                    # ovov[mbje] = ovov_ph[jbme]
    ring += es('ma,jbmi->iajb', t1, Q)
    r2 += antisym_ph(ring)

    x = -es('ie,jeab->ijab', t1, E.ovvv)                      # ovvv[jeba] = -ovvv[jeab]
    r2 += x - x.transpose(1, 0, 2, 3)
    x = es('ma,ijmb->ijab', t1, E.ooov)
    r2 -= x - x.transpose(0, 1, 3, 2)
    return finish(r1, r2, t1, t2, e_o, e_v, alpha, equation)


def energy(E, t1, t2, fsp):
    o = t1.shape[0]
    G = es('menf,nf->me', E.oovv_ph, t1)
    return float(np.sum(fsp[:o, o:] * t1) + 0.25 * np.sum(t2 * E.oovv) + 0.5 * np.sum(t1 * G))


def lupdate(E, fock, t1, t2, l1, l2, fsp=None, alpha=None, equation=False):
    o, v = t1.shape
    if fsp is None:
        fsp = fock
    foo, fov, fvo, fvv = fsp[:o, :o], fsp[:o, o:], fsp[o:, :o], fsp[o:, o:]
    e_o, e_v = np.diagonal(fock)[:o], np.diagonal(fock)[o:]
    shift = (equation is False) and alpha is None

    tau = make_tau(t1, t2, 1.0)
    tau_p = pack_last(pack_first(tau))
    l2_p = pack_last(pack_first(l2))
    t2ph = np.ascontiguousarray(t2.transpose(0, 2, 1, 3))
    l2ph = np.ascontiguousarray(l2.transpose(0, 2, 1, 3))

    # ---- intermediates (CCSD.py:543-623) ----
    G = es('menf,nf->me', E.oovv_ph, t1)
    Fov = fov + G                                             # = fov1 (:474) = tmp (:504) = x (:580)
    v1 = fvv - es('ja,jb->ba', fov, t1)
    v1 = v1 - es('jbac,jc->ba', E.ovvv, t1)
    v1 = v1 - 0.5 * es('jkcb,jkca->ba', tau, E.oovv)
    v2 = foo + es('ib,jb->ij', fov, t1)
    v2 = v2 - es('kijb,kb->ij', E.ooov, t1)
    v2 = v2 + 0.5 * es('ikbc,jkbc->ij', E.oovv, tau)

    v4ph = es('kcld,ldjb->kcjb', t2ph, E.oovv_ph) - E.ovov_ph   # [(kc),(jb)] = v4[j,c,b,k]

    v5T = fvo.T + es('jbkc,kc->jb', t2ph, fov)               # v5T[j,b] = v5[b,j]
    q = es('kc,jc->kj', Fov, t1)
    v5T = v5T + es('kj,kb->jb', q, t1)
    v5T = v5T + 0.5 * es('kljc,klcb->jb', E.ooov, t2)
    v5T = v5T - 0.5 * es('jkdc,kbdc->jb', t2, E.ovvv)

    w3T = v5T + es('kcjb,jb->kc', v4ph, t1)                  # w3T[k,c] = w3[c,k]
    w3T = w3T + es('kb,cb->kc', t1, v1)
    w3T = w3T - es('jk,jc->kc', v2, t1)

    # hole-hole pieces, packed [ij_p, kl_p]
    v3_p = 2.0 * es('if,kf->ik', E.oovv_p, tau_p)
    y = es('jilc,kc->jilk', E.ooov, t1)
    woo_p = 0.5 * E.oooo_p + 0.25 * v3_p + 0.5 * pack_last(pack_first(y - y.transpose(0, 1, 3, 2)))
    lt_p = 2.0 * es('if,kf->ik', l2_p, tau_p)                # (l2.tau)[ij_p, kl_p]

    S = es('ljbd,kd->ljbk', E.oovv, t1)                      # = -sum_d oovv[ljdb] t1[kd]
    wph = v4ph + es('lc,ljbk->kcjb', t1, S)
    wph = wph - es('lc,ljkb->kcjb', t1, E.ooov)
    wph = wph + es('jcbd,kd->kcjb', E.ovvv, t1)

    ovvv_p2 = E.ovvv_p.reshape(o * v, -1)
    wo_p = 0.5 * es('qf,kf->qk', ovvv_p2, tau_p)             # 1/4 ovvv.tau, [(ic), jk_p]
    wovoo = unpack_last(wo_p, o).reshape(o, v, o, o)
    wovoo = wovoo + 0.5 * E.ooov.transpose(2, 3, 0, 1)
    wovoo = wovoo + es('kcib,jb->icjk', v4ph, t1)
    wovoo = wovoo - es('kclb,lijb->icjk', t2ph, E.ooov)

    # ---- m3 (CCSD.py:461-470), packed [ij_p, ab_p] ----
    m3_p = 2.0 * es('ik,ka->ia', woo_p, l2_p)
    m3_p += 0.5 * es('ik,ka->ia', lt_p, E.oovv_p)
    l2t1 = es('ijcd,kd->ijck', l2, t1)
    a_p = pack_first(np.ascontiguousarray(l2t1.transpose(0, 1, 3, 2))).reshape(-1, o * v)   # [ij_p,(kc)]
    m3_p += es('pq,qa->pa', a_p, ovvv_p2)                    # -(kcba) = +(kcab)
    m3_p += es('if,af->ia', l2_p, E.vvvv_p)                  # 1/2 l2.vvvv
    m3 = unpack_first(unpack_last(m3_p, v), o)

    m_vv = 0.5 * es('klcb,klca->ba', t2, l2)
    m_oo = 0.5 * es('kicd,kjcd->ij', l2, t2)
    x_vv = m_vv + es('ka,kb->ba', l1, t1)
    x_oo = m_oo + es('ic,kc->ik', l1, t1)
    if shift:
        v1s = v1 - np.diag(e_v)
        v2s = v2 - np.diag(e_o)
    else:
        v1s, v2s = v1, v2

    # ---- L2 (CCSD.py:472-488) ----
    r2 = E.oovv + m3
    ring = es('iakc,kcjb->iajb', l2ph, wph) + es('ia,jb->iajb', l1, Fov)
    r2 += antisym_ph(ring)
    y = es('ka,ijkb->ijab', l1, E.ooov) - es('ijac,cb->ijab', l2, v1s)
    y += es('ca,ijcb->ijab', x_vv, E.oovv)
    r2 -= y - y.transpose(0, 1, 3, 2)
    y = es('qc,pcrs->pqrs', l1, E.ovvv)                       # 'ic,jcba->jiba' (positional)
    y = y + es('qk,kprs->pqrs', v2s, l2) - es('pk,kqrs->pqrs', x_oo, E.oovv)
    r2 += y - y.transpose(1, 0, 2, 3)

    # ---- L1 (CCSD.py:490-506) ----
    r1 = fov - es('jbia,jb->ia', E.ovov_ph, l1)               # ovvo[ibaj] = -ovov_ph[(jb),(ia)]
    r1 += es('ib,ba->ia', l1, v1s) - es('ij,ja->ia', v2s, l1)
    r1 -= es('icjk,kjca->ia', wovoo, l2)
    # -(l2 . wvvvo), wvvvo never formed:
    r1 += es('ikcj,kcja->ia', l2t1, v4ph)
    lt = unpack_first(unpack_last(lt_p, o), o)                # [i,k,j,l]
    r1 -= 0.25 * es('ikjl,jlka->ia', lt, E.ooov)
    r1 -= 0.5 * es('ikbc,kabc->ia', l2, E.ovvv)
    Xph = es('ibjc,jckd->ibkd', l2ph, t2ph)
    r1 += es('ibkd,kbda->ia', Xph, E.ovvv)
    r1 += es('ijab,jb->ia', m3, t1)
    r1 += es('iajb,jb->ia', l2ph, w3T)
    z = t1 + es('jbkc,kc->jb', t2ph, l1) - es('bd,jd->jb', x_vv, t1) - es('lj,lb->jb', m_oo, t1)
    r1 += es('iajb,jb->ia', E.oovv_ph, z)
    r1 -= es('icba,bc->ia', E.ovvv, x_vv)
    r1 -= es('jika,kj->ia', E.ooov, x_oo)
    r1 -= es('ik,ka->ia', m_oo, Fov)
    r1 -= es('ca,ic->ia', m_vv, Fov)

    if shift:
        Ecc = energy(E, t1, t2, fsp)
        r1 = r1 * (1.0 - Ecc)
        r2 = r2 * (1.0 - Ecc)
    return finish(r1, r2, l1, l2, e_o, e_v, alpha, equation)


def gamma(t1, t2, l1, l2):
    o, v = t1.shape
    D = es('imef,jmef->ij', l2, t2)
    doo = -es('ie,je->ij', l1, t1) - 0.5 * D
    dvv = es('ma,mb->ab', t1, l1) + 0.5 * es('mnea,mneb->ab', t2, l2)
    t2ph = t2.transpose(0, 2, 1, 3)
    dvoT = es('iame,me->ia', t2ph, l1) - 0.5 * es('mi,ma->ia', D, t1) - es('ie,ae->ia', t1, dvv) + t1
    dm = np.empty((o + v, o + v))
    dm[:o, :o] = 0.5 * (doo + doo.T)
    dm[:o, o:] = 0.5 * (l1 + dvoT)
    dm[o:, :o] = dm[:o, o:].T
    dm[o:, o:] = 0.5 * (dvv + dvv.T)
    dm[np.arange(o), np.arange(o)] += 1.0
    return dm


# ======================================================================
# GENERAL path: amplitudes NOT assumed antisymmetric.
#
# The reference's L1 update (utilities.py:53-67: v<=0 is soft-thresholded,
# v>0 gets e+alpha) destroys the antisymmetry of t2/l2 after the first
# regularised iteration, and its dense einsums are evaluated on whatever they
# are given.  The variant below uses only the antisymmetry of the INTEGRALS
# (Eris.py:128), never that of the amplitudes: contracted integral pairs are
# still packed (the amplitude is antisymmetrised while packing), but the (i,j)
# rows of the ladders stay dense (o^2 instead of o(o-1)/2).
# ======================================================================
def make_tau2(t1, t2, c1, c2):
    """t2 + c1 t1[ia]t1[jb] - c2 t1[ib]t1[ja]"""
    x = es('ia,jb->ijab', t1, t1)
    return t2 + c1 * x - c2 * x.transpose(0, 1, 3, 2)


def tupdate_general(E, fock, t1, t2, fsp=None, alpha=None, equation=False):
    o, v = t1.shape
    if fsp is None:
        fsp = fock
    foo, fov, fvv = fsp[:o, :o], fsp[:o, o:], fsp[o:, o:]
    e_o, e_v = np.diagonal(fock)[:o], np.diagonal(fock)[o:]
    shift = (not equation) and alpha is None

    tau = make_tau2(t1, t2, 1.0, 1.0)
    ttl = make_tau2(t1, t2, 0.5, 0.5)
    tau_q = 0.5 * pack_last(tau - tau.transpose(0, 1, 3, 2)).reshape(o * o, -1)   # [(ij), ef_p]
    tau_r = 0.5 * pack_first(tau - tau.transpose(1, 0, 2, 3)).reshape(-1, v * v)  # [mn_p, (ab)]
    t2ph = np.ascontiguousarray(t2.transpose(0, 2, 1, 3))          # [(ia),(jb)] = t2[i,j,a,b]
    t2ph2 = np.ascontiguousarray(t2.transpose(1, 2, 0, 3))         # [(nf),(jb)] = t2[j,n,f,b]

    G = es('menf,nf->me', E.oovv_ph, t1)
    Fov = fov + G
    Fvv = fvv - 0.5 * es('me,ma->ae', fov, t1)
    Fvv = Fvv - es('maef,mf->ae', E.ovvv, t1)
    Fvv = Fvv + 0.5 * es('mnaf,mnfe->ae', ttl, E.oovv)
    Foo = foo + 0.5 * es('me,ie->mi', fov, t1)
    Foo = Foo + es('mnie,ne->mi', E.ooov, t1)
    Foo = Foo + 0.5 * es('mnef,inef->mi', E.oovv, ttl)
    if shift:
        Fvv = Fvv - np.diag(e_v)
        Foo = Foo - np.diag(e_o)

    r1 = fov + es('ie,ae->ia', t1, Fvv) - es('mi,ma->ia', Foo, t1)
    r1 += es('iame,me->ia', t2ph, Fov)
    r1 -= es('ianf,nf->ia', E.ovov_ph, t1)
    r1 -= 0.5 * es('imef,maef->ia', t2, E.ovvv)
    r1 -= 0.5 * es('mnae,mnie->ia', t2, E.ooov)

    F1 = Fvv - 0.5 * es('mb,me->be', t1, Fov)
    x = es('ijae,be->ijab', t2, F1)
    r2 = E.oovv + x - x.transpose(0, 1, 3, 2)
    F2 = Foo + 0.5 * es('je,me->mj', t1, Fov)
    x = es('mj,imab->ijab', F2, t2)
    r2 -= x - x.transpose(1, 0, 2, 3)

    x = es('mnie,je->mnij', E.ooov, t1)
    Woo_p = pack_first(E.oooo + x - x.transpose(0, 1, 3, 2)).reshape(-1, o * o)   # [mn_p, (ij)]
    Woo_p = Woo_p + es('mf,if->mi', E.oovv_p, tau_q)
    r2 += es('mi,ma->ia', Woo_p, tau_r).reshape(o, o, v, v)       # hh ladder, dense output
    acc_q = es('if,af->ia', tau_q, E.vvvv_p)                      # [(ij), ab_p]
    ovvv_p2 = E.ovvv_p.reshape(o * v, -1)
    Y_q = -2.0 * es('if,qf->iq', tau_q, ovvv_p2).reshape(-1, o, v)
    Z = es('pma,mb->pab', Y_q, t1)
    acc_q += -0.5 * pack_last(Z - Z.transpose(0, 2, 1))
    r2 += unpack_last(acc_q, v).reshape(o, o, v, v)

    Wph = -0.5 * es('menf,nfjb->mejb', E.oovv_ph, t2ph2)
    Wph += es('mbef,jf->mejb', E.ovvv, t1)
    Wph += es('nb,mnje->mejb', t1, E.ooov)
    U = es('mnef,jf->mnej', E.oovv, t1)
    Wph -= es('nb,mnej->mejb', t1, U)
    Wph -= E.ovov_ph
    ring = es('iame,mejb->iajb', t2ph, Wph)
    Q = es('jbme,ie->jbmi', E.ovov_ph, t1)
    ring += es('ma,jbmi->iajb', t1, Q)
    r2 += antisym_ph(ring)
    x = -es('ie,jeab->ijab', t1, E.ovvv)
    r2 += x - x.transpose(1, 0, 2, 3)
    x = es('ma,ijmb->ijab', t1, E.ooov)
    r2 -= x - x.transpose(0, 1, 3, 2)
    return finish(r1, r2, t1, t2, e_o, e_v, alpha, equation)


def lupdate_general(E, fock, t1, t2, l1, l2, fsp=None, alpha=None, equation=False):
    o, v = t1.shape
    if fsp is None:
        fsp = fock
    foo, fov, fvo, fvv = fsp[:o, :o], fsp[:o, o:], fsp[o:, :o], fsp[o:, o:]
    e_o, e_v = np.diagonal(fock)[:o], np.diagonal(fock)[o:]
    shift = (equation is False) and alpha is None

    tau = make_tau2(t1, t2, 2.0, 0.0)                               # CCSD.py:565 exactly
    tau_q = 0.5 * pack_last(tau - tau.transpose(0, 1, 3, 2)).reshape(o * o, -1)
    l2_q = 0.5 * pack_last(l2 - l2.transpose(0, 1, 3, 2)).reshape(o * o, -1)
    t2ph = np.ascontiguousarray(t2.transpose(0, 2, 1, 3))
    l2ph = np.ascontiguousarray(l2.transpose(0, 2, 1, 3))          # [(ib),(jc)] = l2[i,j,b,c]
    l2ph2 = np.ascontiguousarray(l2.transpose(1, 3, 0, 2))         # [(ia),(kc)] = l2[k,i,c,a]

    G = es('menf,nf->me', E.oovv_ph, t1)
    Fov = fov + G
    v1 = fvv - es('ja,jb->ba', fov, t1)
    v1 = v1 - es('jbac,jc->ba', E.ovvv, t1)
    v1 = v1 + 0.5 * es('jkca,jkbc->ba', E.oovv, tau)
    v2 = foo + es('ib,jb->ij', fov, t1)
    v2 = v2 - es('kijb,kb->ij', E.ooov, t1)
    v2 = v2 + 0.5 * es('ikbc,jkbc->ij', E.oovv, tau)

    v4ph = es('kcld,ldjb->kcjb', t2ph, E.oovv_ph) - E.ovov_ph
    v5T = fvo.T + es('jbkc,kc->jb', t2ph, fov)
    q = es('kc,jc->kj', Fov, t1)
    v5T = v5T + es('kj,kb->jb', q, t1)
    v5T = v5T - 0.5 * es('kljc,klbc->jb', E.ooov, t2)
    v5T = v5T - 0.5 * es('jkdc,kbdc->jb', t2, E.ovvv)
    w3T = v5T + es('kcjb,jb->kc', v4ph, t1)
    w3T = w3T + es('kb,cb->kc', t1, v1)
    w3T = w3T - es('jk,jc->kc', v2, t1)

    # woooo packed on its antisymmetric (integral) first pair only: [ij_p, (kl)]
    y = es('jilc,kc->jilk', E.ooov, t1)
    woo_p = pack_first(0.5 * E.oooo + y).reshape(-1, o * o)
    woo_p = woo_p + 0.5 * es('if,kf->ik', E.oovv_p, tau_q)         # 1/4 v3, v3 = 2 oovv_p.tau_q
    lt = es('ijcd,klcd->ijkl', l2, tau)                            # dense, no symmetry

    S = es('ljbd,kd->ljbk', E.oovv, t1)
    wph = v4ph + es('lc,ljbk->kcjb', t1, S)
    wph = wph - es('lc,ljkb->kcjb', t1, E.ooov)
    wph = wph + es('jcbd,kd->kcjb', E.ovvv, t1)

    ovvv_p2 = E.ovvv_p.reshape(o * v, -1)
    wovoo = 0.5 * es('qf,kf->qk', ovvv_p2, tau_q).reshape(o, v, o, o)
    wovoo = wovoo + 0.5 * E.ooov.transpose(2, 3, 0, 1)
    wovoo = wovoo + es('kcib,jb->icjk', v4ph, t1)
    wovoo = wovoo - es('kclb,lijb->icjk', t2ph, E.ooov)

    # m3 dense [ij,ab]; pieces with an antisymmetric integral pair are accumulated packed
    m3 = unpack_first(es('ik,ka->ia', woo_p, l2.reshape(o * o, v * v)), o).reshape(o, o, v, v)
    ltp = pack_last(lt - lt.transpose(0, 1, 3, 2)).reshape(o * o, -1)     # [(ij), kl_p]
    acc_q = 0.25 * es('ik,ka->ia', ltp, E.oovv_p)
    l2t1 = es('ijcd,kd->ijck', l2, t1)
    a_q = np.ascontiguousarray(l2t1.transpose(0, 1, 3, 2)).reshape(o * o, o * v)
    acc_q += es('pq,qa->pa', a_q, ovvv_p2)
    acc_q += es('if,af->ia', l2_q, E.vvvv_p)
    m3 += unpack_last(acc_q, v).reshape(o, o, v, v)

    m_vv = 0.5 * es('klcb,klca->ba', t2, l2)
    m_oo = 0.5 * es('kicd,kjcd->ij', l2, t2)
    x_vv = m_vv + es('ka,kb->ba', l1, t1)
    x_oo = m_oo + es('ic,kc->ik', l1, t1)
    if shift:
        v1s = v1 - np.diag(e_v)
        v2s = v2 - np.diag(e_o)
    else:
        v1s, v2s = v1, v2

    r2 = E.oovv + m3
    ring = es('iakc,kcjb->iajb', l2ph2, wph) + es('ia,jb->iajb', l1, Fov)
    r2 += antisym_ph(ring)
    y = es('ka,ijkb->ijab', l1, E.ooov) + es('ijca,cb->ijab', l2, v1s)
    y += es('ca,ijcb->ijab', x_vv, E.oovv)
    r2 -= y - y.transpose(0, 1, 3, 2)
    y = es('qc,pcrs->pqrs', l1, E.ovvv)
    y = y + es('qk,kprs->pqrs', v2s, l2) - es('pk,kqrs->pqrs', x_oo, E.oovv)
    r2 += y - y.transpose(1, 0, 2, 3)

    r1 = fov - es('jbia,jb->ia', E.ovov_ph, l1)
    r1 += es('ib,ba->ia', l1, v1s) - es('ij,ja->ia', v2s, l1)
    r1 -= es('icjk,kjca->ia', wovoo, l2)
    l2t1b = es('ikbc,jb->ikcj', l2, t1)
    r1 -= es('ikcj,kcja->ia', l2t1b, v4ph)
    r1 -= 0.25 * es('ikjl,jlka->ia', lt, E.ooov)
    r1 -= 0.5 * es('ikbc,kabc->ia', l2, E.ovvv)
    Xph = es('ibjc,jckd->ibkd', l2ph, t2ph)
    r1 += es('ibkd,kbda->ia', Xph, E.ovvv)
    r1 += es('ijab,jb->ia', m3, t1)
    r1 += es('iajb,jb->ia', l2ph2, w3T)
    z = t1 + es('kcjb,kc->jb', t2ph, l1) - es('bd,jd->jb', x_vv, t1) - es('lj,lb->jb', m_oo, t1)
    r1 += es('iajb,jb->ia', E.oovv_ph, z)
    r1 -= es('icba,bc->ia', E.ovvv, x_vv)
    r1 -= es('jika,kj->ia', E.ooov, x_oo)
    r1 -= es('ik,ka->ia', m_oo, Fov)
    r1 -= es('ca,ic->ia', m_vv, Fov)

    if shift:
        Ecc = energy(E, t1, t2, fsp)
        r1 = r1 * (1.0 - Ecc)
        r2 = r2 * (1.0 - Ecc)
    return finish(r1, r2, l1, l2, e_o, e_v, alpha, equation)
