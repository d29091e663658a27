"""Sampled-element CPU oracle of the CCSD T / Lambda residuals — TEST INFRASTRUCTURE ONLY.

The full oracle (oracle/ccsd_np.py) follows the reference's factorisation with its dense Wvvvv / wvvvo builds and
cannot run at the benchmark shape (nocc, nvir) = (40, 400): 600 GB.  This module evaluates, straight from the
reference's formulas (CCSD.py:248-338 tupdate, :346-413 T intermediates, :419-535 lupdate, :543-623 Linter),

    T1new (all of it),  T2new[:, :, a, b]  and  L2new[:, :, a, b]  for chosen virtual pairs (a, b),
    L1new[:, a]  for the virtual indices a of those pairs,

with integrals taken from a *provider* that produces blocks on demand (oracle/synth_fast.py: function-defined
synthetic integrals, never a dense vvvv).  Per pair the work is O(o^2 v^2) + O(o^3 v) for the ladders and rings and
O(o^2 v^3) for the Lambda pieces — seconds at (40, 400).  Intermediates that the reference forms as o v^3 / v^4 arrays
are evaluated only on the slices the chosen elements touch; where the reference contracts such an intermediate with a
singles amplitude, the product is re-associated (same terms, same coefficients).  Nothing here shares code, layouts or
plan lowering with the CUDA path.

Pinned by tests/test_oracle_columns_cpu.py: equal to oracle/ccsd_np.OracleGCC (itself pinned to the unmodified
reference) on every element and in every output mode at sizes where both run.
"""
import numpy as np

from .ccsd_np import soft_threshold

es = lambda spec, *ops: np.einsum(spec, *ops, optimize=True)  # noqa: E731


class ColumnOracle(object):
    def __init__(self, provider, pairs):
        """provider: oracle/synth_fast.SynthProvider | ArrayProvider;  pairs: list of (a, b) virtual index pairs."""
        self.p = provider
        self.o, self.v = provider.nocc, provider.nvir
        self.fock = provider.fock
        self.pairs = [(int(a), int(b)) for a, b in pairs]
        self.xs = sorted({x for ab in self.pairs for x in ab})        # virtual indices that occur in the pairs

    # ------------------------------------------------------------------------------------------------------
    def _finish(self, r1, r2cols, a1, a2, alpha, equation, cols2, cols1=None):
        """Output modes of CCSD.py:316-338 / :512-535 on the chosen elements.
        r1: (o, v) or (o, len(cols1)) singles residual, r2cols: (npair, o, o); a1 / a2 the input amplitudes."""
        o = self.o
        e = np.diagonal(self.fock)
        eo, ev = e[:o], e[o:]
        eia = eo[:, None] - ev[None, :]
        amp1 = a1 if cols1 is None else a1[:, cols1]
        d1 = eia if cols1 is None else eia[:, cols1]
        amp2 = np.stack([a2[:, :, a, b] for a, b in cols2])
        d2 = np.stack([eo[:, None] + eo[None, :] - ev[a] - ev[b] for a, b in cols2])
        if alpha is not None:
            w1, w2 = r1, soft_threshold(r2cols, amp2, alpha)             # L1 term on the doubles only (Q3)
            if equation:
                return w1, w2
            return (w1 + amp1 * d1) / d1, (w2 + amp2 * d2) / d2
        if equation:
            return r1, r2cols
        return r1 / d1, r2cols / d2

    # ------------------------------------------------------------------------------------------------------
    def tupdate(self, t1, t2, fsp=None, alpha=None, equation=False):
        """-> (T1new full (o, v), T2new columns (npair, o, o)); CCSD.py:248-338."""
        p, o, v = self.p, self.o, self.v
        fock = self.fock
        if fsp is None:
            fsp = fock
        foo, fov, fvv = fsp[:o, :o], fsp[:o, o:], fsp[o:, o:]
        oovv, ooov, ovov, oooo = p.oovv, p.ooov, p.ovov, p.oooo
        t1t1 = es('ia,jb->ijab', 0.5 * t1, t1)
        t1t1 = t1t1 - t1t1.transpose(1, 0, 2, 3)
        t1t1 = t1t1 - t1t1.transpose(0, 1, 3, 2)
        tau = t2 + t1t1                                                   # make_tau, :346-353
        ttl = t2 + 0.5 * t1t1                                             # fac = 0.5
        del t1t1
        # cc_Fvv / cc_Foo / cc_Fov (:355-387); the vovv term and the T1 ovvv term stream over ovvv[m]
        Fvv = fvv - 0.5 * es('me,ma->ae', fov, t1) - 0.5 * es('mnaf,mnef->ae', ttl, oovv)
        Foo = foo + 0.5 * es('me,ie->mi', fov, t1) + es('ne,mnie->mi', t1, ooov) + 0.5 * es('inef,mnef->mi', ttl, oovv)
        Fov = fov + es('nf,mnef->me', t1, oovv)
        del ttl
        t1_ovvv = np.zeros((o, v))
        for m in range(o):
            blk = p.ovvv_m(m, m + 1)[0]                                   # ovvv[m]  (a, e, f)
            Fvv -= es('f,aef->ae', t1[m], blk)                            # vovv[amef] = -ovvv[maef]
            t1_ovvv += np.ascontiguousarray(t2[:, m]).reshape(o, v * v) @ blk.reshape(v, v * v).T   # 'imef,maef->ia'
        if not equation and alpha is None:                                # :283-285
            Fvv = Fvv - np.diag(np.diagonal(fock[o:, o:]))
            Foo = Foo - np.diag(np.diagonal(fock[:o, :o]))
        r1 = es('ie,ae->ia', t1, Fvv) - es('ma,mi->ia', t1, Foo) + es('imae,me->ia', t2, Fov)
        r1 -= es('nf,naif->ia', t1, ovov)
        r1 -= 0.5 * t1_ovvv
        r1 -= 0.5 * es('mnae,mnie->ia', t2, ooov)
        r1 += fov
        # Woooo (:389-394)
        tmp = es('je,mnie->mnij', t1, ooov)
        Woooo = oooo + tmp - tmp.transpose(0, 1, 3, 2) + 0.25 * es('ijef,mnef->mnij', tau, oovv)
        Ft1 = Fvv - 0.5 * es('mb,me->be', t1, Fov)
        Ft2 = Foo + 0.5 * es('je,me->mj', t1, Fov)
        # per virtual index x: ovvv[:, x, :, :] and Wovvo[:, x, :, :] (:404-413)
        ovvv1, Wx = {}, {}
        oovv_x = np.ascontiguousarray(oovv.transpose(0, 2, 1, 3)).reshape(o * v, o * v)     # [(m,e),(n,f)], one copy
        for x in self.xs:
            A = ovvv1[x] = p.ovvv_x1(x)                                   # [m, e, f]
            W = es('jf,mef->mej', t1, A)
            W += es('n,mnje->mej', t1[:, x], ooov)                        # - t1[nb] oovo[mnej], oovo = -ooov^T
            W -= 0.5 * (oovv_x @ np.ascontiguousarray(t2[:, :, :, x]).reshape(o, o * v).T).reshape(o, v, o)
            W -= es('jf,n,mnef->mej', t1, t1[:, x], oovv)
            W -= ovov[:, x].transpose(0, 2, 1)                            # ovvo[mbej] = -ovov[mbje]
            Wx[x] = W
        out = np.empty((len(self.pairs), o, o))
        for k, (a, b) in enumerate(self.pairs):
            c = es('ije,e->ij', t2[:, :, a, :], Ft1[b]) - es('ije,e->ij', t2[:, :, b, :], Ft1[a])            # :297-299
            x2 = es('im,mj->ij', t2[:, :, a, b], Ft2)
            c -= x2 - x2.T                                                                                   # :300-302
            c += oovv[:, :, a, b]
            c += 0.5 * es('mn,mnij->ij', tau[:, :, a, b], Woooo)
            Wab = p.vvvv_ab(a, b) - es('m,mfe->fe', t1[:, a], ovvv1[b]) + es('m,mfe->fe', t1[:, b], ovvv1[a])  # :396-402
            Wab = Wab + 0.25 * es('mn,mnef->ef', tau[:, :, a, b], oovv)
            c += 0.5 * es('ijef,ef->ij', tau, Wab)

            def ring(a_, b_):                                                                                # :306-307
                r = es('ime,mej->ij', t2[:, :, a_, :], Wx[b_])
                return r + es('ie,m,mje->ij', t1, t1[:, a_], ovov[:, b_])
            r = ring(a, b) - ring(b, a)
            c += r - r.T
            y = es('ie,je->ij', t1, p.ovvv_ef(b, a))                                                         # :311
            c += y - y.T
            c -= es('m,ijm->ij', t1[:, a], ooov[:, :, :, b]) - es('m,ijm->ij', t1[:, b], ooov[:, :, :, a])   # :313-314
            out[k] = c
        return self._finish(r1, out, t1, t2, alpha, equation, self.pairs)

    # ------------------------------------------------------------------------------------------------------
    def lupdate(self, t1, t2, l1, l2, fsp=None, alpha=None, equation=False):
        """-> (L1new[:, xs] (o, len(xs)), L2new columns (npair, o, o)); CCSD.py:419-535 with Linter :543-623."""
        p, o, v = self.p, self.o, self.v
        fock = self.fock
        if fsp is None:
            fsp = fock
        foo, fov, fvo, fvv = fsp[:o, :o], fsp[:o, o:], fsp[o:, :o], fsp[o:, o:]
        oovv, ooov, ovov, oooo = p.oovv, p.ooov, p.ovov, p.oooo
        tau = t2 + 2.0 * es('ia,jb->ijab', t1, t1)                        # :565
        # ---- Linter pieces that are small or can be streamed
        v1 = fvv - es('ja,jb->ba', fov, t1) + 0.5 * es('jkca,jkbc->ba', oovv, tau)
        v5_ovvv = np.zeros((v, o))
        mba = 0.5 * es('klca,klcb->ba', l2, t2)                           # :459
        mij = 0.5 * es('kicd,kjcd->ij', l2, t2)
        tmp1vv = mba + es('ka,kb->ba', l1, t1)
        tmp1oo = mij + es('ic,kc->ik', l1, t1)
        l1_ovvv = np.zeros((o, v))
        for j in range(o):
            blk = p.ovvv_m(j, j + 1)[0]                                   # ovvv[j]  (b, a, c) / (b, d, c) / (c, a, b)
            v1 -= es('bac,c->ba', blk, t1[j])                             # :568
            tk = np.ascontiguousarray(t2[:, j].transpose(0, 2, 1)).reshape(o, v * v)          # [j', (d, c)]
            v5_ovvv += blk.reshape(v, v * v) @ tk.T                      # 'kbdc,jkcd->bj' with k = j of the loop
            l1_ovvv[j] = es('cab,bc->a', blk, tmp1vv)                     # 'icab,bc->ia'  (:501)
        v2 = foo + es('ib,jb->ij', fov, t1) - es('kijb,kb->ij', ooov, t1) + 0.5 * es('ikbc,jkbc->ij', oovv, tau)
        v3 = es('ijcd,klcd->ijkl', oovv, tau)
        g = fov - es('kldc,ld->kc', oovv, t1)
        v5 = fvo + es('kc,jkbc->bj', fov, t2) + es('kc,kb,jc->bj', g, t1, t1) - 0.5 * es('kljc,klbc->bj', ooov, t2)
        v5 = v5 + 0.5 * v5_ovvv
        # w3 = v5 + v4.t1 + v1.t1 - v2.t1 (:588-590); v4.t1 re-associated: sum_jb (oovv[ljdb] t1[jb]) t2[klcd] + ovvo.t1
        h = es('ljdb,jb->ld', oovv, t1)
        w3 = v5 + es('ld,klcd->ck', h, t2) - es('jckb,jb->ck', ovov, t1)
        w3 = w3 + es('cb,jb->cj', v1, t1) - es('jk,jb->bk', v2, t1)
        woooo = 0.5 * oooo + 0.25 * v3 + es('jilc,kc->jilk', ooov, t1)    # :592-594
        E = 0.0
        if equation is False and alpha is None:                           # :449-456 (Q2)
            v1s = v1 - np.diag(np.diagonal(fock[o:, o:]))
            v2s = v2 - np.diag(np.diagonal(fock[:o, :o]))
            E = es('ia,ia', fov, t1) + 0.25 * es('ijab,ijab', t2, oovv) + 0.5 * es('ia,jb,ijab', t1, t1, oovv)
        else:
            v1s, v2s = v1, v2
        lt = es('ijcd,klcd->ijkl', l2, tau)                               # :463
        l2t1 = es('ijcd,kd->ijck', l2, t1)                                # :466
        fov1 = fov + es('kjcb,kc->jb', oovv, t1)                          # :474
        # wovoo (:600-603) in full (o^3 v) but for its ovvv.tau term, which is contracted with l2 per column below;
        # the v4.t1 term is re-associated
        wovoo = 0.5 * ooov.transpose(2, 3, 0, 1) - es('lijb,klcb->icjk', ooov, t2)
        X = es('lidb,jb->lidj', oovv, t1)
        wovoo = wovoo + es('lidj,klcd->icjk', X, t2) - es('ickb,jb->icjk', ovov, t1)        # v4[icbk] t1[jb]
        t2_x = np.ascontiguousarray(t2.transpose(1, 3, 0, 2)).reshape(o * v, o * v)           # [(l,d),(k,c)], one copy
        # ---- per virtual index x: slices of v4 / wovvo / m3 / wvvvo
        ovvv2, wovvo_x, m3_x = {}, {}, {}
        l1cols = np.zeros((o, len(self.xs)))
        for n, x in enumerate(self.xs):
            B2 = ovvv2[x] = p.ovvv_x2(x)                                  # ovvv[:, :, x, :]   [j, c, d]
            ox = np.ascontiguousarray(oovv[:, :, :, x].transpose(1, 0, 2)).reshape(o, o * v)    # [j, (l,d)]
            v4x = (ox @ t2_x).reshape(o, o, v).transpose(0, 2, 1) - ovov[:, :, :, x]  # v4[j,c,x,k], ovvo[jcbk] = -ovov[jckb]
            w = v4x - es('ljd,lc,kd->jck', oovv[:, :, :, x], t1, t1) - es('ljk,lc->jck', ooov[:, :, :, x], t1)
            wovvo_x[x] = w + es('jcd,kd->jck', B2, t1)                    # wovvo[j,c,x,k]   (:596-598)
            # m3[:, :, x, :]  (:461-467)
            m3 = es('klb,ijkl->ijb', l2[:, :, x, :], woooo) + 0.25 * es('klb,ijkl->ijb', oovv[:, :, x, :], lt)
            m3 += es('kcb,ijck->ijb', B2, l2t1)                           # - ovvv[kcba] = + ovvv[kcab]
            m3 += 0.5 * es('ijcd,cdb->ijb', l2, p.vvvv_x3(x))
            m3_x[x] = m3
            # wvvvo[b, c, x, k]  (:605-608)
            wv = es('jck,jb->bck', v4x, t1) + 0.25 * es('jlk,jlbc->bck', ooov[:, :, :, x], tau)
            wv -= 0.5 * p.ovvv_x1(x).transpose(2, 1, 0)                   # - 1/2 ovvv[j, x, c, b] -> [b, c, j]
            wv += (np.ascontiguousarray(B2.transpose(1, 0, 2)).reshape(v, o * v) @ t2_x).reshape(v, o, v).transpose(0, 2, 1)
            c1 = fov[:, x] - es('ibj,jb->i', ovov[:, :, :, x], l1)        # l1[jb] ovvo[ibaj], ovvo[ibaj] = -ovov[ibja]
            c1 += es('ib,b->i', l1, v1s[:, x]) - es('j,ij->i', l1[:, x], v2s)
            c1 -= es('kjc,icjk->i', l2[:, :, :, x], wovoo)
            # the 1/4 ovvv[icdb] tau[jkdb] term of wovoo (:600): l2 and tau first, then one pass over ovvv
            Mx = 0.25 * es('kjc,jkdb->cdb', l2[:, :, :, x], tau)
            for i in range(o):
                c1[i] -= np.vdot(p.ovvv_m(i, i + 1)[0], Mx)
            c1 -= es('ikbc,bck->i', l2, wv)
            c1 += es('ijb,jb->i', m3, t1)
            c1 += es('jib,bj->i', l2[:, :, :, x], w3)
            z = t1 + es('kc,kjcb->jb', l1, t2) - es('bd,jd->jb', tmp1vv, t1) - es('lj,lb->jb', mij, t1)
            c1 += es('jib,jb->i', oovv[:, :, :, x], z)
            c1 += l1_ovvv[:, x]
            c1 -= es('jik,kj->i', ooov[:, :, :, x], tmp1oo)
            g2 = fov - es('kjba,jb->ka', oovv, t1)
            c1 -= es('ik,k->i', mij, g2[:, x]) + es('c,ic->i', mba[:, x], g2)
            l1cols[:, n] = c1
        out = np.empty((len(self.pairs), o, o))
        for k, (a, b) in enumerate(self.pairs):
            c = oovv[:, :, a, b] + m3_x[a][:, :, b]

            def ring(a_, b_):                                             # :475-476
                return np.outer(l1[:, a_], fov1[:, b_]) + es('kic,jck->ij', l2[:, :, :, a_], wovvo_x[b_])
            r = ring(a, b) - ring(b, a)
            c = c + r - r.T

            def vterm(a_, b_):                                            # :479-482
                r = es('k,ijk->ij', l1[:, a_], ooov[:, :, :, b_]) + es('ijc,c->ij', l2[:, :, :, a_], v1s[:, b_])
                return r + es('c,ijc->ij', tmp1vv[:, a_], oovv[:, :, :, b_])
            c = c - (vterm(a, b) - vterm(b, a))
            # :484-488; 'ic,jcba->jiba' read at [i,j,a,b]: l1[jc] ovvv[icab]
            y = es('jc,ic->ij', l1, ovvv2[a][:, :, b]) + es('ki,jk->ij', l2[:, :, a, b], v2s)
            y = y - es('ik,kj->ij', tmp1oo, oovv[:, :, a, b])
            c = c + y - y.T
            out[k] = c
        if E != 0.0:                                                      # :509-510
            l1cols = l1cols - l1cols * E
            out = out - out * E
        return self._finish(l1cols, out, l1, l2, alpha, equation, self.pairs, cols1=self.xs)
