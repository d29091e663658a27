"""HARNESS, not part of the drop-in (see ecw_cc_b200/harness/__init__.py).  Mirror of the reference's excited-state ECW-CCS solver `Solver_ES.Solver_ES` (Solver_ES.py:26-500, the `SCF`
method): coupled T / Lambda / R_n / L_n iteration for a ground state and N excited states with state (V_nn) and
ground-to-excited transition (V_0n, V_n0) experimental potentials.

Every tensor is o x v or n x n and every step is one call into `ecw_cc_b200.Gccs` (§8e: replicas only, latency bound),
so the loop drives the numpy API of `Gccs`; `Vexp` is `exp_pot.Exp` (the reference's or `ecw_cc_b200.exp_pot.Exp`).
DIIS: 'GS' and 'all' through `ecw_cc_b200.diis` (host store); the reference's 'ES' mode cannot run (its vector
bookkeeping fails on the first update, Solver_ES.py:376-388) and is rejected here.  `SCF_diag` sits on the reference's
dead Davidson path and is not provided.
"""
import copy

import numpy as np

from .. import utilities
from ..diis import DIIS


class Solver_ES(object):
    def __init__(self, mycc, Vexp, rn_ini=None, tsini=None, lsini=None, val_core=None, rini_koop_idx=None,
                 conv_var='tl', conv_thres=10 ** -6, diis='', maxiter=40, maxdiis=20, mindiis=2, tablefmt='rst'):
        """Parameters as in the reference (Solver_ES.py:27-44).  rn_ini=None takes Koopmans' guesses for
        val_core = (n_valence, n_core) states (the reference stops with a TypeError on that route, Solver_ES.py:89)."""
        self.mycc = mycc
        self.Vexp_class = Vexp
        self.nbr_states = Vexp.nbr_states
        self.tablefmt = tablefmt
        self.nocc, self.nvir = mycc.nocc, mycc.nvir
        self.dim = self.nocc + self.nvir
        self.EHF = mycc.eris.EHF
        self.tsini = np.zeros((self.nocc, self.nvir)) if tsini is None else tsini
        self.lsini = np.zeros((self.nocc, self.nvir)) if lsini is None else lsini
        e = np.diag(mycc.fock)
        if rn_ini is None:
            rn_ini, de = utilities.koopman_init_guess(e, mycc.eris.mo_occ, val_core, koop_idx=rini_koop_idx)
        elif len(rn_ini) != self.nbr_states - 1:
            raise ValueError('The number of given initial r vectors is not '
                             'consistent with the given experimental data for ES')
        else:
            de = [utilities.get_DE(e, rs) for rs in rn_ini]
        self.rn_ini = rn_ini
        self.ln_ini = [r * 1 for r in rn_ini]
        zero_t, zero_v = np.zeros_like(self.tsini), np.zeros((self.dim, self.dim))
        self.r0_ini = [mycc.r0_fromE(d, zero_t, r, zero_v) for r, d in zip(rn_ini, de)]
        self.l0_ini = [x * 1 for x in self.r0_ini]
        self.E_ini = -np.asarray(de)
        print(' Initial Koopman energies in eV: ', -self.E_ini * 27.2114)
        self.diis, self.maxdiis, self.mindiis = diis, maxdiis, mindiis
        self.maxiter, self.conv_thres = maxiter, conv_thres
        if conv_var not in ('Ep', 'rl', 'tl', 'all'):
            raise ValueError('Accepted convergence parameter is Ep, tl, rl or all')
        self.conv_var = conv_var

    # -- convergence vectors (Solver_ES.py:119-144) -------------------------------------------------------------------
    def _conv_vector(self, amp):
        if self.conv_var == 'Ep':
            return self.mycc.energy_ccs(amp['ts'], amp.get('fsp'))      # fsp is never stored: fails as in the reference
        tl = amp['ts'] + amp['ls']
        if self.conv_var == 'tl':
            return tl
        rl = np.zeros_like(amp['rn'][0])
        for r, l in zip(amp['rn'], amp['ln']):
            rl += r + l
        return rl if self.conv_var == 'rl' else tl + rl

    def SCF(self, L=None, dic_amp_ini=None, diis=None, force_alpha=True, print_ite=True):
        """Solver_ES.py:146-498.  Returns (text, {'ts','ls','rn','ln','r0n','l0n'}, Delta[nm], Ep[n, (right, left)],
        ground-state rdm1)."""
        V = self.Vexp_class
        ns = self.nbr_states
        nes = ns - 1
        L = V.L if L is None else V.L_check(L)
        if dic_amp_ini is None:
            ts, ls, rn, ln = self.tsini, self.lsini, self.rn_ini, self.ln_ini
            r0n, l0n = self.r0_ini, self.l0_ini
            ov = [np.where(r == 1) for r in rn]            # the Koopmans element pins the excitation energy
        else:
            ts, ls, rn, ln = (dic_amp_ini[k] for k in ('ts', 'ls', 'rn', 'ln'))
            r0n, l0n = dic_amp_ini['r0n'], dic_amp_ini['l0n']
            ov = [None] * nes
        amp = {'ts': ts, 'ls': ls, 'rn': rn, 'ln': ln}
        if diis is None:
            diis = self.diis
        if diis == 'ES':
            raise NotImplementedError("diis='ES' cannot run in the reference (Solver_ES.py:376-388); use 'GS' or 'all'")
        amp_diis = None
        if diis:
            amp_diis = DIIS()
            amp_diis.space, amp_diis.min_space = self.maxdiis, self.mindiis
        mycc, o, v = self.mycc, self.nocc, self.nvir
        fsp, rdm1, tr_rdm1 = [None] * ns, [None] * ns, [None] * nes
        Delta, Ep = np.zeros((ns, ns)), np.zeros((ns, 2))
        Spin = np.zeros(nes)
        conv, Dconv, ite = 0., 1., 0
        rows = []
        while Dconv > self.conv_thres:
            conv_old = conv
            # density matrices of all states: GS, ES_n, and both transition matrices of every ES
            rdm1[0] = mycc.gamma(ts, ls)
            for k in range(nes):
                rdm1[k + 1] = mycc.gamma_es(ts, ln[k], rn[k], r0n[k], l0n[k])
                tr_rdm1[k] = [mycc.gamma_tr(ts, ln[k], None, None, l0n[k]), mycc.gamma_tr(ts, ls, rn[k], r0n[k], 1)]
            # experimental potentials and dressed Fock matrices
            if V.exp_data[0]:
                Delta[0, 0], vmax = V.Vexp_update(rdm1[0], tr_rdm1, (0, 0), L=L)
            for n in range(1, ns):
                if not V.exp_data[n]:
                    continue
                if 'trdip' in V.prop_names[n] or 'trmat' in V.prop_names[n]:
                    right, left = tr_rdm1[n - 1]
                    Delta[n, 0], vmax = V.Vexp_update(right, left, (n, 0), L=L)
                    Delta[0, n], vmax = V.Vexp_update(left, right, (0, n), L=L)
                else:
                    Delta[n, n], vmax = V.Vexp_update(rdm1[n], rdm1[0], (n, n), L=L)
                    fsp[n] = np.subtract(mycc.fock, V.Vexp[n, n])
            if V.Vexp[0, 0] is not None:
                fsp[0] = np.subtract(mycc.fock, V.Vexp[0, 0])
            # ground state: T then Lambda, coupled to the excited states through V_0n / V_n0
            ts = mycc.tsupdate(ts, mycc.T1inter(ts, fsp[0]), rsn=rn, r0n=r0n, vn=V.Vexp[0, 1:])
            ls = mycc.lsupdate(ts, ls, mycc.L1inter(ts, fsp[0]), rsn=rn, lsn=ln, r0n=r0n, l0n=l0n, vn=V.Vexp[1:, 0])
            if diis == 'GS':
                ls, ts = (x.reshape(o, v) for x in np.split(amp_diis.update(np.concatenate((np.ravel(ls), np.ravel(ts)))), 2))
            # excited states: energy from the pinned element, then r, r0, l, l0
            rnew, lnew, r0new, l0new = [None] * nes, [None] * nes, [None] * nes, [None] * nes
            for n in range(1, ns):
                k = n - 1
                vexp = V.Vexp[0, n]
                inter = mycc.R1inter(ts, fsp[n], vexp)
                En_r, i, a = mycc.Extract_Em_r(rn[k], r0n[k], inter, ov=ov[k])
                rnew[k] = mycc.rsupdate(rn[k], r0n[k], inter, En_r, force_alpha=force_alpha)
                rnew[k][i, a] = np.ravel(mycc.get_ov(ln[k], l0n[k], rn[k], r0n[k], [i, a]))[0]
                r0new[k] = mycc.r0_fromE(En_r, ts, rn[k], vexp, fsp=fsp[n])
                vexp = V.Vexp[n, 0]
                inter = mycc.es_L1inter(ts, fsp[n], vexp)
                En_l, i, a = mycc.Extract_Em_l(ln[k], l0n[k], inter, ov=ov[k])
                lnew[k] = mycc.es_lsupdate(ln[k], l0n[k], En_l, inter, force_alpha=force_alpha)
                lnew[k][i, a] = np.ravel(mycc.get_ov(rn[k], r0n[k], ln[k], l0n[k], [i, a]))[0]
                l0new[k] = mycc.l0_fromE(En_l, ts, ln[k], vexp, fsp=fsp[n])
                Ep[n, 0], Ep[n, 1] = np.ravel(En_r)[0], np.ravel(En_l)[0]
            if diis == 'all':                              # one vector: ts, ls, all r, all l, all r0, all l0
                vec = np.concatenate((np.ravel(ts), np.ravel(ls), np.ravel([np.ravel(x) for x in rnew]),
                                      np.ravel([np.ravel(x) for x in lnew]), np.ravel([np.ravel(x) for x in r0new]),
                                      np.ravel([np.ravel(x) for x in l0new])))
                vec = amp_diis.update(vec)
                tail = vec[-2 * nes:]
                parts = np.split(vec[:-2 * nes], 2 * nes + 2)
                ts, ls = parts[0].reshape((o, v)), parts[1].reshape((o, v))
                for k in range(nes):
                    rnew[k], lnew[k] = parts[2 + k].reshape((o, v)), parts[2 + k + nes].reshape((o, v))
                    r0new[k], l0new[k] = tail[k], tail[nes + k]
            C_norm = utilities.check_ortho(lnew, rnew, r0new, l0new)
            for k in range(nes):
                Spin[k] = utilities.check_spin(rnew[k], lnew[k])
            rn, ln, r0n, l0n = (copy.deepcopy(x) for x in (rnew, lnew, r0new, l0new))
            amp = {'ts': ts, 'ls': ls, 'rn': rn, 'ln': ln, 'r0n': r0n, 'l0n': l0n}
            Ep[0, 0] = np.ravel(mycc.energy_ccs(ts, fsp[0], rsn=rn, r0n=r0n, vn=[V.Vexp[0, n] for n in range(1, ns)]))[0]
            conv = self._conv_vector(amp)
            if ite > 0:
                Dconv = np.linalg.norm(conv - conv_old)
            if print_ite:
                row = [ite, '%.3e' % Dconv]
                for k in range(nes):
                    row += ['ES %d' % (k + 1), '%.3e' % C_norm[k, k], Delta[k + 1, 0], Delta[0, k + 1], 2 * Spin[k] + 1,
                            float(np.ravel(r0n[k])[0]), float(np.ravel(l0n[k])[0]), Ep[k + 1, 0], Ep[k + 1, 1]]
                    if k:
                        row.append('%.3e' % ((C_norm[0, k] + C_norm[k, 0]) / 2))
                rows.append(row)
            if ite >= self.maxiter:
                Conv_text = 'Max iteration reached'
                break
            if Dconv > 10.:
                Conv_text = 'Diverges for lambda = {} after {} iterations'.format(L, ite)
                break
            ite += 1
        else:
            Conv_text = 'Convergence reached for lambda= {}, after {} iteration'.format(L, ite)
        if print_ite:
            self._print_table(rows, nes)
        return Conv_text, amp, Delta, Ep, rdm1[0]

    def _print_table(self, rows, nes):
        head = ['ite', 'Dconv ' + str(self.conv_var)]
        for k in range(nes):
            head += ['ES %d' % (k + 1), 'norm', 'Delta_r', 'Delta_l', '2S+1', 'r0', 'l0', 'Er', 'El']
            if k:
                head.append('Ortho wrt ES 1')
        try:
            from tabulate import tabulate
            print(tabulate(rows, head, tablefmt=self.tablefmt))
        except ImportError:
            print('  '.join(head))
            for row in rows:
                print('  '.join(str(x) for x in row))
