"""Host-side mirror of the reference `exp_pot.Exp` for the density-matrix target (`'mat'`, exp_pot.py:131-214):
the experimental potential of the ECW ground-state fit, Vexp[0,0] = L (rdm1_exp - rdm1_calc), its relative
deviation Delta (exp_pot.py:392-430) and the weight check (exp_pot.py:459-490).

n x n work on the host: the device-resident solver (`ecw_cc_b200.Solver_CCSD`) sends it the rdm1 (n x n) once per
iteration and takes the dressed Fock back — 3 MB of PCIe traffic per iteration at (40,400).  The property targets
('Ek', 'dip', 'v1e', 'F', 'trdip', 'DEk', 'trmat') need AO integrals from PySCF (exp_pot.py:72-108) and are out of
scope: the unchanged reference class can be passed to the solver instead whenever PySCF is available.
"""
import numpy as np


class Exp(object):
    def __init__(self, L, exp_data, mol=None, mo_coeff=None, Ek_exp_GS=None, Ek_HF_GS=None, HF_prop=False):
        """Same signature as the reference (exp_pot.py:11).  exp_data = [[['mat', rdm1_exp]], ...] (one entry per
        state); HF_prop = [[rdm1_HF]] switches Delta to the deviation relative to |rdm1_exp - rdm1_HF|."""
        self.nbr_states = len(exp_data)
        self.exp_data = exp_data
        self.mol, self.mo_coeff = mol, mo_coeff
        self.prop_names = []
        for st in exp_data:
            for prop in st:
                if prop[0] != 'mat':
                    raise NotImplementedError("only the 'mat' target is provided without PySCF (got %r)" % (prop[0],))
            self.prop_names.append([prop[0] for prop in st])
        if Ek_exp_GS is not None:
            raise NotImplementedError("Ek_exp_GS needs the kinetic-energy integrals (PySCF)")
        self.HF_prop = HF_prop if HF_prop else [[None for _ in st] for st in exp_data]
        self.L = self.L_check(L)
        self.Vexp = np.full((self.nbr_states, self.nbr_states), None)
        self.prop_calc = []

    def L_check(self, L):                                           # exp_pot.py:459-490
        if isinstance(L, (float, int)):
            return [[float(L)] * len(st) for st in self.exp_data]
        if isinstance(L, (list, np.ndarray)):
            if len(L) != self.nbr_states:
                raise SyntaxError('Given constrain weight length does not equal the number of states. '
                                  'You might have forgotten to put L_loop = True.')
            return [list(np.atleast_1d(l)) * (len(st) if len(np.atleast_1d(l)) == 1 else 1)
                    for st, l in zip(self.exp_data, L)]
        raise SyntaxError('L must be a number or a list with one entry per state')

    def Delta(self, n_st, i_prop, prop_diff):                       # exp_pot.py:414-423 ('mat' case)
        exp = self.exp_data[n_st][i_prop][1]
        hf = self.HF_prop[n_st][i_prop]
        if hf is None:
            return np.sum(abs(prop_diff)) / np.sum(abs(exp))
        return np.sum(abs(prop_diff)) / np.sum(abs(exp - hf))

    def Vexp_update(self, rdm1, rdm1_add, index, L=None):           # exp_pot.py:131-214
        n, m = index
        if n != m:
            raise NotImplementedError("transition targets ('trmat', 'trdip') are not provided")
        self.Vexp[n, m] = np.zeros_like(rdm1)
        Delta, vmax = 0., 0.
        self.prop_calc = []
        L = self.L if L is None else self.L_check(L)
        for i, _ in enumerate(self.prop_names[n]):
            diff = np.subtract(self.exp_data[n][i][1], rdm1)
            self.Vexp[n, n] += L[n][i] * diff
            Delta += self.Delta(n, i, diff)
            vmax += np.max(abs(diff))
        return Delta, vmax
