"""Host-side mirror of the reference `exp_pot.Exp` (exp_pot.py:11-490): the experimental potentials Vexp[n, m] that
dress the Fock matrix of the ECW fit, their relative deviations Delta, and the weight check.

Targets (exp_pot.py:131-345): 'mat' (ground or excited state rdm1), 'trmat' (left/right transition rdm1), 'Ek', 'v1e',
'dip' (state properties), 'DEk..' (kinetic-energy difference to the ground state, acts on Vexp[0, 0]) and 'trdip'
(squared transition dipole, acts on Vexp[0, n] / Vexp[n, 0]).  The structure-factor target 'F' needs Fourier-transformed
AO pairs (PySCF `ft_ao`) and is not provided.  Property targets take their AO integrals from `mol`, which may be a PySCF
`Mole` or `ecw_cc_b200.molint.Molecule` (same `intor_symmetric` / `with_common_orig` / `atom_charges` / `atom_coords`).

Scope: only the 'mat' ground-state target is part of the product (`mat_update_device`, `ecw_vexp_mat`).
The other targets are n x n host code kept as test HARNESS (callers of the path, SURVEY §2 "out of scope"): they let the
excited-state configuration of BASELINE.json be driven on a box without PySCF, and are pinned to the reference class
in tests/test_exp_pot_cpu.py.

n x n work on the host, except for the case `Main.CCSD_GS` runs — one density-matrix target for the ground state —
which `mat_update_device` evaluates on the GPU (`ecw_vexp_mat`): the device-resident solver (`ecw_cc_b200.Solver_CCSD`)
then moves only scalars per iteration; for every other target it sends the rdm1 (n x n) to the host once per iteration
and takes the dressed Fock back (3 MB of PCIe traffic per iteration at (40,400)).
"""
import numpy as np

from . import utilities


class Exp(object):
    def __init__(self, L, exp_data, mol=None, mo_coeff=None, Ek_exp_GS=None, Ek_HF_GS=None, HF_prop=False):
        """Same signature as the reference (exp_pot.py:11).  exp_data = [[GS targets], [ES1 targets], ...], a target
        being ['mat', rdm1], ['trmat', (left, right)], ['Ek', x], ['v1e', x], ['dip', [x, y, z]], ['trdip', [x, y, z]],
        ['DEk..', x]; mo_coeff in spin-orbital (G) format; HF_prop in the layout of exp_data switches Delta to the
        deviation relative to |exp - HF|."""
        self.nbr_states = len(exp_data)
        self.exp_data = exp_data
        self.mol, self.mo_coeff = mol, mo_coeff
        self.prop_calc = []
        self.HF_prop = HF_prop if HF_prop else [[None for _ in st] for st in exp_data]
        self.Ek_HF_GS = Ek_HF_GS
        self.L = self.L_check(L)
        self.charge_center = None
        self.Ek_int = self.dip_int = self.v1e_int = self.F_int = None
        self.dic_int = {}
        self.prop_names = []
        for st in exp_data:
            for prop in st:
                name = prop[0]
                if name == 'F':
                    raise NotImplementedError("the structure-factor target 'F' needs PySCF's ft_ao")
                if name not in ('mat', 'trmat') and mol is None:
                    raise ValueError("target %r needs AO integrals: pass mol (PySCF Mole or molint.Molecule)" % (name,))
                if 'dip' in name and self.dip_int is None:          # 'dip' and 'trdip' (exp_pot.py:90-98)
                    charges, coords = mol.atom_charges(), mol.atom_coords()
                    self.charge_center = np.einsum('z,zr->r', charges, coords) / charges.sum()
                    with mol.with_common_orig(self.charge_center):
                        self.dip_int = mol.intor_symmetric('int1e_r', comp=3)
                    self.dic_int['dip'] = utilities.convert_aoint(self.dip_int, mo_coeff)
                if 'v1e' in name and self.v1e_int is None:          # exp_pot.py:101-104
                    self.v1e_int = mol.intor_symmetric('int1e_nuc')
                    self.dic_int['v1e'] = utilities.convert_aoint(self.v1e_int, mo_coeff)
                if 'Ek' in name and self.Ek_int is None:            # 'Ek' and 'DEk..' (exp_pot.py:107-110)
                    self.Ek_int = mol.intor_symmetric('int1e_kin')
                    self.dic_int['Ek'] = utilities.convert_aoint(self.Ek_int, mo_coeff)
            self.prop_names.append([prop[0] for prop in st])
        self.DEk_GS_idx = None                                      # weight slot of the GS 'DEk' target, if any
        for i, name in enumerate(self.prop_names[0] if self.prop_names else []):
            if 'DEk' in name:
                self.DEk_GS_idx = i
        self.Ek_exp_GS = Ek_exp_GS
        self.Ek_calc_GS = None
        self.Delta_Ek_GS = None
        self.Vexp = np.full((self.nbr_states, self.nbr_states), None)

    # -- weights (exp_pot.py:459-490) -------------------------------------------------------------------------------
    def L_check(self, L):
        if isinstance(L, (float, int)):
            return [[float(L)] * len(st) for st in self.exp_data]
        if isinstance(L, (list, np.ndarray)):
            if len(L) != self.nbr_states:
                raise SyntaxError('Given constrain weight length does not equal the number of states. '
                                  'You might have forgotten to put L_loop = True.')
            out = []
            for k, (st, l) in enumerate(zip(self.exp_data, L)):
                l = list(np.atleast_1d(l))
                if len(st) != len(l) and len(l) == 1:               # one weight for all targets of the state
                    l = l * len(st)
                elif len(st) != len(l):
                    raise SyntaxError("Wrong syntax for L list")
                out.append(l)
            return out
        raise SyntaxError('L must be a number or a list with one entry per state')

    # -- <A> and <A>_nm <A>_mn (exp_pot.py:347-390) -----------------------------------------------------------------
    def calc_prop(self, prop, rdm1, g_format=True, rdm1_add=None):
        fn, ints = {'Ek': (utilities.Ekin, self.Ek_int), 'v1e': (utilities.v1e, self.v1e_int),
                    'dip': (utilities.dipole, self.dip_int)}.get(prop, (None, None))
        if fn is None:
            raise NotImplementedError('The possible properties are: Ek, v1e and dip')
        one = fn(self.mol, rdm1, g_format, False, self.mo_coeff, ints)
        if rdm1_add is None:
            return list(one) if prop == 'dip' else one
        two = fn(self.mol, rdm1_add.transpose(), g_format, False, self.mo_coeff, np.conj(ints))
        return (list(one * two), list(two)) if prop == 'dip' else (one * two, two)

    # -- relative deviation (exp_pot.py:392-446) --------------------------------------------------------------------
    def Delta(self, n_st, i_prop, prop_diff, comp_idx=1, threshold=10 ** -6):
        exp = self.exp_data[n_st][i_prop][1]
        hf = self.HF_prop[n_st][i_prop]
        if isinstance(prop_diff, np.ndarray) and n_st == 0:        # ground-state density matrix
            ref = exp if hf is None else exp - hf
            return np.sum(abs(prop_diff)) / np.sum(abs(ref))
        if isinstance(exp, list) and abs(exp[comp_idx]) > threshold:          # vector property, one component
            return prop_diff / np.abs(exp[comp_idx] if hf is None else exp[comp_idx] - hf[comp_idx])
        if isinstance(exp, float) and abs(exp) > threshold:                   # scalar property
            return prop_diff / np.abs(exp if hf is None else exp - hf)
        return 0.                                                   # incl. excited-state 'mat' (as in the reference)

    # -- Vexp[n, m] (exp_pot.py:131-345) ----------------------------------------------------------------------------
    def Vexp_update(self, rdm1, rdm1_add, index, L=None):
        n, m = index
        self.Vexp[n, m] = np.zeros_like(rdm1)
        Delta, vmax = 0., 0.
        self.prop_calc = []
        L = self.L if L is None else self.L_check(L)
        st = max(n, m)                                              # the state whose target list applies
        for i, name in enumerate(self.prop_names[st]):
            w = L[st][i]
            target = self.exp_data[st][i][1]
            if name == 'mat' and n == m:
                diff = np.subtract(target, rdm1)
                self.Vexp[n, n] += w * diff
                Delta += self.Delta(n, i, diff)
                vmax += np.max(abs(diff))
                if n == 0 and self.Ek_exp_GS is not None:           # kinetic energy of the fitted ground state
                    self.Ek_calc_GS = utilities.Ekin(self.mol, rdm1, aobasis=False, mo_coeff=self.mo_coeff,
                                                     ek_int=self.Ek_int, g=True)
                    ref = self.Ek_exp_GS if self.Ek_HF_GS is None else self.Ek_exp_GS - self.Ek_HF_GS
                    self.Delta_Ek_GS = np.abs(self.Ek_exp_GS - self.Ek_calc_GS) / np.abs(ref)
            if name == 'trmat' and n != m:
                print("WARNING: The use of experimental transition density matrix is not tested")
                if n != 0 and m != 0:
                    raise ValueError("Only transition properties between GS and ES are implemented: m or n must be = 0")
                diff = np.subtract(target[0] if n == 0 else target[1], rdm1)
                self.Vexp[n, m] += w * diff
                # the reference sums |target| over ROWS only (builtin sum), so Delta becomes a vector here
                avg = np.sum(abs(target[1]), axis=0) + np.sum(abs(target[0]), axis=0)
                Delta += np.sum(abs(diff)) / (avg / 2.)
                vmax += np.max(abs(diff))
            if name in ('Ek', 'v1e') and n == m:
                calc = self.calc_prop(name, rdm1)
                diff = np.abs(target - calc)
                Delta += self.Delta(n, i, diff)
                diff = diff * self.dic_int[name]
                self.Vexp[n, n] += w * diff
                vmax += np.max(abs(diff))
                self.prop_calc.append([name, calc])
            if 'DEk' in name and n == m and n != 0:                 # -ES rdm1 + GS rdm1; feeds the GS potential
                calc = self.calc_prop('Ek', np.subtract(rdm1_add, rdm1))
                diff = np.abs(target - calc)
                Delta += self.Delta(st, i, diff)
                diff = diff * self.dic_int['Ek']
                if self.Vexp[0, 0] is None:
                    self.Vexp[0, 0] = 0.
                self.Vexp[0, 0] += (L[0][self.DEk_GS_idx] if self.DEk_GS_idx is not None else w) * diff
                vmax += np.max(np.abs(diff))
                self.prop_calc.append([name, calc])
            if name == 'dip' and n == m:
                calc = self.calc_prop('dip', rdm1)
                for j in range(3):
                    diff = np.abs(target[j] - calc[j])
                    Delta += self.Delta(st, i, diff, comp_idx=j)
                    diff = diff * self.dic_int['dip'][j]
                    self.Vexp[n, m] += w * diff
                    vmax += np.max(np.abs(diff))
                self.prop_calc.append([name, calc])
            if name == 'trdip' and n != m:
                calc, scale = self.calc_prop('dip', rdm1, rdm1_add=rdm1_add)
                for j in range(3):
                    diff = np.abs(target[j] - calc[j])
                    Delta += self.Delta(st, i, diff, comp_idx=j)
                    diff = diff * (self.dic_int['dip'][j] * scale[j])
                    self.Vexp[n, m] += w * diff
                    vmax += np.max(np.abs(diff))
                self.prop_calc.append([name, calc])
        return Delta, vmax

    # -- the 'mat' ground-state target without leaving the GPU ------------------------------------------
    def device_mat_ready(self):
        """True when Vexp[0, 0] is exactly one density-matrix target (what `Main.CCSD_GS` fits): then the potential and
        the dressed Fock matrix can be formed on the device by `mat_update_device`."""
        return bool(self.prop_names) and self.prop_names[0] == ['mat'] and self.Ek_exp_GS is None

    def mat_update_device(self, rdm1_dev, fock_dev, L=None):
        """Device version of `Vexp_update(rdm1, rdm1, (0, 0), L)` + `fsp = fock - Vexp[0, 0]` (exp_pot.py:185-195,
        Solver_GS.py:690-692) through `ecw_vexp_mat`: returns (Delta, vmax, fsp) with fsp a device tensor; only the two
        reduction results (16 bytes) come to the host.  `self.Vexp[0, 0]` holds the DEVICE potential afterwards; call
        `sync_host()` to turn it into the numpy array the host path leaves there."""
        import torch
        from ._lib import lib, EcwError
        if not self.device_mat_ready():
            raise EcwError("mat_update_device needs a single 'mat' ground-state target")
        L = self.L if L is None else self.L_check(L)
        dev = rdm1_dev.device
        st = getattr(self, "_dev", None)
        if st is None or st["device"] != dev:
            target = np.ascontiguousarray(self.exp_data[0][0][1], dtype=np.float64)
            hf = self.HF_prop[0][0]
            st = {"device": dev, "target": torch.from_numpy(target).to(dev),
                  "norm": float(np.sum(abs(target if hf is None else target - hf))),
                  "vexp": torch.empty_like(rdm1_dev), "stats": torch.zeros(2, dtype=torch.float64, device=dev)}
            self._dev = st
        if tuple(rdm1_dev.shape) != tuple(st["target"].shape) or not rdm1_dev.is_contiguous():
            raise ValueError("rdm1 must be a contiguous device matrix of the target's shape")
        fsp = torch.empty_like(rdm1_dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.ecw_vexp_mat(rdm1_dev.data_ptr(), st["target"].data_ptr(), fock_dev.data_ptr(), float(L[0][0]),
                              st["vexp"].data_ptr(), fsp.data_ptr(), st["stats"].data_ptr(), rdm1_dev.numel(), stream)
        if rc != 0:
            raise EcwError("ecw_vexp_mat failed")
        self.Vexp[0, 0] = st["vexp"]
        self.prop_calc = []
        s = st["stats"].cpu()
        return float(s[0]) / st["norm"], float(s[1]), fsp

    def sync_host(self):
        """Replace a device-resident Vexp[0, 0] by its numpy copy."""
        v = self.Vexp[0, 0]
        if v is not None and not isinstance(v, np.ndarray) and hasattr(v, "cpu"):
            self.Vexp[0, 0] = v.cpu().numpy()
