"""Golden runs of the UNMODIFIED reference ECW-CCS solvers on H2O/6-31G (integrals: ecw_cc_b200.molint):
`Solver_GS.Solver_CCS.SCF` (ground state, Solver_GS.py:101-239) and `Solver_ES.Solver_ES.SCF` (ground + excited states
with transition-dipole / state-property potentials, Solver_ES.py:146-498 — config 3 of BASELINE.json in the 6-31G
basis), driving the reference `CCS.Gccs` and `exp_pot.Exp`.  Build container only:

    python -m oracle.make_golden_ccs_solvers

tests/golden/ccs_solvers_h2o.npz: per case the convergence text, energies, Delta, final rdm1 and amplitudes.
"""
import os

import numpy as np

from . import ref_loader
from .make_golden_h2o import H2O
from .make_golden_solver import target_rdm1

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
# ground state: (tag, L, alpha, diis, maxiter)
GS_CASES = [("gs_L05", 0.05, None, "", 60), ("gs_L2_tl", 0.2, None, "tl", 60), ("gs_L05_rdm1", 0.05, None, "rdm1", 25),
            ("gs_L05_a", 0.05, 1e-3, "", 15)]
# excited states: (tag, exp_data builder, koopman (val_core, koop_idx), conv_var, diis, [L values run in sequence], maxiter)
TRDIP = [[], [['trdip', [0.523742, 0., 0.]]], [['trdip', [0., 0., 0.622534]]]]          # test/Test_ECW_ES.py:43-44


def es_cases(o, v):
    state = [[['mat', target_rdm1(o, v)]], [['dip', [0., 0., -0.9]], ['Ek', 75.6]]]
    return [("es_trdip", TRDIP, ([2, 0], [0, 2]), "rl", "", [0.0, 0.05], 12),
            ("es_trdip_all", TRDIP, ([2, 0], [0, 2]), "all", "all", [0.05], 12),
            ("es_state_gs", state, ([1, 0], [0]), "tl", "GS", [0.02], 12)]


def water():
    from ecw_cc_b200 import molint
    g = np.load(os.path.join(OUT, "h2o_631g.npz"))
    mol = molint.Molecule(H2O, "6-31g")
    return mol, molint.geris(mol, (float(g["EHF"]), g["mo_energy"], g["mo_coeff"], molint.integrals(mol)))


def run_gs(Solver_CCS, Gccs, Exp, er):
    o, v = er.nocc, er.fock.shape[0] - er.nocc
    out = {}
    for tag, L, alpha, diis, maxiter in GS_CASES:
        vx = Exp(L, [[["mat", target_rdm1(o, v)]]], None, None)
        text, ep, delta, conv, rdm1, (ts, ls) = Solver_CCS(Gccs(er), vx, conv="tl", conv_thres=1e-9, maxiter=maxiter,
                                                            maxdiis=8).SCF(L, alpha=alpha, diis=diis)
        out[tag + "_text"] = np.array(text)
        out[tag + "_Ep"], out[tag + "_Delta"], out[tag + "_conv"] = ep, delta, conv
        out[tag + "_rdm1"], out[tag + "_ts"], out[tag + "_ls"] = rdm1, ts, ls
    return out


def run_es(Solver_ES, Gccs, Exp, koopman, mol, er, cases=None):
    o, v = er.nocc, er.fock.shape[0] - er.nocc
    out = {}
    for tag, exp_data, (val_core, kidx), conv_var, diis, Ls, maxiter in (cases or es_cases(o, v)):
        rn, _ = koopman(er.mo_energy, er.mo_occ, val_core, koop_idx=kidx)
        vx = Exp(Ls[0], exp_data, mol, er.mo_coeff_g)
        solver = Solver_ES(Gccs(er), vx, rn_ini=rn, conv_var=conv_var, conv_thres=1e-9, maxiter=maxiter, diis=diis,
                           maxdiis=6, mindiis=2)
        amp = None
        for k, L in enumerate(Ls):
            text, amp, delta, ep, rdm1 = solver.SCF(L=L, dic_amp_ini=amp, print_ite=False)
            key = "%s_%d" % (tag, k)
            out[key + "_text"] = np.array(text)
            out[key + "_Delta"], out[key + "_Ep"], out[key + "_rdm1"] = np.array(delta), np.array(ep), rdm1
            out[key + "_ts"], out[key + "_ls"] = amp["ts"], amp["ls"]
            out[key + "_rn"], out[key + "_ln"] = np.array(amp["rn"]), np.array(amp["ln"])
            out[key + "_r0n"] = np.array([float(np.ravel(x)[0]) for x in amp["r0n"]])
            out[key + "_l0n"] = np.array([float(np.ravel(x)[0]) for x in amp["l0n"]])
    return out


def main():
    CCS, Solver_GS, Solver_ES, exp_pot, utilities = ref_loader.load("CCS", "Solver_GS", "Solver_ES", "exp_pot", "utilities")
    mol, er = water()
    out = run_gs(Solver_GS.Solver_CCS, CCS.Gccs, exp_pot.Exp, er)
    out.update(run_es(Solver_ES.Solver_ES, CCS.Gccs, exp_pot.Exp, utilities.koopman_init_guess, mol, er))
    np.savez_compressed(os.path.join(OUT, "ccs_solvers_h2o.npz"), **out)
    for k in sorted(out):
        if k.endswith("_text"):
            print(k, out[k])
        if k.endswith("_Ep") and out[k].ndim == 2:
            print(k, out[k].tolist())
        if k.endswith("_Ep") and out[k].ndim == 1:
            print(k, out[k][-1])


if __name__ == "__main__":
    main()
