"""Config 3 of BASELINE.json in its named basis — H2O/6-31+G* ECW-CCS excited states with transition-dipole potentials
V^{0n}, V^{n0} (Solver_ES) — with the UNMODIFIED reference `Solver_ES.Solver_ES.SCF`, `CCS.Gccs` and `exp_pot.Exp` on
the integrals of ecw_cc_b200.molint (s, p and five spherical d functions per shell; (nocc, nvir) = (10, 34)).
Two valence states from Koopmans' guesses with the transition dipoles of test/Test_ECW_ES.py:43-44.
Build container only:

    python -m oracle.make_golden_h2o_es

tests/golden/h2o_631pgs_es.npz: the RHF solution (so the spin-orbital integrals can be rebuilt) and per case the
convergence text, Delta, right/left energies, ground-state rdm1 and all amplitudes.
"""
import os

import numpy as np

from . import ref_loader
from .make_golden_ccs_solvers import TRDIP, run_es
from .make_golden_h2o import H2O

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
BASIS = "6-31+g*"
CASES = [("es_trdip", TRDIP, ([2, 0], [0, 2]), "rl", "", [0.0, 0.01], 10),
         ("es_trdip_all", TRDIP, ([2, 0], [0, 2]), "all", "all", [0.05], 10)]


def water_diffuse(scf=None):
    from ecw_cc_b200 import molint
    mol = molint.Molecule(H2O, BASIS)
    ints = molint.integrals(mol)
    if scf is None:
        scf = molint.rhf(mol, ints)
    return mol, molint.geris(mol, tuple(scf[:3]) + (ints,)), scf


def main():
    CCS, Solver_ES, exp_pot, utilities = ref_loader.load("CCS", "Solver_ES", "exp_pot", "utilities")
    mol, er, scf = water_diffuse()
    print("H2O/%s: EHF %.10f, (nocc, nvir) = (%d, %d)" % (BASIS, scf[0], er.nocc, er.fock.shape[0] - er.nocc))
    out = {"EHF": scf[0], "mo_energy": scf[1], "mo_coeff": scf[2]}
    out.update(run_es(Solver_ES.Solver_ES, CCS.Gccs, exp_pot.Exp, utilities.koopman_init_guess, mol, er, cases=CASES))
    np.savez_compressed(os.path.join(OUT, "h2o_631pgs_es.npz"), **out)
    for k in sorted(out):
        if k.endswith("_text"):
            print(k, out[k])
        if k.endswith("_Ep"):
            print(k, out[k].tolist())


if __name__ == "__main__":
    main()
