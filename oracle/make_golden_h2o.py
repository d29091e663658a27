"""Config 1 of BASELINE.json — H2O/6-31G L1-ECW-CCSD ground state fitted to a target rdm1 — with the UNMODIFIED
reference solver (Solver_GS.Solver_CCSD.SCF), CCSD.GCC and exp_pot.Exp on the integrals of ecw_cc_b200.molint
(the PySCF-free integral source; E_HF anchor -75.9839, ECW_CC/__init__.py:39).  Build container only:

    python -m oracle.make_golden_h2o

tests/golden/h2o_631g.npz holds the RHF solution (so that the spin-orbital integrals can be rebuilt bit-compatibly:
eigenvector signs are LAPACK-build dependent) and, per (L, alpha) case, the histories and final states.
"""
import os

import numpy as np

from . import ref_loader
from .make_golden_solver import target_rdm1

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
H2O = [(8, (0., 0., 0.)), (1, (0., -0.757, 0.587)), (1, (0., 0.757, 0.587))]          # Main.py:104-109
CASES = [("L0", 0.0, None, 60), ("L05", 0.05, None, 60), ("L05_a", 0.05, 2e-4, 12)]
# DIIS-accelerated runs of the same solver (tag, L, alpha, maxiter, diis, maxdiis); pyscf.lib.diis comes from
# oracle/pyscf_stub (restated published algorithm — "parity unpinned" against PySCF itself)
DIIS_CASES = [("L0_dtl", 0.0, None, 60, "tl", 15), ("L05_dtl", 0.05, None, 60, "tl", 4),
              ("L05_drdm1", 0.05, None, 60, "rdm1", 15), ("L05_a_dboth", 0.05, 2e-4, 12, ("tl", "rdm1"), 6)]


def main():
    from ecw_cc_b200 import molint
    CCSD, Solver_GS, exp_pot = ref_loader.load("CCSD", "Solver_GS", "exp_pot")
    mol = molint.Molecule(H2O, "6-31g")
    scf = molint.rhf(mol)
    er = molint.geris(mol, scf)
    o, v = er.nocc, er.fock.shape[0] - er.nocc
    out = {"nocc": o, "nvir": v, "EHF": scf[0], "mo_energy": scf[1], "mo_coeff": scf[2], "conv_thres": 1e-9}
    for tag, L, alpha, maxiter, diis, maxdiis in [c + ("", 15) for c in CASES] + DIIS_CASES:
        mycc = CCSD.GCC(er)
        vx = exp_pot.Exp(L, [[["mat", target_rdm1(o, v)]]], None, None)
        text, ep, delta, conv, rdm1, amps = Solver_GS.Solver_CCSD(mycc, vx, conv="tl", conv_thres=1e-9, maxiter=maxiter,
                                                                  maxdiis=maxdiis).SCF(L, alpha=alpha, diis=diis)
        print("H2O/6-31G %s: %s | E_corr %.10f | Delta %.6f" % (tag, text, ep[-1], delta[-1][0]))
        out[tag + "_text"] = np.array(text)
        out[tag + "_Ep"], out[tag + "_Delta"], out[tag + "_conv"], out[tag + "_rdm1"] = ep, delta, conv, rdm1
        for k, a in zip(("ts", "ls", "td", "ld"), amps):
            out[tag + "_" + k] = a
    np.savez_compressed(os.path.join(OUT, "h2o_631g.npz"), **out)


if __name__ == "__main__":
    main()
