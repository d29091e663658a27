"""Config 2 of BASELINE.json in its NAMED basis — C2H2/cc-pVDZ, (nocc, nvir) = (14, 62) spin orbitals — driven like
`Main.ECW.CCSD_GS` (Main.py:730-763) by the UNMODIFIED reference solver / CCSD.GCC / exp_pot.Exp on the integrals of
ecw_cc_b200.molint: the first two weights of the 6-31G sweep (oracle/make_golden_c2h2.py: L = 0 and 0.2333, the second
started from the amplitudes of the first) without the L1 term, and one L1-regularised run (alpha = 2e-4) at L = 0.2333
cut at 6 iterations (the reference's Python `subdiff` loop costs 6 s per call at this size).  Build container only:

    python -m oracle.make_golden_c2h2_ccpvdz          (about 5 minutes, 1 minute of it integral evaluation)

The doubles amplitudes of the plain sweep stay antisymmetric and are stored packed (i<j, a<b)."""
import os

import numpy as np

from . import ref_loader
from .make_golden_c2h2 import LARRAY, OUT, acetylene, pack, sweep

BASIS = "cc-pvdz"
LS_PLAIN = LARRAY[:2]
L1_CASE = (LARRAY[1], 2e-4, 6)          # L, alpha, maxiter


def pack_pairs(x):
    o, v = x.shape[0], x.shape[2]
    i, j = np.triu_indices(o, 1)
    a, b = np.triu_indices(v, 1)
    return x[i[:, None], j[:, None], a[None, :], b[None, :]]


def shrink(out, tag):
    """final doubles -> packed pairs (+ the antisymmetry defect that justifies it)."""
    for name in ("td", "ld"):
        x = out.pop(tag + "_final_" + name)
        defect = max(np.abs(x + x.transpose(1, 0, 2, 3)).max(), np.abs(x + x.transpose(0, 1, 3, 2)).max())
        out[tag + "_final_" + name + "_p"] = pack_pairs(x)
        out[tag + "_final_" + name + "_defect"] = defect


def main():
    CCSD, Solver_GS, exp_pot = ref_loader.load("CCSD", "Solver_GS", "exp_pot")
    mol, er, scf = acetylene(basis=BASIS)
    o, v = er.nocc, er.fock.shape[0] - er.nocc
    assert (o, v) == (14, 62)
    out = {"EHF": scf[0], "mo_energy": scf[1], "mo_coeff": scf[2]}
    res = sweep(Solver_GS.Solver_CCSD, CCSD.GCC, exp_pot.Exp, er, None, larray=LS_PLAIN)
    pack(res, "plain", out)
    shrink(out, "plain")
    for L, r in zip(LS_PLAIN, res):
        print("C2H2/cc-pVDZ plain L=%.3f: %s | Ep %.10f | Delta %.6f" % (L, r[0], r[1][-1], r[2][-1][0]))
    L, alpha, maxiter = L1_CASE
    res = sweep(Solver_GS.Solver_CCSD, CCSD.GCC, exp_pot.Exp, er, alpha, larray=[L], maxiter=maxiter)
    pack(res, "l1", out)
    for name in ("td", "ld"):               # L1 breaks the antisymmetry (Q11): keep a strided sample + norms
        x = out.pop("l1_final_" + name)
        out["l1_final_%s_sample" % name] = x[::3, ::3, ::5, ::5].copy()
        out["l1_final_%s_norm" % name] = np.linalg.norm(x)
        out["l1_final_%s_nnz" % name] = np.count_nonzero(x)
    print("C2H2/cc-pVDZ l1 L=%.3f: %s | Ep %.10f | Delta %.6f" % (L, res[0][0], res[0][1][-1], res[0][2][-1][0]))
    np.savez_compressed(os.path.join(OUT, "c2h2_ccpvdz.npz"), **out)


if __name__ == "__main__":
    main()
