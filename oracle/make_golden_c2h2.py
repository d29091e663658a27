"""Config 2 of BASELINE.json in the 6-31G basis — C2H2 L1-ECW-CCSD ground state with a sweep of the Vexp weight L —
as `Main.ECW.CCSD_GS` drives it (Main.py:730-763): one `Solver_CCSD`, one `exp_pot.Exp` with the HF reference values
(`HF_prop`, test/Test_ECW_GS.py:34), `SCF(L, ts, ls, td, ld, alpha)` per L with the previous amplitudes as the start.
UNMODIFIED reference solver / CCSD.GCC / exp_pot.Exp on the integrals of ecw_cc_b200.molint.  Build container only:

    python -m oracle.make_golden_c2h2

(6-31G instead of cc-pVDZ to keep the reference sweep at minutes: the sweep logic and the kernels are basis independent.)
"""
import os

import numpy as np

from . import ref_loader
from .make_golden_solver import target_rdm1

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
C2H2 = [(6, (0., 0., 0.6034010)), (6, (0., 0., -0.6034010)), (1, (0., 0., 1.6667490)), (1, (0., 0., -1.6667490))]   # Main.py:65-70
LARRAY = np.linspace(0., 0.7, 4)                                   # lambi, lambf of test/Test_ECW_GS.py:9-12
SWEEPS = [("plain", None), ("l1", 2e-4)]                           # (tag, alpha)
CONV, MAXITER = 1e-8, 25


def sweep(Solver_CCSD, GCC, Exp, er, alpha, device=False, larray=None, maxiter=MAXITER):
    """The L loop of Main.CCSD_GS; `device` keeps the amplitudes on the GPU between L values (product only)."""
    larray = LARRAY if larray is None else larray
    o, v = er.nocc, er.fock.shape[0] - er.nocc
    hf = np.diag(er.mo_occ)
    vx = Exp(larray[0], [[["mat", target_rdm1(o, v)]]], None, None, HF_prop=[[hf]])
    solver = Solver_CCSD(GCC(er), vx, conv="tl", conv_thres=CONV, tsini=np.zeros((o, v)), lsini=np.zeros((o, v)),
                         diis="tl", maxdiis=15, maxiter=maxiter)
    ts, ls, td, ld = np.zeros((o, v)), np.zeros((o, v)), None, None
    out = []
    for L in larray:
        kw = {"return_device": True} if device else {}
        res = solver.SCF(L, ts=ts, ls=ls, td=td, ld=ld, alpha=alpha, **kw)
        ts, ls, td, ld = res[5]
        out.append(res)
    return out


def pack(results, tag, out):
    for k, res in enumerate(results):
        key = "%s_L%d" % (tag, k)
        out[key + "_text"] = np.array(res[0])
        out[key + "_Ep"], out[key + "_Delta"], out[key + "_conv"], out[key + "_rdm1"] = res[1], res[2], res[3], res[4]
    last = results[-1][5]
    for name, a in zip(("ts", "ls", "td", "ld"), last):
        out[tag + "_final_" + name] = np.asarray(a.cpu().numpy() if hasattr(a, "cpu") else a)


def acetylene(scf=None, basis="6-31g"):
    from ecw_cc_b200 import molint
    mol = molint.Molecule(C2H2, basis)
    ints = molint.integrals(mol)
    if scf is None:
        scf = molint.rhf(mol, ints)
    return mol, molint.geris(mol, tuple(scf[:3]) + (ints,)), scf


def main():
    CCSD, Solver_GS, exp_pot = ref_loader.load("CCSD", "Solver_GS", "exp_pot")
    mol, er, scf = acetylene()
    out = {"EHF": scf[0], "mo_energy": scf[1], "mo_coeff": scf[2]}
    for tag, alpha in SWEEPS:
        res = sweep(Solver_GS.Solver_CCSD, CCSD.GCC, exp_pot.Exp, er, alpha)
        pack(res, tag, out)
        for L, r in zip(LARRAY, res):
            print("C2H2/6-31G %s L=%.3f: %s | Ep %.10f | Delta %.6f" % (tag, L, r[0], r[1][-1], r[2][-1][0]))
    np.savez_compressed(os.path.join(OUT, "c2h2_631g_sweep.npz"), **out)


if __name__ == "__main__":
    main()
