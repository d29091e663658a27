"""Golden vectors of CONVERGED L1-ECW-CCSD ground states: the UNMODIFIED reference solver
(Solver_GS.Solver_CCSD.SCF, Solver_GS.py:621-742) driving the UNMODIFIED reference CCSD.GCC with the reference's own
exp_pot.Exp ('mat' target, exp_pot.py:185-214) on the synthetic integrals of oracle/synth.py.  Build container only:

    python -m oracle.make_golden_solver

tests/golden/solver_ccsd_*.npz hold, per (L, alpha) case: Ep / Delta / conv histories, the final rdm1 and the final
amplitudes — what "the converged energies and rdm1 must also match" (BASELINE.json north_star) is checked against.
The target rdm1 is function defined (target_rdm1 below), so the fixtures carry outputs only.
"""
import os

import numpy as np

from . import ref_loader, synth

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
# (tag, L, alpha, maxiter): the unregularised cases converge to 1e-9 in 5-8 iterations; with the L1 term the reference's
# quasi-Newton iteration creeps (conv ~ 1/it), so those cases are cut at 12 iterations ("Max iteration reached")
CASES = [("L0", 0.0, None, 60), ("L05", 0.05, None, 60), ("L05_a", 0.05, 2e-4, 12), ("L2_a", 0.2, 5e-4, 12)]


def target_rdm1(o, v):
    """'Experimental' rdm1 of the fit: the HF rdm1 plus a fixed symmetric, trace-free perturbation."""
    n = o + v
    g = np.diag(np.concatenate([np.ones(o), np.zeros(v)]))
    rng = np.random.default_rng(4242 + 31 * o + v)
    p = 0.02 * rng.standard_normal((n, n))
    p = 0.5 * (p + p.T)
    p -= np.eye(n) * np.trace(p) / n
    return g + p


def solver_case(o, v, conv_thres=1e-9):
    CCSD, Solver_GS, exp_pot = ref_loader.load("CCSD", "Solver_GS", "exp_pot")
    er = synth.SynthEris(o, v)
    out = {"nocc": o, "nvir": v, "conv_thres": conv_thres}
    for tag, L, alpha, maxiter in CASES:
        mycc = CCSD.GCC(er)
        vx = exp_pot.Exp(L, [[["mat", target_rdm1(o, v)]]], None, None)
        solver = Solver_GS.Solver_CCSD(mycc, vx, conv="tl", conv_thres=conv_thres, maxiter=maxiter)
        text, ep, delta, conv, rdm1, amps = solver.SCF(L, alpha=alpha)
        print("(%d,%d) %s: %s | Ep %.12f | Delta %.6f" % (o, v, tag, text, ep[-1], delta[-1][0]))
        out[tag + "_text"] = np.array(text)
        out[tag + "_Ep"], out[tag + "_Delta"], out[tag + "_conv"], out[tag + "_rdm1"] = ep, delta, conv, rdm1
        for k, a in zip(("ts", "ls", "td", "ld"), amps):
            out[tag + "_" + k] = a
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for (o, v) in [(4, 6), (8, 16)]:
        np.savez_compressed(os.path.join(OUT, "solver_ccsd_o%dv%d.npz" % (o, v)), **solver_case(o, v))


if __name__ == "__main__":
    main()
