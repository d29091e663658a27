"""Single L1-regularised update steps of config 2 (C2H2/6-31G) recorded from a run of the UNMODIFIED reference:
the L1 sweep of oracle/make_golden_c2h2.py is replayed with a recording subclass of the reference `CCSD.GCC`
(no method is changed: the subclass stores the arguments and the return value of chosen `tupdate` / `lupdate` calls).

Why: `utilities.subdiff` (utilities.py:53-67, Q1) branches on the sign of the INPUT amplitude.  Over a whole sweep
symmetry-forbidden amplitudes are +-1e-17 rounding noise of numpy's summation order, a different implementation takes
the other branch for single elements and the trajectories separate by O(alpha) (tests/test_gpu_c2h2.py keeps loose
bounds for the sweep).  ONE step from bit-identical inputs has no such freedom — the branch is decided by the recorded
input — so the general (non-antisymmetric, Q11) tupdate / lupdate plans and the soft-threshold finish are held to 1e-10
on a real molecule.  Build container only:   python -m oracle.make_golden_c2h2_step
"""
import os

import numpy as np

from . import ref_loader
from .make_golden_c2h2 import LARRAY, OUT, acetylene, sweep

ALPHA = 2e-4
RECORD = (1, 103)            # call indices (0-based, over the whole sweep: 4 L values x 26 iterations) to record


def main():
    CCSD, Solver_GS, exp_pot = ref_loader.load("CCSD", "Solver_GS", "exp_pot")
    out = {"alpha": ALPHA, "record": np.array(RECORD)}

    class Recording(CCSD.GCC):
        ncall_t = 0
        ncall_l = 0

        def tupdate(self, t1, t2, fsp=None, alpha=None, equation=False):
            r = CCSD.GCC.tupdate(self, t1, t2, fsp=fsp, alpha=alpha, equation=equation)
            k = Recording.ncall_t
            Recording.ncall_t += 1
            if k in RECORD:
                for n, a in zip(("t1", "t2", "fsp", "t1new", "t2new"), (t1, t2, fsp, r[0], r[1])):
                    out["t%d_%s" % (k, n)] = np.array(a)
            return r

        def lupdate(self, t1, t2, l1, l2, fsp=None, alpha=None, equation=False):
            r = CCSD.GCC.lupdate(self, t1, t2, l1, l2, fsp=fsp, alpha=alpha, equation=equation)
            k = Recording.ncall_l
            Recording.ncall_l += 1
            if k in RECORD:
                # (t1, t2) of this call are the (t1new, t2new) of tupdate call k (Solver_GS.py:701-705): stored once
                assert np.array_equal(t2, out["t%d_t2new" % k]) and np.array_equal(fsp, out["t%d_fsp" % k])
                for n, a in zip(("l1", "l2", "l1new", "l2new"), (l1, l2, r[0], r[1])):
                    out["l%d_%s" % (k, n)] = np.array(a)
            return r

    mol, er, scf = acetylene()
    out.update(EHF=scf[0], mo_energy=scf[1], mo_coeff=scf[2])
    sweep(Solver_GS.Solver_CCSD, Recording, exp_pot.Exp, er, ALPHA)
    for k in RECORD:
        t2 = out["t%d_t2new" % k]
        print("call %d: antisymmetry defect of t2 %.2e, |t2|max %.3f" % (
            k, np.abs(t2 + t2.transpose(1, 0, 2, 3)).max(), np.abs(t2).max()))
    np.savez_compressed(os.path.join(OUT, "c2h2_631g_l1_steps.npz"), **out)


if __name__ == "__main__":
    main()
