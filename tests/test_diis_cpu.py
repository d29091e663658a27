"""`ecw_cc_b200.diis.DIIS` (host store) against the restated PySCF algorithm in oracle/pyscf_stub, and the oracle loop
with DIIS against DIIS-accelerated runs of the unmodified reference solver (tests/golden/h2o_631g.npz).  PySCF itself is
absent: DIIS parity is pinned to the restatement ("parity unpinned" against PySCF)."""
import numpy as np
import pytest

from helpers import load_golden
from oracle.pyscf_stub.pyscf.lib.diis import DIIS as StubDIIS


def _fixed_point_problem(n, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    A = 0.8 * A / np.abs(np.linalg.eigvals(A)).max()
    b = rng.standard_normal(n)
    return A, b, np.linalg.solve(np.eye(n) - A, b)


@pytest.mark.parametrize("space,min_space", [(4, 2), (6, 1), (15, 2)])
def test_host_diis_equals_restated_pyscf(space, min_space):
    from ecw_cc_b200.diis import DIIS
    A, b, sol = _fixed_point_problem(40, space)
    mine, ref = DIIS(), StubDIIS()
    mine.space = ref.space = space
    mine.min_space = ref.min_space = min_space
    x = y = np.zeros((8, 5))
    plain = np.zeros(40)
    for it in range(40):
        x = mine.update((A @ x.ravel() + b).reshape(8, 5))
        y = ref.update((A @ y.ravel() + b).reshape(8, 5))
        plain = A @ plain + b
        assert x.shape == (8, 5) and np.abs(x - y).max() < 1e-9 * max(1.0, np.abs(y).max()), it
    assert np.abs(x.ravel() - sol).max() < 0.1 * np.abs(plain - sol).max()           # it does accelerate


def test_oracle_loop_with_diis_reproduces_reference_runs():
    from ecw_cc_b200 import molint
    from oracle.ccsd_np import OracleGCC
    from oracle.make_golden_h2o import DIIS_CASES, H2O
    from oracle.make_golden_solver import target_rdm1
    from oracle.solver_np import ExpMat, scf_loop
    g = load_golden("h2o_631g.npz")
    mol = molint.Molecule(H2O, "6-31g")
    er = molint.geris(mol, (float(g["EHF"]), g["mo_energy"], g["mo_coeff"], molint.integrals(mol)))
    # DIIS reaches the fixed point of the plain iteration (same energies to the convergence threshold) in fewer steps
    assert abs(g["L0_dtl_Ep"][-1] - g["L0_Ep"][-1]) < 1e-9 and len(g["L0_dtl_Ep"]) < 0.6 * len(g["L0_Ep"])
    assert abs(g["L05_dtl_Ep"][-1] - g["L05_Ep"][-1]) < 1e-9 and np.abs(g["L05_dtl_rdm1"] - g["L05_rdm1"]).max() < 1e-8
    for tag, L, alpha, maxiter, diis, maxdiis in (DIIS_CASES[0], DIIS_CASES[3]):
        out = scf_loop(OracleGCC(er), ExpMat(L, target_rdm1(10, 16)), L, alpha=alpha, conv_thres=float(g["conv_thres"]),
                       maxiter=maxiter, diis=diis, maxdiis=maxdiis)
        assert out[0] == str(g[tag + "_text"])
        assert np.abs(out[1] - g[tag + "_Ep"]).max() < 1e-11 and np.abs(out[4] - g[tag + "_rdm1"]).max() < 1e-10
