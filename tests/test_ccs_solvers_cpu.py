"""The mirrored ECW-CCS solver loops (`ecw_cc_b200.Solver_CCS`, `ecw_cc_b200.Solver_ES`, with `ecw_cc_b200.exp_pot.Exp`
and the host DIIS) driving the numpy ORACLE `Gccs`, against runs of the UNMODIFIED reference solvers with the
reference `Gccs` / `Exp` on H2O/6-31G (tests/golden/ccs_solvers_h2o.npz, oracle/make_golden_ccs_solvers.py).
This pins the loop logic on the CPU; tests/test_gpu_ccs_solvers.py runs the same loops over the CUDA `Gccs`."""
import numpy as np
import pytest

from helpers import load_golden
from oracle.ccs_np import OracleGccs
from oracle.make_golden_ccs_solvers import run_es, run_gs, water

TOL = 1e-10


def compare(out, g, prefix, tol=TOL):
    keys = [k for k in g if k.startswith(prefix)]
    assert keys and sorted(keys) == sorted(k for k in out if k.startswith(prefix))
    worst = 0.0
    for k in keys:
        if k.endswith("_text"):
            assert str(out[k]) == str(g[k]), k
            continue
        want, got = np.asarray(g[k], dtype=float), np.asarray(out[k], dtype=float)
        assert want.shape == got.shape, k
        worst = max(worst, np.abs(want - got).max())
        assert np.abs(want - got).max() < tol, k
    return worst


@pytest.fixture(scope="module")
def h2o():
    return water()


def test_ground_state_solver_loop(h2o):
    import ecw_cc_b200 as ecw
    mol, er = h2o
    g = load_golden("ccs_solvers_h2o.npz")
    compare(run_gs(ecw.Solver_CCS, OracleGccs, ecw.exp_pot.Exp, er), g, "gs_")
    assert "Convergence reached" in str(g["gs_L05_text"]) and "Convergence reached" in str(g["gs_L2_tl_text"])
    assert len(g["gs_L2_tl_Ep"]) < 30


def test_excited_state_solver_loop(h2o, capsys):
    import ecw_cc_b200 as ecw
    mol, er = h2o
    g = load_golden("ccs_solvers_h2o.npz")
    compare(run_es(ecw.Solver_ES, OracleGccs, ecw.exp_pot.Exp, ecw.utilities.koopman_init_guess, mol, er), g, "es_")
    # CIS-like excitation energies of water at L = 0 (right = left), shifted by the transition-dipole potentials
    assert abs(g["es_trdip_0_Ep"][1, 0] - 0.32915335607814) < 1e-10 and g["es_trdip_0_Ep"][1, 0] == g["es_trdip_0_Ep"][1, 1]
    assert abs(g["es_trdip_1_Ep"][2, 0] - g["es_trdip_0_Ep"][2, 0]) > 1e-2


def test_solver_es_surface(h2o):
    import ecw_cc_b200 as ecw
    mol, er = h2o
    vx = ecw.exp_pot.Exp(0.0, [[], [['trdip', [0.5, 0., 0.]]]], mol, er.mo_coeff_g)
    cc = OracleGccs(er)
    with pytest.raises(ValueError):
        ecw.Solver_ES(cc, vx, rn_ini=[np.zeros((10, 16))] * 2)
    with pytest.raises(ValueError):
        ecw.Solver_ES(cc, vx, val_core=[1, 0], conv_var="x")
    s = ecw.Solver_ES(cc, vx, val_core=[1, 0], maxiter=1)          # Koopmans route (a TypeError in the reference)
    assert np.argwhere(s.rn_ini[0]).tolist() == [[9, 1]]
    with pytest.raises(NotImplementedError):
        s.SCF(L=0.0, diis='ES', print_ite=False)
    text, amp, delta, ep, rdm1 = s.SCF(L=0.0, print_ite=True)
    assert text == 'Max iteration reached' and abs(np.trace(rdm1) - 10) < 1e-12 and set(amp) == {'ts', 'ls', 'rn', 'ln', 'r0n', 'l0n'}
