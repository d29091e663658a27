"""Config 2 of BASELINE.json in the 6-31G basis on the GPU: C2H2 L1-ECW-CCSD ground state over a sweep of the Vexp
weight L, driven like `Main.ECW.CCSD_GS` (Main.py:730-763: one solver, previous amplitudes as the next start — here
they stay on the device between L values), against the UNMODIFIED reference solver / CCSD.GCC / exp_pot.Exp with HF
reference values on the same integrals (tests/golden/c2h2_631g_sweep.npz, oracle/make_golden_c2h2.py).
(o, v) = (14, 30); 4 L values x 26 iterations, with and without the L1 term."""
import numpy as np
import pytest

from helpers import load_golden
from oracle.make_golden_c2h2 import LARRAY, SWEEPS, acetylene, pack, sweep

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def c2h2():
    g = load_golden("c2h2_631g_sweep.npz")
    mol, er, _ = acetylene((float(g["EHF"]), g["mo_energy"], g["mo_coeff"]))
    return g, er


@pytest.mark.parametrize("tag,alpha", SWEEPS)
def test_weight_sweep_matches_reference(built_lib, engine, c2h2, tag, alpha):
    import ecw_cc_b200 as ecw
    g, er = c2h2
    assert (er.nocc, er.fock.shape[0] - er.nocc) == (14, 30)
    out = {}
    pack(sweep(ecw.Solver_CCSD, ecw.GCC, ecw.exp_pot.Exp, er, alpha, device=True), tag, out)
    # deviation per kind of output; texts must agree exactly
    worst = {}
    for k, want in ((k, g[k]) for k in g if k.startswith(tag + "_")):
        if k.endswith("_text"):
            assert str(out[k]) == str(want), k
            continue
        kind = "amplitudes" if "_final_" in k else k.rsplit("_", 1)[1]
        dev = np.abs(np.asarray(out[k], dtype=float) - np.asarray(want, dtype=float)).max()
        worst[kind] = max(worst.get(kind, 0.0), dev)
    print("C2H2 sweep %s, engine %s: max deviation %s" % (tag, engine, {k: "%.2e" % v for k, v in worst.items()}))
    if alpha is None:
        assert max(worst.values()) < TOL, worst
    else:
        # Q1 makes the L1 update discontinuous at v = 0: `v > 0 -> e + alpha`, else the soft threshold.  Acetylene's
        # symmetry-forbidden amplitudes are +-1e-17 rounding noise in numpy's einsum and (different) noise or exact
        # zeros here, so single elements take the other branch: those amplitudes then differ by O(alpha / denominator),
        # the observables by products of two such elements.  (Water's forbidden elements are exact zeros on both
        # sides: tests/test_gpu_h2o.py holds 1e-10 with the L1 term.)
        # Measured (identical under the DMMA and INT8 engines, i.e. a branch difference, not arithmetic noise; with the
        # previous summation order of the integral code the same sweep agreed to 6e-13): Ep 9.5e-8, Delta 2.4e-6,
        # conv 7.9e-7, rdm1 8.5e-6, amplitudes 1.2e-4.
        assert worst["Ep"] < 1e-6 and worst["Delta"] < 2e-5 and worst["conv"] < 1e-5 and worst["rdm1"] < 5e-5, worst
        assert worst["amplitudes"] < 10 * alpha, worst
    assert len(LARRAY) == 4 and g[tag + "_L3_Delta"][-1][0] < g[tag + "_L0_Delta"][-1][0]      # the fit tightens with L
