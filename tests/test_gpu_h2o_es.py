"""Config 3 of BASELINE.json in its named basis on the GPU: H2O/6-31+G* ((nocc, nvir) = (10, 34), d functions from
ecw_cc_b200.molint) ECW-CCS excited states with transition-dipole potentials — `ecw_cc_b200.Solver_ES` over the CUDA
`Gccs` against the UNMODIFIED reference solver / `CCS.Gccs` / `exp_pot.Exp` (tests/golden/h2o_631pgs_es.npz)."""
import pytest

from helpers import load_golden
from oracle.make_golden_ccs_solvers import run_es
from oracle.make_golden_h2o_es import CASES, water_diffuse
from test_ccs_solvers_cpu import compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("eng", ["int8", "dmma"])
def test_excited_states_in_the_named_basis(built_lib, eng, monkeypatch):
    import ecw_cc_b200 as ecw
    monkeypatch.setenv("ECW_GEMM", eng)
    g = load_golden("h2o_631pgs_es.npz")
    mol, er, _ = water_diffuse((float(g["EHF"]), g["mo_energy"], g["mo_coeff"]))
    out = run_es(ecw.Solver_ES, ecw.Gccs, ecw.exp_pot.Exp, ecw.utilities.koopman_init_guess, mol, er, cases=CASES)
    worst = compare(out, g, "es_")
    print("H2O/6-31+G* ES solver, engine %s: max deviation %.2e" % (eng, worst))
