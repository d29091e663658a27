"""Config 3 of BASELINE.json in its named basis (H2O/6-31+G*, (nocc, nvir) = (10, 34)): the mirrored `Solver_ES` +
`exp_pot.Exp` over the numpy oracle `Gccs` against the UNMODIFIED reference solver / `CCS.Gccs` / `exp_pot.Exp`
(tests/golden/h2o_631pgs_es.npz, oracle/make_golden_h2o_es.py).  tests/test_gpu_h2o_es.py: the same over the CUDA `Gccs`."""
import numpy as np

from helpers import load_golden
from oracle.ccs_np import OracleGccs
from oracle.make_golden_ccs_solvers import run_es
from oracle.make_golden_h2o_es import CASES, water_diffuse
from test_ccs_solvers_cpu import compare


def test_excited_states_in_the_named_basis():
    import ecw_cc_b200 as ecw
    g = load_golden("h2o_631pgs_es.npz")
    mol, er, _ = water_diffuse((float(g["EHF"]), g["mo_energy"], g["mo_coeff"]))
    assert (er.nocc, er.fock.shape[0] - er.nocc) == (10, 34)
    out = run_es(ecw.Solver_ES, OracleGccs, ecw.exp_pot.Exp, ecw.utilities.koopman_init_guess, mol, er, cases=CASES)
    compare(out, g, "es_")
    assert abs(g["es_trdip_0_Ep"][1, 0] - 0.328497032335) < 1e-9      # first excitation energy at L = 0 (Eh)
