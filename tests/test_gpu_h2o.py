"""Config 1 of BASELINE.json on the GPU: H2O/6-31G (integrals from ecw_cc_b200.molint, no PySCF) L1-ECW-CCSD ground
state fitted to a target rdm1, `ecw_cc_b200.Solver_CCSD` + `ecw_cc_b200.exp_pot.Exp` against the UNMODIFIED reference
solver / CCSD.GCC / exp_pot.Exp on the same integrals (tests/golden/h2o_631g.npz, oracle/make_golden_h2o.py):
converged energies, rdm1 and amplitudes to 1e-10, same iteration counts — under every GEMM engine."""
import numpy as np
import pytest

from helpers import load_golden
from oracle.make_golden_h2o import CASES, H2O
from oracle.make_golden_solver import target_rdm1

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_water_ground_states_match_reference(built_lib, engine):
    import ecw_cc_b200 as ecw
    from ecw_cc_b200 import molint
    g = load_golden("h2o_631g.npz")
    mol = molint.Molecule(H2O, "6-31g")
    er = molint.geris(mol, (float(g["EHF"]), g["mo_energy"], g["mo_coeff"], molint.integrals(mol)))
    o, v = er.nocc, er.fock.shape[0] - er.nocc
    assert (o, v) == (10, 16)
    for tag, L, alpha, maxiter in CASES:
        mycc = ecw.GCC(er)
        vx = ecw.exp_pot.Exp(L, [[["mat", target_rdm1(o, v)]]], None, None)
        text, ep, delta, conv, rdm1, amps = ecw.Solver_CCSD(mycc, vx, conv="tl", conv_thres=float(g["conv_thres"]),
                                                            maxiter=maxiter).SCF(L, alpha=alpha)
        assert text == str(g[tag + "_text"]), tag
        assert np.abs(ep - g[tag + "_Ep"]).max() < TOL, tag
        assert np.abs(delta - g[tag + "_Delta"]).max() < TOL, tag
        assert np.abs(rdm1 - g[tag + "_rdm1"]).max() < TOL, tag
        for k, a in zip(("ts", "ls", "td", "ld"), amps):
            assert np.abs(a - g[tag + "_" + k]).max() < TOL, (tag, k)
    # total energy of the unconstrained CCSD ground state
    assert abs(float(g["EHF"]) + float(g["L0_Ep"][-1]) - (-76.1193463836)) < 1e-8
