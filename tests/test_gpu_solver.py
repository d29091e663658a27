"""Converged L1-ECW-CCSD ground states (BASELINE.json north_star: "the converged energies and rdm1 must also match"):
the device-resident solver mirror `ecw_cc_b200.Solver_CCSD` (amplitudes stay on the GPU; per iteration only the rdm1 and
the dressed Fock cross PCIe) against runs of the UNMODIFIED reference solver + CCSD.GCC + exp_pot.Exp
(tests/golden/solver_ccsd_*.npz, oracle/make_golden_solver.py), under every GEMM engine."""
import numpy as np
import pytest

from helpers import load_golden
from oracle import synth
from oracle.make_golden_solver import CASES, target_rdm1
from oracle.solver_np import ExpMat

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.mark.parametrize("name", ["solver_ccsd_o4v6.npz", "solver_ccsd_o8v16.npz"])
def test_solver_matches_reference_runs(built_lib, name, engine):
    import ecw_cc_b200 as ecw
    g = load_golden(name)
    o, v = int(g["nocc"]), int(g["nvir"])
    er = synth.SynthEris(o, v)
    for tag, L, alpha, maxiter in CASES:
        mycc = ecw.GCC(er)
        solver = ecw.Solver_CCSD(mycc, ExpMat(L, target_rdm1(o, v)), conv="tl", conv_thres=float(g["conv_thres"]),
                                 maxiter=maxiter)
        text, ep, delta, conv, rdm1, amps = solver.SCF(L, alpha=alpha)
        assert text == str(g[tag + "_text"]), tag
        assert ep.shape == g[tag + "_Ep"].shape and np.abs(ep - g[tag + "_Ep"]).max() < TOL, tag
        assert np.abs(delta - g[tag + "_Delta"]).max() < TOL, tag
        assert np.allclose(conv, g[tag + "_conv"], rtol=1e-4, atol=1e-11), tag
        assert np.abs(rdm1 - g[tag + "_rdm1"]).max() < TOL, tag
        for k, a in zip(("ts", "ls", "td", "ld"), amps):
            assert np.abs(a - g[tag + "_" + k]).max() < TOL, (tag, k)


def test_solver_numpy_api_and_device_api_agree(built_lib):
    """The reference loop written against the numpy API of GCC (what the unchanged reference solver does) and the
    device-resident mirror walk through the same iterates; other convergence measures and start vectors work."""
    import ecw_cc_b200 as ecw
    from oracle.solver_np import scf_loop
    o, v = 6, 10
    er = synth.SynthEris(o, v)
    L, alpha = 0.05, None
    ref = scf_loop(ecw.GCC(er), ExpMat(L, target_rdm1(o, v)), L, alpha=alpha, conv_thres=1e-8, maxiter=30)
    for conv in ("tl", "l", "Ep"):
        r2 = scf_loop(ecw.GCC(er), ExpMat(L, target_rdm1(o, v)), L, alpha=alpha, conv_thres=1e-8, maxiter=30, conv=conv)
        s = ecw.Solver_CCSD(ecw.GCC(er), ExpMat(L, target_rdm1(o, v)), conv=conv, conv_thres=1e-8, maxiter=30)
        out = s.SCF(L, alpha=alpha)
        assert out[0] == r2[0] and np.abs(out[1] - r2[1]).max() < 1e-12
        assert np.allclose(out[3], r2[3], rtol=1e-6, atol=1e-13)
        for a, b in zip(out[5], r2[5]):
            assert np.abs(a - b).max() < 1e-12
    # restart from converged amplitudes: one more pass stays converged
    s = ecw.Solver_CCSD(ecw.GCC(er), ExpMat(L, target_rdm1(o, v)), conv_thres=1e-8, maxiter=30)
    out = s.SCF(L, ref[5][0], ref[5][1], ref[5][2], ref[5][3])
    assert abs(out[1][-1] - ref[1][-1]) < 1e-9
    # Q5 (Solver_GS.py:648-649): the constructor's diis is only used when SCF gets diis=None; the numpy-API loop with
    # the restated pyscf DIIS and the device-resident DIIS walk through the same iterates
    s = ecw.Solver_CCSD(ecw.GCC(er), ExpMat(L, target_rdm1(o, v)), conv_thres=1e-8, maxiter=30, diis='tl', maxdiis=5)
    plain = s.SCF(L, alpha=alpha)
    assert plain[0] == ref[0] and np.abs(plain[1] - ref[1]).max() < 1e-12
    acc = s.SCF(L, alpha=alpha, diis=None)
    r3 = scf_loop(ecw.GCC(er), ExpMat(L, target_rdm1(o, v)), L, alpha=alpha, conv_thres=1e-8, maxiter=30, diis='tl', maxdiis=5)
    assert acc[0] == r3[0] and np.abs(acc[1] - r3[1]).max() < 1e-11 and abs(acc[1][-1] - ref[1][-1]) < 1e-8


def test_density_matrix_potential_on_the_device(built_lib):
    """`ecw_vexp_mat`: potential, dressed Fock, Delta and vmax of the 'mat' target formed on the device —
    against the host route of the same class, bit for bit for the matrices; and the solver takes that route."""
    import torch
    import ecw_cc_b200 as ecw
    o, v = 6, 10
    n = o + v
    er = synth.SynthEris(o, v)
    cc = ecw.GCC(er)
    rng = np.random.default_rng(11)
    rdm1 = np.diag(np.concatenate([np.ones(o), np.zeros(v)])) + 0.01 * rng.standard_normal((n, n))
    hf = np.diag(np.concatenate([np.ones(o), np.zeros(v)]))
    for hf_prop in (False, [[hf]]):
        host = ecw.exp_pot.Exp(0.3, [[["mat", target_rdm1(o, v)]]], None, None, HF_prop=hf_prop)
        devp = ecw.exp_pot.Exp(0.3, [[["mat", target_rdm1(o, v)]]], None, None, HF_prop=hf_prop)
        assert devp.device_mat_ready()
        d_h, vmax_h = host.Vexp_update(rdm1, rdm1, (0, 0), L=0.7)
        d_d, vmax_d, fsp = devp.mat_update_device(torch.from_numpy(rdm1).cuda(), cc.eris.fock_dev, L=0.7)
        assert abs(d_h - d_d) < 1e-14 and vmax_h == vmax_d
        assert np.array_equal(fsp.cpu().numpy(), np.subtract(cc.fock, host.Vexp[0, 0]))
        devp.sync_host()
        assert isinstance(devp.Vexp[0, 0], np.ndarray) and np.array_equal(devp.Vexp[0, 0], host.Vexp[0, 0])
    assert not ecw.exp_pot.Exp(0.3, [[["mat", hf], ["mat", hf]]], None, None).device_mat_ready()
    # the solver: device route (our Exp) and host route (any other object) walk through the same iterates
    L = 0.05
    a = ecw.Solver_CCSD(ecw.GCC(er), ecw.exp_pot.Exp(L, [[["mat", target_rdm1(o, v)]]], None, None), conv_thres=1e-8,
                        maxiter=30).SCF(L)
    b = ecw.Solver_CCSD(ecw.GCC(er), ExpMat(L, target_rdm1(o, v)), conv_thres=1e-8, maxiter=30).SCF(L)
    assert a[0] == b[0] and np.abs(a[1] - b[1]).max() < 1e-13 and np.abs(a[2] - b[2]).max() < 1e-13
    assert isinstance(a[4], np.ndarray) and np.abs(a[4] - b[4]).max() < 1e-13
