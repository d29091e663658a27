"""Pin the CPU oracle: against the golden vectors produced by the reference
itself (always), and against the live reference + its raw-equation second
implementation when /root/reference is present (build container only)."""
import numpy as np
import pytest

from helpers import load_golden, MODES
from oracle import synth, ref_loader
from oracle.ccsd_np import OracleGCC, soft_threshold
from oracle import refactored_np as R

TOL = 1e-13


@pytest.mark.parametrize("name", ["ccsd_o4v6.npz", "ccsd_o5v8.npz"])
def test_oracle_matches_golden(name):
    g = load_golden(name)
    o, v = int(g["nocc"]), int(g["nvir"])
    er = synth.SynthEris(o, v)
    t1, t2, l1, l2 = synth.amplitudes(o, v)
    cc = OracleGCC(er)
    for fname, fsp in (("sym", synth.fsp(o, v)), ("ns", g["fsp_ns"])):
        for tag, alpha, eq in MODES:
            a, b = cc.tupdate(t1, t2, fsp=fsp, alpha=alpha, equation=eq)
            assert np.abs(a - g["T1_%s_%s" % (fname, tag)]).max() < TOL
            assert np.abs(b - g["T2_%s_%s" % (fname, tag)]).max() < TOL
            a, b = cc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha, equation=eq)
            assert np.abs(a - g["L1_%s_%s" % (fname, tag)]).max() < TOL
            assert np.abs(b - g["L2_%s_%s" % (fname, tag)]).max() < TOL
        assert abs(cc.energy(t1, t2, fsp) - float(g["E_%s" % fname])) < TOL
    assert np.abs(cc.gamma(t1, t2, l1, l2) - g["gamma"]).max() < TOL
    # identity the reference itself checks (CCSD.py:675-699): factorised == raw equations
    a, b = cc.tupdate(t1, t2, equation=True)
    assert np.abs(a - g["rawT1"]).max() < 1e-12 and np.abs(b - g["rawT2"]).max() < 1e-12
    a, b = cc.lupdate(t1, t2, l1, l2, equation=True)
    assert np.abs(a - g["rawL1"]).max() < 1e-12 and np.abs(b - g["rawL2"]).max() < 1e-12


@pytest.mark.parametrize("name", ["ccsd_o4v6.npz", "ccsd_o5v8.npz"])
def test_refactored_spec_matches_golden(name):
    """The refactored algorithm (what the device executes) equals the reference."""
    g = load_golden(name)
    o, v = int(g["nocc"]), int(g["nvir"])
    er = synth.SynthEris(o, v)
    E = R.DeviceErisSpec(er)
    t1, t2, l1, l2 = synth.amplitudes(o, v)
    for fname, fsp in (("sym", synth.fsp(o, v)), ("ns", g["fsp_ns"])):
        for tag, alpha, eq in MODES:
            a, b = R.tupdate(E, er.fock, t1, t2, fsp=fsp, alpha=alpha, equation=eq)
            assert np.abs(a - g["T1_%s_%s" % (fname, tag)]).max() < TOL
            assert np.abs(b - g["T2_%s_%s" % (fname, tag)]).max() < TOL
            a, b = R.lupdate(E, er.fock, t1, t2, l1, l2, fsp=fsp, alpha=alpha, equation=eq)
            assert np.abs(a - g["L1_%s_%s" % (fname, tag)]).max() < TOL
            assert np.abs(b - g["L2_%s_%s" % (fname, tag)]).max() < TOL
    assert np.abs(R.gamma(t1, t2, l1, l2) - g["gamma"]).max() < TOL


def test_quirks():
    """Q1: v<0 behaves like v==0; Q2: lupdate(alpha=0) != lupdate(alpha=None)."""
    e = np.array([0.5, -0.5, 0.05, 0.5, -0.5, 0.05, 0.5, -0.05])
    v = np.array([1.0, 1.0, 1.0, -1.0, -1.0, -1.0, 0.0, 0.0])
    w = soft_threshold(e, v, 0.1)
    assert np.allclose(w, [0.6, -0.4, 0.15, 0.4, -0.4, 0.0, 0.4, 0.0])
    with pytest.raises(ValueError):
        soft_threshold(np.zeros(3), np.zeros(4), 0.1)
    o, v_ = 4, 6
    er = synth.SynthEris(o, v_)
    t1, t2, l1, l2 = synth.amplitudes(o, v_)
    cc = OracleGCC(er)
    a0 = cc.lupdate(t1, t2, l1, l2, alpha=0.0)
    an = cc.lupdate(t1, t2, l1, l2, alpha=None)
    assert np.abs(a0[1] - an[1]).max() > 1e-6
    b0 = cc.tupdate(t1, t2, alpha=0.0)
    bn = cc.tupdate(t1, t2, alpha=None)
    assert np.abs(b0[1] - bn[1]).max() < 1e-14
    g = cc.gamma(t1, t2, l1, l2)
    assert abs(np.trace(g) - o) < 1e-12 and np.abs(g - g.T).max() < 1e-15


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("ov", [(4, 6), (6, 9)])
def test_oracle_matches_live_reference(ov):
    o, v = ov
    CCSD, U, RAW = ref_loader.load("CCSD", "utilities", "CC_raw_equations")
    er = synth.SynthEris(o, v)
    t1, t2, l1, l2 = synth.amplitudes(o, v)
    fsp = synth.fsp(o, v)
    ref, orc, orf = CCSD.GCC(er), OracleGCC(er), OracleGCC(er, faithful=True)
    for tag, alpha, eq in MODES:
        for cc in (orc, orf):
            a, b = ref.tupdate(t1, t2, fsp=fsp, alpha=alpha, equation=eq)
            c, d = cc.tupdate(t1, t2, fsp=fsp, alpha=alpha, equation=eq)
            assert np.abs(a - c).max() < TOL and np.abs(b - d).max() < TOL
            a, b = ref.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha, equation=eq)
            c, d = cc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha, equation=eq)
            assert np.abs(a - c).max() < TOL and np.abs(b - d).max() < TOL
    assert np.abs(ref.gamma(t1, t2, l1, l2) - orc.gamma(t1, t2, l1, l2)).max() < TOL
    assert abs(ref.energy(t1, t2, fsp) - orc.energy(t1, t2, fsp)) < TOL
    rng = np.random.default_rng(0)
    e, var = rng.standard_normal((7, 5)), rng.standard_normal((7, 5))
    var[0] = 0.0
    for al in (0.0, 1e-3, 0.5):
        assert np.array_equal(U.subdiff(e, var, al), soft_threshold(e, var, al))
