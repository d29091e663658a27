"""CPU check of the device *lowering*: replay the op list the product library
would launch (ecw_plan_dump) with the numpy interpreter in tests/plan_interp.py
and compare with the oracle.  No CUDA call is made."""
import numpy as np
import pytest

from helpers import plan_json, eris_slots, flags_of, MODES, load_golden
from oracle import synth
from oracle.ccsd_np import OracleGCC
from plan_interp import Interp

TOL = 1e-13


@pytest.mark.parametrize("ov", [(2, 3), (4, 6), (5, 7), (6, 11)])
def test_ccsd_plans_match_oracle(built_lib, ov):
    o, v = ov
    er = synth.SynthEris(o, v)
    t1, t2, l1, l2 = synth.amplitudes(o, v)
    fsp = synth.fsp(o, v)
    orc = OracleGCC(er)
    base = eris_slots(er)
    base.update(t1=t1, t2=t2, l1=l1, l2=l2, fsp=fsp, fock=er.fock.copy())
    for tag, alpha, eq in MODES:
        for fn in ("tupdate", "lupdate"):
            pl = plan_json(built_lib, o, v, fn, flags_of(alpha, eq))
            sl = dict(base)
            sl["out1"] = np.full((o, v), np.nan)
            sl["out2"] = np.full((o, o, v, v), np.nan)
            Interp(pl, sl, alpha=alpha or 0.0).run()
            if fn == "tupdate":
                ref = orc.tupdate(t1, t2, fsp=fsp, alpha=alpha, equation=eq)
            else:
                ref = orc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha, equation=eq)
            assert np.abs(sl["out1"] - ref[0]).max() < TOL, (fn, tag)
            assert np.abs(sl["out2"] - ref[1]).max() < TOL, (fn, tag)
    pl = plan_json(built_lib, o, v, "gamma", 0)
    sl = dict(base)
    sl["rdm1"] = np.full((o + v, o + v), np.nan)
    Interp(pl, sl).run()
    assert np.abs(sl["rdm1"] - orc.gamma(t1, t2, l1, l2)).max() < TOL
    pl = plan_json(built_lib, o, v, "energy", 0)
    it = Interp(pl, dict(base)).run()
    assert abs(it.slots["scal"][0] - orc.energy(t1, t2, fsp)) < TOL


def test_ccsd_plan_matches_golden(built_lib):
    g = load_golden("ccsd_o5v8.npz")
    o, v = 5, 8
    er = synth.SynthEris(o, v)
    t1, t2, l1, l2 = synth.amplitudes(o, v)
    base = eris_slots(er)
    base.update(t1=t1, t2=t2, l1=l1, l2=l2, fsp=g["fsp_ns"], fock=er.fock.copy())
    for tag, alpha, eq in MODES:
        for fn, k1, k2 in (("tupdate", "T1", "T2"), ("lupdate", "L1", "L2")):
            sl = dict(base)
            sl["out1"] = np.full((o, v), np.nan)
            sl["out2"] = np.full((o, o, v, v), np.nan)
            Interp(plan_json(built_lib, o, v, fn, flags_of(alpha, eq)), sl, alpha=alpha or 0.0).run()
            assert np.abs(sl["out1"] - g["%s_ns_%s" % (k1, tag)]).max() < TOL
            assert np.abs(sl["out2"] - g["%s_ns_%s" % (k2, tag)]).max() < TOL


def test_north_star_plan_has_no_integral_permutes(built_lib):
    """At (40,400) the lowering must never permute an o v^3 / v^4-sized tensor, must fit one
    B200 and must execute fewer GEMM flops than the dense reference factorisation."""
    o, v = 40, 400
    total = 0.0
    for fn in ("tupdate", "lupdate"):
        pl = plan_json(built_lib, o, v, fn, 0)
        assert pl["workspace_elems"] * 8 < 40e9
        for op in pl["ops"]:
            if op["kind"] == "permute":
                n = int(np.prod(op["c"]["dim"]))
                assert n <= o * o * v * v, op["note"]
        total += pl["gemm_flops"]
    assert total < 9.93e13          # F_alg of SURVEY.md §8(d)
    assert total > 5e13
