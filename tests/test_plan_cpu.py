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


@pytest.mark.parametrize("ov", [(3, 4), (4, 6), (5, 9)])
def test_general_plans_with_unsymmetric_amplitudes(built_lib, ov):
    """General path (no ECW_ANTISYM flag): t2/l2 without any permutational symmetry, as the
    reference's dense einsums accept them — the state of the amplitudes after an L1 update."""
    o, v = ov
    er = synth.SynthEris(o, v)
    fsp = synth.fsp(o, v)
    rng = np.random.default_rng(17 * o + v)
    t1, l1 = 0.05 * rng.standard_normal((o, v)), 0.05 * rng.standard_normal((o, v))
    t2, l2 = 0.02 * rng.standard_normal((o, o, v, v)), 0.02 * rng.standard_normal((o, o, v, v))
    orc = OracleGCC(er)
    base = eris_slots(er)
    base.update(t1=t1, t2=t2, l1=l1, l2=l2, fsp=fsp, fock=er.fock.copy())
    for tag, alpha, eq in MODES:
        for fn in ("tupdate", "lupdate"):
            sl = dict(base)
            sl["out1"] = np.full((o, v), np.nan)
            sl["out2"] = np.full((o, o, v, v), np.nan)
            Interp(plan_json(built_lib, o, v, fn, flags_of(alpha, eq, antisym=False)), sl, alpha=alpha or 0.0).run()
            if fn == "tupdate":
                ref = orc.tupdate(t1, t2, fsp=fsp, alpha=alpha, equation=eq)
            else:
                ref = orc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha, equation=eq)
            assert np.abs(sl["out1"] - ref[0]).max() < TOL, (fn, tag)
            assert np.abs(sl["out2"] - ref[1]).max() < TOL, (fn, tag)
    # gamma / energy never assume antisymmetry
    sl = dict(base)
    sl["rdm1"] = np.full((o + v, o + v), np.nan)
    Interp(plan_json(built_lib, o, v, "gamma", 0), sl).run()
    assert np.abs(sl["rdm1"] - orc.gamma(t1, t2, l1, l2)).max() < TOL
    it = Interp(plan_json(built_lib, o, v, "energy", 0), dict(base)).run()
    assert abs(it.slots["scal"][0] - orc.energy(t1, t2, fsp)) < TOL


def test_l1_update_breaks_antisymmetry_and_general_path_tracks_it(built_lib):
    """Q1 consequence: after one L1-regularised tupdate the reference's t2 is no longer
    antisymmetric; the packed path would then be wrong at the 1e-8 level, the general path is exact."""
    o, v = 5, 9
    er = synth.SynthEris(o, v)
    fsp = synth.fsp(o, v)
    orc = OracleGCC(er)
    e = np.diagonal(er.fock)
    d1 = e[:o, None] - e[None, o:]
    d2 = d1[:, None, :, None] + d1[None, :, None, :]
    ts, ls, td = np.zeros((o, v)), np.zeros((o, v)), er.oovv / d2
    ld = td.copy()
    ts, td = orc.tupdate(ts, td, fsp=fsp, alpha=5e-4)
    assert np.abs(td + td.transpose(1, 0, 2, 3)).max() > 1e-6
    ref = orc.lupdate(ts, td, ls, ld, fsp=fsp, alpha=5e-4)
    base = eris_slots(er)
    base.update(t1=ts, t2=td, l1=ls, l2=ld, fsp=fsp, fock=er.fock.copy())
    err = {}
    for anti in (True, False):
        sl = dict(base)
        sl["out1"] = np.full((o, v), np.nan)
        sl["out2"] = np.full((o, o, v, v), np.nan)
        Interp(plan_json(built_lib, o, v, "lupdate", flags_of(5e-4, False, antisym=anti)), sl, alpha=5e-4).run()
        err[anti] = max(np.abs(sl["out1"] - ref[0]).max(), np.abs(sl["out2"] - ref[1]).max())
    assert err[False] < TOL and err[True] > 1e-9


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
        pl = plan_json(built_lib, o, v, fn, 4)
        assert pl["workspace_elems"] * 8 < 40e9
        for op in pl["ops"]:
            if op["kind"] == "permute":
                n = int(np.prod(op["c"]["dim"]))
                assert n <= o * o * v * v, op["note"]
        total += pl["gemm_flops"]
    assert total < 9.93e13          # F_alg of SURVEY.md §8(d)
    assert total > 5e13


# ---------------------------------------------------------------- INT8 tensor-core engine (csrc/ozaki.cu)
from plan_interp import oz_const_slots


def _run_modes(built_lib, o, v, t1, t2, l1, l2, fsp, er, antisym, ns, tol, vvvv_planes=True, ovvv_planes=False,
               splitk_min_k=None, cut_cache_min=None):
    orc = OracleGCC(er)
    base = eris_slots(er)
    base.update(t1=t1, t2=t2, l1=l1, l2=l2, fsp=fsp, fock=er.fock.copy())
    if vvvv_planes:
        base["vvvv_oz"], base["vvvv_ozs"] = oz_const_slots(base["vvvv_p"], ns)
        base["vvvv_p"] = np.full(1, np.nan)          # FP64 vvvv is not bound in this mode
    if ovvv_planes:
        pv = v * (v - 1) // 2
        O = base["ovvv_p"].reshape(o * v, pv)
        base["ovvv_oz1"], base["ovvv_oz1s"] = oz_const_slots(O, ns)
        base["ovvv_oz2"], base["ovvv_oz2s"] = oz_const_slots(np.ascontiguousarray(O.T), ns, K1=o)
        base["ovvv_p"] = np.full(1, np.nan)
    worst = 0.0
    n_oz = 0
    for tag, alpha, eq in MODES:
        for fn in ("tupdate", "lupdate"):
            pl = plan_json(built_lib, o, v, fn, flags_of(alpha, eq, antisym=antisym), int8_digits=ns, min_flops=-1.0,
                           vvvv_planes=vvvv_planes, ovvv_planes=ovvv_planes, splitk_min_k=splitk_min_k,
                           cut_cache_min=cut_cache_min)
            if splitk_min_k:
                assert any(op["kind"] == "oz_gemm" and "split-K" in op["note"] for op in pl["ops"]), fn
            if ovvv_planes:
                assert any(op["kind"] == "oz_gemm" and op["batch"] > 1 for op in pl["ops"])
            n_oz += sum(op["kind"] == "oz_gemm" for op in pl["ops"])
            sl = dict(base)
            sl["out1"] = np.full((o, v), np.nan)
            sl["out2"] = np.full((o, o, v, v), np.nan)
            Interp(pl, sl, alpha=alpha or 0.0).run()
            if fn == "tupdate":
                ref = orc.tupdate(t1, t2, fsp=fsp, alpha=alpha, equation=eq)
            else:
                ref = orc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha, equation=eq)
            worst = max(worst, np.abs(sl["out1"] - ref[0]).max(), np.abs(sl["out2"] - ref[1]).max())
    assert n_oz > 0
    assert worst < tol, worst
    return worst


@pytest.mark.parametrize("ov", [(3, 4), (4, 6), (6, 11)])
def test_int8_engine_plans_match_oracle(built_lib, ov):
    """Every unbatched GEMM forced onto the digit-plane route (6 base-256 digits), packed vvvv bound as planes:
    the replayed plan (exact integer arithmetic in numpy) matches the oracle far below the 1e-10 bar."""
    o, v = ov
    er = synth.SynthEris(o, v)
    t1, t2, l1, l2 = synth.amplitudes(o, v)
    _run_modes(built_lib, o, v, t1, t2, l1, l2, synth.fsp(o, v), er, True, 6, 1e-12)


def test_int8_engine_general_path_and_digit_count(built_lib):
    o, v = 4, 7
    er = synth.SynthEris(o, v)
    rng = np.random.default_rng(5)
    t1, l1 = 0.05 * rng.standard_normal((o, v)), 0.05 * rng.standard_normal((o, v))
    t2, l2 = 0.02 * rng.standard_normal((o, o, v, v)), 0.02 * rng.standard_normal((o, o, v, v))
    e7 = _run_modes(built_lib, o, v, t1, t2, l1, l2, synth.fsp(o, v), er, False, 6, 1e-12)
    e8 = _run_modes(built_lib, o, v, t1, t2, l1, l2, synth.fsp(o, v), er, False, 7, 1e-13, vvvv_planes=False)
    # fewer digits = a coarser product: the truncation error is visible and grows by 2^8 per digit
    o4 = OracleGCC(er)
    base = eris_slots(er)
    base.update(t1=t1, t2=t2, l1=l1, l2=l2, fsp=synth.fsp(o, v), fock=er.fock.copy())
    sl = dict(base)
    sl["out1"], sl["out2"] = np.full((o, v), np.nan), np.full((o, o, v, v), np.nan)
    Interp(plan_json(built_lib, o, v, "tupdate", flags_of(None, True, antisym=False), int8_digits=3), sl).run()
    ref = o4.tupdate(t1, t2, fsp=synth.fsp(o, v), equation=True)
    e4 = np.abs(sl["out2"] - ref[1]).max()
    assert e4 > 100 * max(e7, e8, 1e-16) and e4 < 1e-5


def test_int8_engine_north_star_plan(built_lib):
    """(40,400) with the default threshold: both ladders, the five o^3v^3 rings and R4/R6/R9 run on the
    INT8 pipe (> 90 % of the flops); the vvvv planes replace the FP64 layout; everything fits one B200."""
    o, v = 40, 400
    oz = tot = 0.0
    for fn in ("tupdate", "lupdate"):
        pl = plan_json(built_lib, o, v, fn, 4, int8_digits=6, min_flops=2e10, vvvv_planes=True, ovvv_planes=True)
        assert pl["workspace_elems"] * 8 < 45e9
        oz += pl["oz_flops"]
        tot += pl["gemm_flops"]
        for op in pl["ops"]:
            for k in "abcde":
                assert not (op[k] and op[k]["slot"] in ("vvvv_p", "ovvv_p")), op["note"]
        big = [op for op in pl["ops"] if op["kind"] == "gemm" and 2.0 * op["M"] * op["N"] * op["K"] * op["batch"] / op["splitk"] > 5e11]
        assert not big, [b["note"] for b in big]
    assert oz / tot > 0.9


@pytest.mark.parametrize("antisym", [True, False])
def test_int8_engine_ovvv_planes(built_lib, antisym):
    """nocc, nvir multiples of 8: ovvv_p bound as digit planes in both orientations; R4/R6/R9 read them and the
    ovvv-streaming terms run as batched INT8 products over row / k1 sub-blocks of the constant plane sets."""
    o, v = 8, 16
    er = synth.SynthEris(o, v)
    if antisym:
        t1, t2, l1, l2 = synth.amplitudes(o, v)
    else:
        rng = np.random.default_rng(9)
        t1, l1 = 0.05 * rng.standard_normal((o, v)), 0.05 * rng.standard_normal((o, v))
        t2, l2 = 0.02 * rng.standard_normal((o, o, v, v)), 0.02 * rng.standard_normal((o, o, v, v))
    _run_modes(built_lib, o, v, t1, t2, l1, l2, synth.fsp(o, v), er, antisym, 6, 1e-12, ovvv_planes=True)


def test_int8_engine_split_k(built_lib):
    """Few-tile products with a long contraction index are cut into K chunks (two-level index, per-chunk row sums,
    one product per chunk, fixed-order sum)."""
    o, v = 8, 16
    er = synth.SynthEris(o, v)
    t1, t2, l1, l2 = synth.amplitudes(o, v)
    _run_modes(built_lib, o, v, t1, t2, l1, l2, synth.fsp(o, v), er, True, 6, 1e-12, splitk_min_k=64)


@pytest.mark.parametrize("antisym", [True, False])
def test_int8_engine_cut_cache(built_lib, antisym):
    """Operands that are cut twice keep their digit planes (Plan::oz_cut cache).  The threshold is lowered so that
    every cut of this small case goes through the cache: planes must survive until the last product that was handed
    them, also when their source buffer is released first (the (16,96) regression of round 2)."""
    o, v = 8, 16
    er = synth.SynthEris(o, v)
    if antisym:
        t1, t2, l1, l2 = synth.amplitudes(o, v)
    else:
        rng = np.random.default_rng(10)
        t1, l1 = 0.05 * rng.standard_normal((o, v)), 0.05 * rng.standard_normal((o, v))
        t2, l2 = 0.02 * rng.standard_normal((o, o, v, v)), 0.02 * rng.standard_normal((o, o, v, v))
    for ovvv_planes in (True, False):
        _run_modes(built_lib, o, v, t1, t2, l1, l2, synth.fsp(o, v), er, antisym, 6, 1e-12, ovvv_planes=ovvv_planes,
                   cut_cache_min=1)
