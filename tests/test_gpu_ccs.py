"""GPU parity tests of the CCS ground/excited-state path (ecw_cc_b200.Gccs -> ecw_op_* kernels)
against the CPU oracle and the reference-generated golden vectors, through the same call list
(`oracle/make_golden.ccs_calls`) that produced the fixtures."""
import types

import numpy as np
import pytest

from helpers import load_golden
from test_oracle_pins_ccs import inputs_from_golden

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _mod(cc):
    return types.SimpleNamespace(gamma_CCS=cc.gamma, gamma_unsym_CCS=cc.gamma_unsym,
                                 gamma_es_CCS=cc.gamma_es, gamma_tr_CCS=cc.gamma_tr)


@pytest.mark.parametrize("name", ["ccs_o4v6.npz", "ccs_o6v9.npz"])
def test_ccs_matches_golden(built_lib, name, engine):
    import ecw_cc_b200 as ecw
    from oracle import synth
    from oracle.make_golden import ccs_calls
    g = load_golden(name)
    o, v = int(g["nocc"]), int(g["nvir"])
    cc = ecw.Gccs(synth.SynthEris(o, v))
    out = ccs_calls(cc, _mod(cc), inputs_from_golden(g))
    assert len(out) > 60
    for k, val in out.items():
        assert np.abs(val - g[k]).max() < TOL, k


@pytest.mark.parametrize("ov", [(3, 5), (7, 12), (10, 34)])
def test_ccs_matches_oracle(built_lib, ov, engine):
    import ecw_cc_b200 as ecw
    from oracle import synth, ccs_np
    from oracle.make_golden import ccs_calls, ccs_inputs
    o, v = ov
    er = synth.SynthEris(o, v)
    d = ccs_inputs(o, v)
    cc = ecw.Gccs(er)
    orc = ccs_np.OracleGccs(er)
    omod = types.SimpleNamespace(gamma_CCS=ccs_np.gamma_CCS, gamma_unsym_CCS=ccs_np.gamma_unsym_CCS,
                                 gamma_es_CCS=ccs_np.gamma_es_CCS, gamma_tr_CCS=ccs_np.gamma_tr_CCS)
    a = ccs_calls(cc, _mod(cc), d)
    b = ccs_calls(orc, omod, d)
    for k in b:
        assert np.abs(a[k] - b[k]).max() < TOL, k


def test_ccs_quirks_and_reference_tuples(built_lib):
    """Q6 in-place shift of the passed intermediates; plain numpy tuples (as the reference would
    produce) are accepted; argument-check errors mirror CCS.py:317-320, 542-545."""
    import ecw_cc_b200 as ecw
    from oracle import synth, ccs_np
    from oracle.make_golden import ccs_inputs
    o, v = 5, 8
    er = synth.SynthEris(o, v)
    d = ccs_inputs(o, v)
    cc, orc = ecw.Gccs(er), ccs_np.OracleGccs(er)
    inter = cc.T1inter(d["ts"], d["fsp"])
    before = inter[0].copy()
    cc.tsupdate(d["ts"], inter)
    assert np.abs(np.diagonal(inter[0] - before) + np.diagonal(er.fock)[o:]).max() < 1e-13
    ref_inter = orc.L1inter(d["ts"], d["fsp"])                      # reference-style numpy tuple
    a = cc.lsupdate(d["ts"], d["ls"], tuple(x.copy() if hasattr(x, "copy") else x for x in ref_inter))
    b = orc.lsupdate(d["ts"], d["ls"], ref_inter)
    assert np.abs(a - b).max() < TOL
    with pytest.raises(ValueError):
        cc.tsupdate(d["ts"], cc.T1inter(d["ts"], d["fsp"]), [d["rs"]], None, [d["vm"]])
    with pytest.raises(ValueError):
        cc.lsupdate(d["ts"], d["ls"], cc.L1inter(d["ts"], d["fsp"]), [d["rs"]], [d["rl"], d["rl"]], [0.1], [0.1],
                    [d["vm"]])
    new = cc.rsupdate(d["rs"], 0.3, cc.R1inter(d["ts"], d["fsp"], d["vm"]), np.array([0.7]))
    assert np.all(new[0::2] == 0.0) and np.any(new[1::2] != 0.0)    # Q9


def test_ccs_gs_loop_tracks_oracle(built_lib):
    """Body of Solver_CCS.SCF (Solver_GS.py:166-204) for a few iterations, with and without L1."""
    import ecw_cc_b200 as ecw
    from oracle import synth, ccs_np
    o, v = 6, 10
    er = synth.SynthEris(o, v)
    fsp = synth.fsp(o, v)
    for alpha in (None, 1e-3):
        objs = [ecw.Gccs(er), ccs_np.OracleGccs(er)]
        st = [[np.zeros((o, v)), np.zeros((o, v))] for _ in objs]
        for it in range(4):
            outs = []
            for cc, s in zip(objs, st):
                ts, ls = s
                ti = cc.T1inter(ts, fsp)
                ts = cc.tsupdate(ts, ti) if alpha is None else cc.tsupdate_L1(ts, ti, alpha)
                li = cc.L1inter(ts, fsp)
                ls = cc.lsupdate(ts, ls, li) if alpha is None else cc.lsupdate_L1(ls, li, alpha)
                s[:] = [ts, ls]
                outs.append((ts, ls, cc.gamma(ts, ls), cc.energy_ccs(ts, fsp)))
            for x, y in zip(outs[0], outs[1]):
                assert np.abs(np.asarray(x) - np.asarray(y)).max() < TOL, (alpha, it)


def test_extract_r0_matches_reference_fixture(built_lib, engine):
    """`Gccs.Extract_r0` (CCS.py:1036-1079) against outputs of the unmodified reference: same values
    (relative 1e-10: the roots divide by a small c), same ValueError cases."""
    import ecw_cc_b200 as ecw
    from oracle import synth
    from oracle.make_golden import ccs_inputs
    from oracle.make_golden_ccs_r0 import SIZES, call, r1_variants
    g = load_golden("ccs_extract_r0.npz")
    for o, v in SIZES:
        cc = ecw.Gccs(synth.SynthEris(o, v))
        d = ccs_inputs(o, v)
        got = [call(cc, r1, d, vm) for vm in (d["vm"], d["vm2"]) for r1 in r1_variants(d)]
        got.append(call(cc, d["rs"], dict(d, fsp=None), d["vm"]))
        want, stat = g["r0_o%dv%d" % (o, v)], g["status_o%dv%d" % (o, v)]
        assert [s for _, s in got] == list(stat)
        for (x, s), w in zip(got, want):
            assert s == 1 or abs(x - w) < TOL * max(1.0, abs(w))
