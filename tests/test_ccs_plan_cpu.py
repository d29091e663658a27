"""CPU check of the lowering of the ECW-CCS intermediate plans (csrc/ccs_plan.cpp behind ecw_ccs_t1inter / l1inter /
r1inter / esl1inter): the op list the library would launch is replayed with the numpy interpreter and compared with the
oracle (oracle/ccs_np.py, pinned to the unmodified reference CCS.Gccs).  No CUDA call is made."""
import numpy as np
import pytest

from helpers import plan_json, eris_slots
from oracle import synth
from oracle.ccs_np import OracleGccs
from plan_interp import Interp

TOL = 1e-13


def _run(lib, o, v, func, flags, ts, fsp, vm, er, **kw):
    n = o + v
    sl = eris_slots(er)
    sl.update(t1=ts.copy(), fsp=fsp.copy(), fock=(vm.copy() if vm is not None else np.full((n, n), np.nan)),
              rdm1=np.full((n, n), np.nan), out2=np.full(o * o * v * v, np.nan), out1=np.full((o, v), np.nan),
              scal=np.full(16, np.nan))
    it = Interp(plan_json(lib, o, v, func, flags, **kw), sl).run()
    F = it.slots["rdm1"]
    return (F[:o, :o], F[:o, o:], F[o:, :o], F[o:, o:], it.slots["out2"].reshape(v, o, o, v), it.slots["out1"],
            it.slots["scal"][0])


@pytest.mark.parametrize("ov", [(3, 4), (4, 6), (5, 9), (8, 16)])
@pytest.mark.parametrize("int8", [0, 6])
def test_ccs_intermediate_plans_match_oracle(built_lib, ov, int8):
    o, v = ov
    n = o + v
    er = synth.SynthEris(o, v)
    rng = np.random.default_rng(100 * o + v)
    ts = 0.05 * rng.standard_normal((o, v))
    fsp = er.fock + 0.02 * rng.standard_normal((n, n))
    vm = 0.02 * rng.standard_normal((n, n))
    orc = OracleGccs(er)
    kw = dict(int8_digits=int8, min_flops=-1.0) if int8 else {}
    tol = 1e-12 if int8 else TOL

    def close(a, b):
        assert np.abs(np.asarray(a) - np.asarray(b)).max() < tol

    Fab, Fji, Fai = orc.T1inter(ts, fsp)
    oo, ov_, vo, vv, W, X, e = _run(built_lib, o, v, "ccs_t1inter", 0, ts, fsp, None, er, **kw)
    close(vv, Fab); close(oo, Fji); close(vo, Fai)
    assert np.abs(ov_).max() == 0.0

    for e_term in (True, False):
        Fia, Fba, Fij, Wr, E = orc.L1inter(ts, fsp, E_term=e_term)
        oo, ov_, vo, vv, W, X, e = _run(built_lib, o, v, "ccs_l1inter", int(e_term), ts, fsp, None, er, **kw)
        close(ov_, Fia); close(vv, Fba); close(oo, Fij); close(W, Wr)
        if e_term:
            assert abs(e - E) < tol
        assert np.abs(vo).max() == 0.0

    for pot in (vm, None):
        Fab, Fji, Wr, Er, Tia, Pia = orc.R1inter(ts, fsp, pot)
        oo, ov_, vo, vv, W, X, e = _run(built_lib, o, v, "ccs_r1inter", int(pot is not None), ts, fsp, pot, er, **kw)
        close(vv, Fab); close(oo, Fji); close(W, Wr); close(ov_, Tia); close(X, Pia)
        assert abs(e - Er) < tol
        Fba, Fij, Wr, El, Zia, Pl = orc.es_L1inter(ts, fsp, pot)
        oo, ov_, vo, vv, W, X, e = _run(built_lib, o, v, "ccs_esl1inter", int(pot is not None), ts, fsp, pot, er, **kw)
        close(vv, Fba); close(oo, Fij); close(W, Wr); close(ov_, Zia); close(X, Pl)
        assert abs(e - El) < tol
