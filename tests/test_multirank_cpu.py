"""N>1 path on CPU: two processes (gloo), each replays ITS rank's plan (vvvv_p row shard, heavy
contractions owner-computes, all-gathers through torch.distributed) with the numpy plan
interpreter; every rank must reproduce the unsharded oracle result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import plan_json, eris_slots, flags_of
from oracle import synth
from oracle.ccsd_np import OracleGCC
from plan_interp import Interp, oz_const_slots

TOL = 1e-13


def _worker(rank, world, port, o, v, antisym, q, int8=0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ecw_cc_b200
        lib = ecw_cc_b200.lib
        er = synth.SynthEris(o, v)
        fsp = synth.fsp(o, v)
        if antisym:
            t1, t2, l1, l2 = synth.amplitudes(o, v)
        else:
            rng = np.random.default_rng(5)
            t1, l1 = 0.05 * rng.standard_normal((o, v)), 0.05 * rng.standard_normal((o, v))
            t2, l2 = 0.02 * rng.standard_normal((o, o, v, v)), 0.02 * rng.standard_normal((o, o, v, v))
        base = eris_slots(er)
        pv = v * (v - 1) // 2
        nshmax = (pv + world - 1) // world
        n0, n1 = min(pv, rank * nshmax), min(pv, (rank + 1) * nshmax)
        base["vvvv_p"] = np.ascontiguousarray(base["vvvv_p"][n0:n1])        # this rank's rows only
        base.update(t1=t1, t2=t2, l1=l1, l2=l2, fsp=fsp, fock=er.fock.copy())
        if int8:      # INT8 engine: the rank's shard is bound as digit planes, every unbatched GEMM on that route
            base["vvvv_oz"], base["vvvv_ozs"] = oz_const_slots(base["vvvv_p"], int8)
            base["vvvv_p"] = np.full(1, np.nan)
            if o % 8 == 0 and v % 8 == 0:          # ovvv_p as digit planes too (R4/R6/R9 read row ranges of them)
                O = base["ovvv_p"].reshape(o * v, pv)
                base["ovvv_oz1"], base["ovvv_oz1s"] = oz_const_slots(O, int8)
                base["ovvv_oz2"], base["ovvv_oz2s"] = oz_const_slots(np.ascontiguousarray(O.T), int8, K1=o)
                base["ovvv_p"] = np.full(1, np.nan)

        def allgather(send, recv):
            out = torch.from_numpy(recv)
            dist.all_gather_into_tensor(out, torch.from_numpy(np.ascontiguousarray(send)))

        orc = OracleGCC(er)
        worst = 0.0
        ncoll = 0
        for alpha, eq in ((None, False), (1e-3, True)):
            for fn in ("tupdate", "lupdate"):
                pl = plan_json(lib, o, v, fn, flags_of(alpha, eq, antisym), rank=rank, world=world, int8_digits=int8,
                               vvvv_planes=bool(int8), ovvv_planes=bool(int8) and o % 8 == 0 and v % 8 == 0)
                ncoll += sum(1 for op in pl["ops"] if op["kind"] == "allgather")
                sl = dict(base)
                sl["out1"] = np.full((o, v), np.nan)
                sl["out2"] = np.full((o, o, v, v), np.nan)
                Interp(pl, sl, alpha=alpha or 0.0, allgather=allgather).run()
                ref = (orc.tupdate(t1, t2, fsp=fsp, alpha=alpha, equation=eq) if fn == "tupdate"
                       else orc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha, equation=eq))
                worst = max(worst, np.abs(sl["out1"] - ref[0]).max(), np.abs(sl["out2"] - ref[1]).max())
        # the rdm1 plan: the products over the doubles are split over an occupied index and summed over ranks
        pl = plan_json(lib, o, v, "gamma", 0, rank=rank, world=world, int8_digits=int8, vvvv_planes=bool(int8),
                       ovvv_planes=bool(int8) and o % 8 == 0 and v % 8 == 0)
        ncoll += sum(1 for op in pl["ops"] if op["kind"] == "allgather")
        sl = dict(base)
        sl["rdm1"] = np.full((o + v, o + v), np.nan)
        Interp(pl, sl, allgather=allgather).run()
        worst = max(worst, np.abs(sl["rdm1"] - orc.gamma(t1, t2, l1, l2)).max())
        q.put((rank, float(worst), ncoll))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("antisym", [True, False])
@pytest.mark.parametrize("world,ov,int8", [(2, (4, 6), 0), (2, (5, 7), 0), (3, (5, 7), 0), (2, (5, 7), 7), (2, (8, 16), 6), (8, (5, 9), 6)])
def test_sharded_plans_match_oracle(built_lib, world, ov, antisym, int8):
    o, v = ov
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 7 * world + (3 if antisym else 0) + o + (11 if int8 else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, o, v, antisym, q, int8)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, worst, ncoll in res:
        assert worst < (1e-12 if int8 else TOL), (rank, worst)
        assert ncoll >= 10          # K1/K2 ladders, R1-R4, R6-R9 are distributed
