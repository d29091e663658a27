"""Multi-GPU parity check (launch with torchrun, one rank per GPU): every rank runs the sharded
CUDA path (vvvv_p row shard + distributed GEMMs + NCCL all-gathers) and compares with the CPU oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
import numpy as np
import torch
import torch.distributed as dist

import ecw_cc_b200 as ecw
from oracle import synth
from oracle.ccsd_np import OracleGCC

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
worst = 0.0
for (o, v) in [(5, 9), (8, 20), (10, 33), (8, 16), (16, 24)]:     # the last two: ovvv digit planes (nocc, nvir % 8 == 0)
    er = synth.SynthEris(o, v)
    fsp = synth.fsp(o, v)
    orc = OracleGCC(er)
    cc = ecw.GCC(er, rank=rank, world=world)
    for antisym in (True, False):
        if antisym:
            t1, t2, l1, l2 = synth.amplitudes(o, v)
        else:
            rng = np.random.default_rng(5)
            t1, l1 = 0.05 * rng.standard_normal((o, v)), 0.05 * rng.standard_normal((o, v))
            t2, l2 = 0.02 * rng.standard_normal((o, o, v, v)), 0.02 * rng.standard_normal((o, o, v, v))
        for alpha, eq in ((None, False), (1e-3, False), (None, True)):
            a, b = cc.tupdate(t1, t2, fsp=fsp, alpha=alpha, equation=eq)
            c, d = orc.tupdate(t1, t2, fsp=fsp, alpha=alpha, equation=eq)
            worst = max(worst, np.abs(a - c).max(), np.abs(b - d).max())
            a, b = cc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha, equation=eq)
            c, d = orc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha, equation=eq)
            worst = max(worst, np.abs(a - c).max(), np.abs(b - d).max())
        worst = max(worst, np.abs(cc.gamma(t1, t2, l1, l2) - orc.gamma(t1, t2, l1, l2)).max())
        worst = max(worst, abs(cc.energy(t1, t2, fsp) - orc.energy(t1, t2, fsp)))
# the collectives were enqueued by the library's own executor on its own ncclComm_t (no return to Python)
assert cc.eris.own_nccl and ecw.lib.ecw_ctx_nccl_ops(cc.eris._h) > 20, (cc.eris.own_nccl, ecw.lib.ecw_ctx_nccl_ops(cc.eris._h))
t = torch.tensor([worst], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("multi-GPU parity (world=%d): max abs diff vs oracle over all ranks = %.3e" % (world, float(t[0])), flush=True)
assert float(t[0]) < 1e-10
dist.destroy_process_group()
