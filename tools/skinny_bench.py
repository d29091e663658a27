"""The short-K / skinny FP64 products of the residual at (40,400) through ecw_dgemm, CUDA-event timed (and the target of
an `ncu --set full` capture).  Usage: python tools/skinny_bench.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ecw_cc_b200 import lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
st = torch.cuda.current_stream().cuda_stream
shapes = [("[pk,kqrs->pqrs] rank-40 update", 40, 6400000, 40, 0, 0, 1.0),
          ("[jbmi,ma->jiba] rank-40 update", 640000, 400, 40, 0, 0, 1.0),
          ("ovvv.t1 [mbef,jf->jbme]", 6400000, 40, 400, 0, 1, 0.0)]
for name, M, N, K, ta, tb, beta in shapes:
    A = torch.randn((K, M) if ta else (M, K), dtype=torch.float64, device="cuda")
    B = torch.randn((N, K) if tb else (K, N), dtype=torch.float64, device="cuda")
    C = torch.zeros((M, N), dtype=torch.float64, device="cuda")
    best = 1e9
    for r in range(reps + 1):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = lib.ecw_dgemm(ta, tb, M, N, K, 1.0, A.data_ptr(), A.shape[1], B.data_ptr(), B.shape[1], beta, C.data_ptr(), N, -1, st)
        e.record()
        e.synchronize()
        assert rc == 0
        if r:
            best = min(best, s.elapsed_time(e))
    gb = 8.0 * (M * K + N * K + M * N * (2 if beta else 1)) / 1e9
    print("%-34s M%-8d N%-8d K%-4d  %6.2f ms  %5.1f TFLOP/s  %5.2f TB/s (algorithmic bytes)" % (
        name, M, N, K, best, 2.0 * M * N * K / best / 1e9, gb / best), flush=True)
    del A, B, C
    torch.cuda.empty_cache()
