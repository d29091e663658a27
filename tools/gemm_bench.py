"""FP64 GEMM microbenchmark: ecw_dgemm tile configurations vs cuBLAS DGEMM (torch.matmul fp64),
CUDA-event timed.  Usage: python tools/gemm_bench.py [out.json]"""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ecw_cc_b200
from ecw_cc_b200 import lib


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts)


def main():
    out = []
    st = torch.cuda.current_stream().cuda_stream
    shapes = [("cube8192", 8192, 8192, 8192, 0, 0), ("cube8192_nt", 8192, 8192, 8192, 0, 1),
              ("ring16000", 16000, 16000, 16000, 0, 0), ("ladder780", 780, 16384, 79800, 0, 1),
              ("R6_nn", 780, 79800, 16000, 0, 0)]
    for name, M, N, K, ta, tb in shapes:
        A = torch.randn((K, M) if ta else (M, K), dtype=torch.float64, device="cuda")
        B = torch.randn((N, K) if tb else (K, N), dtype=torch.float64, device="cuda")
        C = torch.empty((M, N), dtype=torch.float64, device="cuda")
        fl = 2.0 * M * N * K
        Aop = A.t() if ta else A
        Bop = B.t() if tb else B
        ms = timeit(lambda: torch.matmul(Aop, Bop, out=C))
        ref = C.clone()
        rec = {"shape": name, "M": M, "N": N, "K": K, "ta": ta, "tb": tb, "cublas_tflops": fl / ms / 1e9}
        for cfg in (8, 10, 20, 21):
            def run():
                rc = lib.ecw_dgemm(ta, tb, M, N, K, 1.0, A.data_ptr(), A.shape[1], B.data_ptr(), B.shape[1], 0.0,
                                   C.data_ptr(), N, cfg, st)
                assert rc == 0
            ms = timeit(run)
            rec["cfg%d_tflops" % cfg] = fl / ms / 1e9
            rec["cfg%d_maxerr" % cfg] = float((C - ref).abs().max())
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del A, B, C, ref
        torch.cuda.empty_cache()
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
