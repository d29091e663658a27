"""Microbenchmark / spot check of the INT8-tcgen05 FP64 GEMM (csrc/ozaki.cu) on a B200.

  python tools/ozaki_check.py [digits ...]       # default: 6 7

For the shapes of the CC residual at (nocc, nvir) = (40, 400): time of the cut and of the product, FP64-equivalent
and int8 rates, max deviation from cuBLAS DGEMM (torch.matmul) where that fits.  Writes gpurun_out/ozaki_perf.json.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ecw_cc_b200
from ecw_cc_b200 import lib


def split_dev(st, X, ns):
    R, K = X.shape
    planes = torch.empty(lib.ecw_ozaki_plane_bytes(R, K, ns), dtype=torch.int8, device="cuda")
    stat = torch.empty(lib.ecw_ozaki_stat_elems(R), dtype=torch.float64, device="cuda")
    assert lib.ecw_ozaki_split(X.data_ptr(), R, K, X.stride(0), 1, ns, planes.data_ptr(), stat.data_ptr(), st) == 0
    return planes, stat


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
    digits = [int(x) for x in sys.argv[1:]] or [6, 7]
    st = torch.cuda.current_stream().cuda_stream
    out = []
    shapes = [("cube8192", 8192, 8192, 8192), ("ring16000", 16000, 16000, 16000), ("ladder", 79800, 780, 79800),
              ("R9", 16000, 780, 79800), ("hh", 79800, 780, 780), ("fvv", 400, 400, 640000)]
    for name, M, N, K in shapes:
        for ns in digits:
            A = torch.randn((M, K), dtype=torch.float64, device="cuda") * 0.02
            B = torch.randn((N, K), dtype=torch.float64, device="cuda") * 0.01
            C = torch.empty((M, N), dtype=torch.float64, device="cuda")
            pa, sa = split_dev(st, A, ns)
            pb, sb = split_dev(st, B, ns)

            def run_split():
                assert lib.ecw_ozaki_split(A.data_ptr(), M, K, K, 1, ns, pa.data_ptr(), sa.data_ptr(), st) == 0

            def run_gemm():
                assert lib.ecw_ozaki_gemm(pa.data_ptr(), sa.data_ptr(), pb.data_ptr(), sb.data_ptr(), M, N, K,
                                          C.data_ptr(), N, 1, 1.0, 0.0, ns, st) == 0
            ms_s, ms_g = timeit(run_split), timeit(run_gemm)
            fl = 2.0 * M * N * K
            rec = {"shape": name, "M": M, "N": N, "K": K, "digits": ns, "gemm_ms": ms_g,
                   "fp64_equiv_tflops": fl / ms_g / 1e9, "int8_tops": fl * ns * (ns + 1) / 2 / ms_g / 1e9,
                   "split_A_ms": ms_s, "split_GBs": M * K * (8 + ns) / ms_s / 1e6}
            if M * N <= 16000 * 16000 and M * K <= 16000 * 16000:
                ref = torch.matmul(A, B.t())
                rec["maxerr_vs_cublas"] = float((C - ref).abs().max())
                rec["cublas_ms"] = timeit(lambda: torch.matmul(A, B.t(), out=ref))
                del ref
            print(json.dumps(rec), flush=True)
            out.append(rec)
            del A, B, C, pa, pb
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ozaki_perf.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
