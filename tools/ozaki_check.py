"""Bring-up and microbenchmark of the INT8-tcgen05 FP64 GEMM (csrc/ozaki.cu) on a B200.

  python tools/ozaki_check.py            # runs every stage in its own subprocess (a hang cannot block the rest)
  python tools/ozaki_check.py <stage>    # split | mma | probe | gemm | perf

Reference for the numbers is cuBLAS DGEMM (torch.matmul, fp64) and a numpy restatement of the digit cut.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def np_scale(X):
    mx = np.abs(X).max(axis=1)
    _, e = np.frexp(mx)
    return np.where(mx > 0, np.ldexp(1.0, e - 6), 1.0)


def np_digits(X, ns):
    """oracle of ozaki_split_kernel: digits[p][r][k] (int8) and row scales."""
    s = np_scale(X)
    x = X / s[:, None]
    out = []
    for _ in range(ns):
        d = np.rint(x)
        x = (x - d) * 128.0
        out.append(d.astype(np.int8))
    return np.stack(out), s


def to_planes(D):
    """digits [ns][R][K] -> device plane order [kb][p][rg][j][ri][16] with rows padded to 128, k to 32."""
    ns, R, K = D.shape
    Rp, Kp = (R + 127) // 128 * 128, (K + 31) // 32 * 32
    P = np.zeros((ns, Rp, Kp), dtype=np.int8)
    P[:, :R, :K] = D
    return np.ascontiguousarray(P.reshape(ns, Rp // 8, 8, Kp // 32, 2, 16).transpose(3, 0, 1, 4, 2, 5))


def from_planes(buf, ns, R, K):
    Rp, Kp = (R + 127) // 128 * 128, (K + 31) // 32 * 32
    P = buf.reshape(Kp // 32, ns, Rp // 8, 2, 8, 16).transpose(1, 2, 4, 0, 3, 5).reshape(ns, Rp, Kp)
    return P[:, :R, :K]


def dev():
    import torch
    import ecw_cc_b200
    from ecw_cc_b200 import lib
    return torch, lib, torch.cuda.current_stream().cuda_stream


def split_dev(torch, lib, st, X, ns, transposed=False):
    """X: torch [R,K] (or its [K,R] storage when transposed). returns (planes int8 tensor, scale tensor)."""
    if transposed:
        K, R = X.shape
        rs, ks = 1, X.stride(0)
    else:
        R, K = X.shape
        rs, ks = X.stride(0), 1
    nb = lib.ecw_ozaki_plane_bytes(R, K, ns)
    planes = torch.empty(nb, dtype=torch.int8, device="cuda")
    scale = torch.empty(lib.ecw_ozaki_padded_rows(R), dtype=torch.float64, device="cuda")
    rc = lib.ecw_ozaki_split(X.data_ptr(), R, K, rs, ks, ns, planes.data_ptr(), scale.data_ptr(), st)
    assert rc == 0, rc
    return planes, scale


def stage_split():
    torch, lib, st = dev()
    rng = np.random.default_rng(0)
    for (R, K, ns) in [(128, 32, 3), (200, 100, 7), (780, 333, 6), (64, 4096, 8)]:
        X = rng.standard_normal((R, K)) * np.exp(rng.uniform(-8, 2, (R, 1)))
        X[3, :] = 0.0
        X[5, 7] = 0.0
        Xd = torch.from_numpy(X).cuda()
        for tr in (False, True):
            src = Xd.t().contiguous() if tr else Xd
            planes, scale = split_dev(torch, lib, st, src, ns, transposed=tr)
            torch.cuda.synchronize()
            D, s = np_digits(X, ns)
            got = from_planes(planes.cpu().numpy(), ns, R, K)
            ok_d = np.array_equal(got, D)
            ok_s = np.array_equal(scale.cpu().numpy()[:R], s)
            rec = (s[:, None] * sum(D[p].astype(np.float64) * 128.0 ** (-p) for p in range(ns)))
            print("split R=%d K=%d ns=%d transposed=%s digits_equal=%s scales_equal=%s recon_err=%.2e pad_zero=%s" % (
                R, K, ns, tr, ok_d, ok_s, np.abs(rec - X).max() / np.abs(X).max(),
                bool((planes.cpu().numpy().astype(np.int64) != 0).sum() == (D != 0).sum())), flush=True)


def raw_gemm(torch, lib, st, DA, DB, ns, M, N, K, env=None):
    """C = sum_{p+q<ns} 128^-(p+q) A_p B_q^T from explicit digit arrays (scales = 1)."""
    pa = torch.from_numpy(to_planes(DA)).cuda()
    pb = torch.from_numpy(to_planes(DB)).cuda()
    sa = torch.ones((M + 127) // 128 * 128, dtype=torch.float64, device="cuda")
    sb = torch.ones((N + 127) // 128 * 128, dtype=torch.float64, device="cuda")
    C = torch.full((M, N), -777.0, dtype=torch.float64, device="cuda")
    for k in ("ECW_OZ_LBO", "ECW_OZ_SBO"):
        os.environ.pop(k, None)
    if env:
        os.environ.update(env)
    rc = lib.ecw_ozaki_gemm(pa.data_ptr(), sa.data_ptr(), pb.data_ptr(), sb.data_ptr(), M, N, K, C.data_ptr(), N, 1,
                            1.0, 0.0, ns, st)
    assert rc == 0, rc
    torch.cuda.synchronize()
    return C.cpu().numpy()


def raw_ref(DA, DB, ns):
    C = 0.0
    for p in range(ns):
        for q in range(ns - p):
            C = C + 128.0 ** (-(p + q)) * (DA[p].astype(np.float64) @ DB[q].astype(np.float64).T)
    return C


def stage_mma():
    torch, lib, st = dev()
    rng = np.random.default_rng(1)
    for (M, N, K, ns) in [(128, 64, 32, 3), (128, 64, 32, 7), (128, 64, 256, 7), (256, 192, 512, 7), (300, 100, 70, 6)]:
        DA = rng.integers(-64, 65, (ns, M, K)).astype(np.int8)
        DB = rng.integers(-64, 65, (ns, N, K)).astype(np.int8)
        ref = raw_ref(DA, DB, ns)
        for env in (None, {"ECW_OZ_LBO": "256", "ECW_OZ_SBO": "128"}):
            got = raw_gemm(torch, lib, st, DA, DB, ns, M, N, K, env)
            print("mma M=%d N=%d K=%d ns=%d desc=%s maxerr=%.3e (ref max %.3e)" % (
                M, N, K, ns, env or "default", np.abs(got - ref).max(), np.abs(ref).max()), flush=True)


def stage_probe():
    """one-hot A against an index-coded B: shows which B element the hardware pairs with each A position."""
    torch, lib, st = dev()
    M, N, K, ns = 128, 64, 32, 3
    kk, nn = np.meshgrid(np.arange(K), np.arange(N))
    code = ((kk + 32 * nn) % 127 - 63).astype(np.int8)          # B[n,k] coded
    DB = np.zeros((ns, N, K), np.int8)
    DB[0] = code
    for (m0, k0) in [(0, 0), (0, 1), (0, 16), (1, 0), (8, 0), (9, 17), (64, 5), (127, 31)]:
        DA = np.zeros((ns, M, K), np.int8)
        DA[0, m0, k0] = 1
        got = raw_gemm(torch, lib, st, DA, DB, ns, M, N, K)
        rows = np.nonzero(np.abs(got).sum(axis=1))[0]
        print("probe A[%d,%d]=1: nonzero rows %s; C[row,0:4]=%s expect row %d values %s" % (
            m0, k0, rows[:8].tolist(), got[rows[0], :4].tolist() if len(rows) else None, m0, code[:4, k0].tolist()),
            flush=True)


def stage_gemm():
    torch, lib, st = dev()
    torch.manual_seed(0)
    cases = [(128, 64, 32, 7, 0, 1), (256, 256, 1024, 7, 0, 1), (780, 1000, 4000, 7, 0, 1), (1000, 780, 4000, 7, 1, 0),
             (500, 300, 70000, 7, 0, 1), (2048, 2048, 2048, 6, 0, 0), (2048, 2048, 2048, 8, 1, 1),
             (4096, 4096, 4096, 7, 0, 1)]
    for (M, N, K, ns, ta, tb) in cases:
        A = torch.randn((K, M) if ta else (M, K), dtype=torch.float64, device="cuda") * 0.02
        B = torch.randn((N, K) if tb else (K, N), dtype=torch.float64, device="cuda") * 0.01
        Aop = A.t() if ta else A
        Bop = B.t() if tb else B
        C0 = torch.randn((M, N), dtype=torch.float64, device="cuda")
        ref = 0.5 * torch.matmul(Aop, Bop) + 0.25 * C0
        pa, sa = split_dev(torch, lib, st, A, ns, transposed=bool(ta))
        pb, sb = split_dev(torch, lib, st, B, ns, transposed=not bool(tb))
        C = C0.clone()
        rc = lib.ecw_ozaki_gemm(pa.data_ptr(), sa.data_ptr(), pb.data_ptr(), sb.data_ptr(), M, N, K, C.data_ptr(), N, 1,
                                0.5, 0.25, ns, st)
        assert rc == 0
        # swapped operand roles (tile rows run along n): same [M,N] result written with strides (1, N)
        Ct = C0.clone()
        rc = lib.ecw_ozaki_gemm(pb.data_ptr(), sb.data_ptr(), pa.data_ptr(), sa.data_ptr(), N, M, K, Ct.data_ptr(), 1, N,
                                0.5, 0.25, ns, st)
        assert rc == 0
        torch.cuda.synchronize()
        bound = K * float(Aop.abs().max()) * float(Bop.abs().max())
        print("gemm M=%d N=%d K=%d ns=%d ta=%d tb=%d maxerr=%.3e maxerr_T=%.3e (|C|max %.2e, K*amax*bmax %.2e)" % (
            M, N, K, ns, ta, tb, float((C - ref).abs().max()), float((Ct - ref).abs().max()),
            float(ref.abs().max()), bound), flush=True)


def timeit(torch, fn, reps=3, warm=1):
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


def stage_perf():
    torch, lib, st = dev()
    out = []
    shapes = [("cube8192", 8192, 8192, 8192), ("ring16000", 16000, 16000, 16000), ("ladderT", 79800, 780, 79800),
              ("R9", 16000, 780, 79800)]
    for name, M, N, K in shapes:
        for ns in (6, 7, 8):
            if name.startswith("ladder") and ns != 7:
                continue
            A = torch.randn((M, K), dtype=torch.float64, device="cuda") * 0.02
            B = torch.randn((N, K), dtype=torch.float64, device="cuda") * 0.01
            C = torch.empty((M, N), dtype=torch.float64, device="cuda")
            pa, sa = split_dev(torch, lib, st, A, ns)
            pb, sb = split_dev(torch, lib, st, B, ns)
            nbA = lib.ecw_ozaki_plane_bytes(M, K, ns)

            def run_split():
                assert lib.ecw_ozaki_split(A.data_ptr(), M, K, K, 1, ns, pa.data_ptr(), sa.data_ptr(), st) == 0

            def run_gemm():
                assert lib.ecw_ozaki_gemm(pa.data_ptr(), sa.data_ptr(), pb.data_ptr(), sb.data_ptr(), M, N, K,
                                          C.data_ptr(), N, 1, 1.0, 0.0, ns, st) == 0
            ms_s = timeit(torch, run_split)
            ms_g = timeit(torch, run_gemm)
            fl = 2.0 * M * N * K
            rec = {"shape": name, "M": M, "N": N, "K": K, "ns": ns, "gemm_ms": ms_g, "fp64_equiv_tflops": fl / ms_g / 1e9,
                   "int8_tops": fl * ns * (ns + 1) / 2 / ms_g / 1e9, "split_A_ms": ms_s,
                   "split_GBs": (M * K * 8 + nbA) / ms_s / 1e6}
            if M * N <= 16000 * 16000 and ns == 7 and M <= 16000:
                ref = torch.matmul(A, B.t())
                rec["maxerr"] = float((C - ref).abs().max())
                rec["cublas_ms"] = timeit(torch, lambda: torch.matmul(A, B.t(), out=ref))
                del ref
            print(json.dumps(rec), flush=True)
            out.append(rec)
            del A, B, C, pa, pb
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ozaki_perf.json"), "w"), indent=1)


STAGES = {"split": stage_split, "mma": stage_mma, "probe": stage_probe, "gemm": stage_gemm, "perf": stage_perf}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        STAGES[sys.argv[1]]()
    else:
        for name in ("split", "mma", "probe", "gemm", "perf"):
            print("=== stage %s" % name, flush=True)
            try:
                rc = subprocess.call([sys.executable, os.path.abspath(__file__), name], timeout=300 if name == "perf" else 120)
                print("=== stage %s rc=%s" % (name, rc), flush=True)
            except subprocess.TimeoutExpired:
                print("=== stage %s TIMEOUT (hang)" % name, flush=True)
