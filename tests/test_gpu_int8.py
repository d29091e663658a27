"""GPU tests of the INT8 tensor-core GEMM engine (csrc/ozaki.cu: tcgen05.mma kind::i8 + TMEM) that the
contraction plans use for the large products of the CC residual (CCSD.py:305, 411, 470, 484, 602):
digit cut bit-exact vs numpy, the product vs exact integer arithmetic and vs FP64, chunked plane sets,
and the whole residual at sizes where every tile path (multi-wave persistent loop, k flushes, padding)
is exercised, against the FP64 DMMA engine."""
import numpy as np
import pytest

from plan_interp import oz_digits, oz_to_planes, oz_stats, oz_value

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ecw(built_lib):
    import ecw_cc_b200
    return ecw_cc_b200


def _split(ecw, X, ns, transposed=False, row0=0, total=0, planes=None, scale=None):
    import torch
    lib = ecw.lib
    st = torch.cuda.current_stream().cuda_stream
    if transposed:
        K, R = X.shape
        rs, ks = 1, X.stride(0)
    else:
        R, K = X.shape
        rs, ks = X.stride(0), 1
    tot = total or R
    if planes is None:
        planes = torch.zeros(lib.ecw_ozaki_plane_bytes(tot, K, ns), dtype=torch.int8, device="cuda")
        scale = torch.zeros(lib.ecw_ozaki_stat_elems(tot), dtype=torch.float64, device="cuda")
    assert lib.ecw_ozaki_split_rows(X.data_ptr(), R, K, rs, ks, ns, planes.data_ptr(), scale.data_ptr(), row0, tot, st) == 0
    return planes, scale


@pytest.mark.parametrize("R,K,ns", [(1, 1, 3), (128, 32, 6), (200, 100, 7), (780, 333, 6), (64, 4096, 8), (257, 31, 5)])
def test_digit_cut_bit_exact(ecw, R, K, ns):
    import torch
    rng = np.random.default_rng(R + K)
    X = rng.standard_normal((R, K)) * np.exp(rng.uniform(-8, 2, (R, 1)))
    if R > 5 and K > 7:
        X[3, :] = 0.0
        X[5, 7] = 0.0
    D, s = oz_digits(X, ns)
    ref = oz_to_planes(D)
    Xd = torch.from_numpy(X).cuda()
    for tr in (False, True):
        planes, scale = _split(ecw, Xd.t().contiguous() if tr else Xd, ns, transposed=tr)
        assert np.array_equal(planes.cpu().numpy()[: ref.size], ref), tr
        st, want = scale.cpu().numpy(), oz_stats(X, s)
        Rp = want.size // 2
        assert np.array_equal(st[:Rp], want[:Rp])                                   # power-of-two scales
        assert np.abs(st[Rp:] - want[Rp:]).max() <= 1e-13 * max(1.0, K ** 0.5)     # row sums (summation order differs)
    # the digits stand for X up to half a unit of the last digit of y = (x/s+1)/2 — a whole unit in the rare
    # case where rounding the last digit up would carry (it is clamped at 255) — plus the FP64 rounding of y
    err = np.abs(oz_value(D, s, ns) - X) / s[:, None]
    assert np.all(err <= 2 * 256.0 ** (-ns) + 2.0 ** -52)
    assert np.mean(err <= 256.0 ** (-ns) + 2.0 ** -52) > 0.99


def test_chunked_plane_set_equals_whole(ecw):
    """ecw_eris_vvvv_planes cuts the packed vvvv in row chunks: same bytes as one cut of the whole."""
    import torch
    rng = np.random.default_rng(2)
    R, K, ns = 700, 150, 7
    X = torch.from_numpy(rng.standard_normal((R, K))).cuda()
    whole, sw = _split(ecw, X, ns)
    planes = torch.full_like(whole, 77)
    scale = torch.full_like(sw, -1.0)
    for r0 in range(0, R, 256):
        nr = min(256, R - r0)
        _split(ecw, X[r0:r0 + nr], ns, row0=r0, total=R, planes=planes, scale=scale)
    assert torch.equal(planes, whole) and torch.equal(scale, sw)


@pytest.mark.parametrize("M,N,K,ns", [(1, 1, 1, 6), (128, 80, 32, 6), (130, 90, 40, 6), (300, 100, 70, 7), (256, 192, 512, 8),
                                      (100, 300, 100, 5), (200, 96, 64, 4), (64, 64, 64, 3),
                                      (780, 1000, 4000, 6), (500, 300, 70000, 6), (19000, 200, 64, 6)])
def test_int8_gemm_vs_fp64(ecw, M, N, K, ns):
    """alpha A B^T + beta C from digit planes vs numpy FP64 (long-double accumulate for the reference);
    K = 70000 spans nine int32 drain intervals, 19000 rows need a second wave of the persistent CTAs."""
    import torch
    lib = ecw.lib
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(M * 7 + N)
    A = rng.standard_normal((M, K)) * 0.02
    B = rng.standard_normal((N, K)) * 0.01
    C0 = rng.standard_normal((M, N))
    ref = 0.5 * (A @ B.T) + 0.25 * C0
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    pa, sa = _split(ecw, dA, ns)
    pb, sb = _split(ecw, dB, ns)
    # worst case (ozaki.cu header): (NS+3) K 256^-NS sA sB with sA sB < 4 max|A| max|B|, plus FP64 rounding
    bound = 4 * (ns + 3) * K * np.abs(A).max() * np.abs(B).max() * 256.0 ** (-ns) + 1e-15 * K ** 0.5
    C = torch.from_numpy(C0).cuda()
    assert lib.ecw_ozaki_gemm(pa.data_ptr(), sa.data_ptr(), pb.data_ptr(), sb.data_ptr(), M, N, K, C.data_ptr(), N, 1,
                              0.5, 0.25, ns, st) == 0
    assert np.abs(C.cpu().numpy() - ref).max() < bound
    # swapped roles, column-major target: the same matrix through strides (1, N)
    Ct = torch.from_numpy(C0).cuda()
    assert lib.ecw_ozaki_gemm(pb.data_ptr(), sb.data_ptr(), pa.data_ptr(), sa.data_ptr(), N, M, K, Ct.data_ptr(), 1, N,
                              0.5, 0.25, ns, st) == 0
    assert np.abs(Ct.cpu().numpy() - ref).max() < bound


def test_int8_gemm_exact_on_digit_operands(ecw):
    """Operands that ARE short digit strings (multiples of 2^-15 of the row scale): cut without loss, and the
    product is exact up to the FP64 rounding of the epilogue."""
    import torch
    lib = ecw.lib
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(11)
    M, N, K, ns = 260, 130, 100, 7
    A = rng.integers(-60, 61, (M, K)).astype(np.float64) + rng.integers(-60, 61, (M, K)) / 128.0
    B = rng.integers(-60, 61, (N, K)).astype(np.float64)
    pa, sa = _split(ecw, torch.from_numpy(A).cuda(), ns)
    pb, sb = _split(ecw, torch.from_numpy(B).cuda(), ns)
    C = torch.zeros((M, N), dtype=torch.float64, device="cuda")
    assert lib.ecw_ozaki_gemm(pa.data_ptr(), sa.data_ptr(), pb.data_ptr(), sb.data_ptr(), M, N, K, C.data_ptr(), N, 1,
                              1.0, 0.0, ns, st) == 0
    ref = A @ B.T                                   # exact in FP64: integers below 2^53 after scaling by 2^14
    assert np.abs(C.cpu().numpy() - ref).max() <= 2.0 ** -48 * K * 64.0 * 64.0      # FP64 epilogue at scale K sA sB


@pytest.mark.parametrize("ov,antisym", [((16, 96), True), ((12, 72), False)])
def test_residual_int8_vs_dmma_engine(ecw, ov, antisym):
    """Whole T/Lambda residuals at a size where the GEMMs span many tiles and waves: INT8 engine (every
    unbatched GEMM, packed vvvv as digit planes) vs the FP64 DMMA engine on the same device inputs."""
    import torch
    o, v = ov
    n = o + v
    res = {}
    for eng in ("dmma", "int8"):
        de = ecw.DeviceEris.synthetic(o, v, gemm=eng, int8_min_flops=-1.0)
        cc = ecw.GCC(de, assume_antisym=antisym)
        t1, t2 = de.synth_tensor("t1", (o, v)), de.synth_tensor("t2", (o, o, v, v))
        l1, l2 = de.synth_tensor("l1", (o, v)), de.synth_tensor("l2", (o, o, v, v))
        if not antisym:
            g = torch.Generator(device="cuda").manual_seed(1)
            t2 = t2 + 0.01 * torch.randn(t2.shape, dtype=torch.float64, device="cuda", generator=g)
            l2 = l2 + 0.01 * torch.randn(l2.shape, dtype=torch.float64, device="cuda", generator=g)
        fsp = de.synth_tensor("fsp", (n, n))
        res[eng] = (cc.tupdate(t1, t2, fsp=fsp, equation=True) + cc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=1e-3)
                    + (cc.gamma(t1, t2, l1, l2),))
        del cc, de
        torch.cuda.empty_cache()
    for x, y in zip(res["dmma"], res["int8"]):
        assert float((x - y).abs().max()) < 1e-11
