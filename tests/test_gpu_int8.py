"""GPU tests of the INT8 tensor-core GEMM engine (csrc/ozaki.cu: tcgen05.mma kind::i8 + TMEM) that the
contraction plans use for the large products of the CC residual (CCSD.py:305, 411, 470, 484, 602):
digit cut bit-exact vs numpy, the product vs exact integer arithmetic and vs FP64, chunked plane sets,
and the whole residual at sizes where every tile path (multi-wave persistent loop, k flushes, padding)
is exercised, against the FP64 DMMA engine."""
import numpy as np
import pytest

from plan_interp import oz_digits, oz_to_planes, oz_stats, oz_value, oz_const_slots, oz_product

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
    assert torch.equal(planes[:-4096], whole[:-4096]) and torch.equal(scale, sw)      # the last 4096 bytes are slack


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


def test_two_level_cut_and_batched_products(ecw):
    """k = (k1, k2) with k2 padded per k1, per-k1 row sums, and batches over row blocks / single k1 values of a plane
    set — the access patterns of the ovvv terms (CCSD.py:294, 311-312) — against numpy."""
    import ctypes
    import torch
    lib = ecw.lib
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(4)
    ns, R, K1, K2, nb, Mb, Nb = 6, 96, 5, 40, 3, 32, 16
    X = rng.standard_normal((R, K1, K2)) * 0.03                  # rows r, k = (k1, k2)
    Xd = torch.from_numpy(X).cuda()
    planes = torch.zeros(lib.ecw_ozaki_plane_bytes2(R, K1, K2, ns), dtype=torch.int8, device="cuda")
    stats = torch.zeros(lib.ecw_ozaki_stat_elems2(R, K1), dtype=torch.float64, device="cuda")
    assert lib.ecw_ozaki_split2(Xd.data_ptr(), R, K1, K2, K1 * K2, K2, 1, ns, planes.data_ptr(), stats.data_ptr(), st) == 0
    pl, want = oz_const_slots(X.reshape(R, K1 * K2), ns, K1=K1)
    nbytes = planes.numel() - 4096
    assert np.array_equal(planes.cpu().numpy()[:nbytes], pl.view(np.int8)[:nbytes])
    assert np.abs(stats.cpu().numpy() - want).max() < 1e-13
    # the same set cut from the transposed storage (rows contiguous)
    Xt = torch.from_numpy(np.ascontiguousarray(X.transpose(1, 2, 0))).cuda()     # [k1, k2, r]
    planes2, stats2 = torch.zeros_like(planes), torch.zeros_like(stats)
    assert lib.ecw_ozaki_split2(Xt.data_ptr(), R, K1, K2, 1, K2 * R, R, ns, planes2.data_ptr(), stats2.data_ptr(), st) == 0
    assert torch.equal(planes2[:nbytes], planes[:nbytes]) and float((stats2 - stats).abs().max()) < 1e-13
    # (1) batch over row blocks: C_b = X[b*Mb:(b+1)*Mb, :] . Y[b*Nb:(b+1)*Nb, :]^T over the whole k
    Y = rng.standard_normal((nb * Nb, K1, K2)) * 0.02
    Yd = torch.from_numpy(Y).cuda()
    pY = torch.zeros(lib.ecw_ozaki_plane_bytes2(nb * Nb, K1, K2, ns), dtype=torch.int8, device="cuda")
    sY = torch.zeros(lib.ecw_ozaki_stat_elems2(nb * Nb, K1), dtype=torch.float64, device="cuda")
    assert lib.ecw_ozaki_split2(Yd.data_ptr(), nb * Nb, K1, K2, K1 * K2, K2, 1, ns, pY.data_ptr(), sY.data_ptr(), st) == 0
    C = torch.zeros((nb, Mb, Nb), dtype=torch.float64, device="cuda")
    nkb2 = (K2 + 31) // 32
    bt = (ctypes.c_int64 * 15)(nb, 0, Mb, 0, Nb, 0, 0, 0, 0, 128, 0, 128, 0, Mb * Nb, K1 * nkb2)
    assert lib.ecw_ozaki_gemm_batched(planes.data_ptr(), stats.data_ptr(), R, pY.data_ptr(), sY.data_ptr(), nb * Nb, Mb, Nb,
                                      K1 * K2, C.data_ptr(), Nb, 1, 1.0, 0.0, ns, bt, st) == 0
    Xf, Yf = X.reshape(R, -1), Y.reshape(nb * Nb, -1)
    ref = np.stack([Xf[b * Mb:(b + 1) * Mb] @ Yf[b * Nb:(b + 1) * Nb].T for b in range(nb)])
    assert np.abs(C.cpu().numpy() - ref).max() < 1e-13
    # (2) batch over single k1 values of X against ONE small operand Z[Nz, K2]: C_b = X[:, b, :] . Z^T
    Nz = 24
    Z = rng.standard_normal((Nz, K2)) * 0.05
    Zd = torch.from_numpy(Z).cuda()
    pZ = torch.zeros(lib.ecw_ozaki_plane_bytes2(Nz, 1, K2, ns), dtype=torch.int8, device="cuda")
    sZ = torch.zeros(lib.ecw_ozaki_stat_elems2(Nz, 1), dtype=torch.float64, device="cuda")
    assert lib.ecw_ozaki_split2(Zd.data_ptr(), Nz, 1, K2, K2, 0, 1, ns, pZ.data_ptr(), sZ.data_ptr(), st) == 0
    C2 = torch.zeros((K1, Nz, R), dtype=torch.float64, device="cuda")          # [b][n][m]: tile rows contiguous
    bt = (ctypes.c_int64 * 15)(K1, 0, 0, 0, 0, 0, nkb2, 0, 0, 2 * 128, 128, 128, 0, Nz * R, nkb2)
    assert lib.ecw_ozaki_gemm_batched(planes.data_ptr(), stats.data_ptr(), R, pZ.data_ptr(), sZ.data_ptr(), Nz, R, Nz, K2,
                                      C2.data_ptr(), 1, R, 1.0, 0.0, ns, bt, st) == 0
    ref2 = np.stack([(X[:, b, :] @ Z.T).T for b in range(K1)])
    assert np.abs(C2.cpu().numpy() - ref2).max() < 1e-13


def test_ovvv_plane_sets_bit_exact(ecw):
    """ecw_eris_ovvv_planes: both orientations of ovvv_p == the numpy statement of the cut; the FP64 layout is dropped."""
    from oracle import synth
    from oracle import refactored_np as R
    o, v = 8, 16
    de = ecw.DeviceEris.synthetic(o, v, gemm="int8")
    assert de.use_ovvv_planes and "ovvv_p" not in de.buf
    E = R.DeviceErisSpec(synth.SynthEris(o, v))
    O = E.ovvv_p.reshape(o * v, -1)
    for name, X, K1 in (("ovvv_oz1", O, 1), ("ovvv_oz2", np.ascontiguousarray(O.T), o)):
        pl, st = oz_const_slots(X, de.int8_digits, K1=K1)
        nb = pl.size * 8 - 4096 - (pl.size * 8 - 4096) % 8
        assert np.array_equal(de.buf[name].cpu().numpy()[:nb], pl.view(np.int8)[:nb]), name
        got = de.buf[name + "s"].cpu().numpy()
        assert np.array_equal(got[: st.size // (2 + K1 if K1 > 1 else 2)], st[: st.size // (2 + K1 if K1 > 1 else 2)])
        assert np.abs(got - st).max() < 1e-13, name


@pytest.mark.parametrize("ns", [6, 7, 8])
def test_int8_accumulator_headroom(ecw, ns):
    """Worst case of the int32 accumulators: every digit of both operands is -128 (x = -(1 - 2^-49) s_r), so each of
    the up to NS products per k adds +2^14 and nothing cancels between the drains to FP64 (csrc/ozaki.cu, oz_kflush).  K spans several drains; the result must be the exact K x^2
    to FP64 rounding."""
    import torch
    lib = ecw.lib
    st = torch.cuda.current_stream().cuda_stream
    M, N, K = 130, 90, 40000
    x = -(1.0 - 2.0 ** -49)
    A = torch.full((M, K), x, dtype=torch.float64, device="cuda")
    B = torch.full((N, K), x, dtype=torch.float64, device="cuda")
    pa, sa = _split(ecw, A, ns)
    pb, sb = _split(ecw, B, ns)
    assert int(pa[:1024].to(torch.int64).max()) == -128          # k-block 0, digit 0, rows 0..31: every digit is -128
    C = torch.zeros((M, N), dtype=torch.float64, device="cuda")
    assert lib.ecw_ozaki_gemm(pa.data_ptr(), sa.data_ptr(), pb.data_ptr(), sb.data_ptr(), M, N, K, C.data_ptr(), N, 1,
                              1.0, 0.0, ns, st) == 0
    want = K * x * x
    # representation error of x with NS digits: |delta| <= 256^-NS per element, twice K of them at |x| ~ 1
    assert float((C - want).abs().max()) <= 2.5 * K * 256.0 ** -ns + 1e-9
