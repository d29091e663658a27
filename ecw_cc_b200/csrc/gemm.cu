// gemm.cu — FP64 tensor-core GEMM for sm_100a.
//
// Blackwell's tcgen05 path has no FP64 kind; the FP64 tensor pipe is reached
// through mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 — every larger PTX f64 shape is
// split into it by ptxas on sm_100a).  The kernel is therefore a multi-stage
// cp.async (LDGSTS) pipeline feeding register accumulators:
//   * CTA tile BM x BN x 16, STAGES-deep shared-memory ring, one barrier per
//     k-tile;
//   * both operands in either orientation (K-contiguous or M/N-contiguous), so
//     the index permutes of the contraction engine are absorbed by the loads;
//   * 64-bit fragments cannot use ldmatrix: tiles are padded by 4 doubles per
//     row, which makes every half-warp LDS.64 fragment read conflict-free;
//   * batched (free index or reduction index) and split-K launches write
//     partial products that reduce.cu sums deterministically;
//   * grouped tile rasterisation keeps an A row-panel group L2-resident.
// C = alpha * op(A) op(B) + beta * C, row-major C.
#include "kernels.h"

#include <cstdlib>

namespace ecw {

namespace {

constexpr int PAD = 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Load a tile whose contiguous direction is the K direction: smem[r][k], r < ROWS.
template <int ROWS, int THREADS, int BK>
__device__ __forceinline__ void load_kmajor(double* sm, const double* __restrict__ g, int64_t ld,
                                            int64_t row0, int64_t nrows, int64_t k0, int64_t kend,
                                            int vec, int tid) {
  constexpr int LDS = BK + PAD;
  if (vec) {
    constexpr int CH = ROWS * (BK / 2);
#pragma unroll
    for (int c = tid; c < CH; c += THREADS) {
      int r = c / (BK / 2), kc = (c % (BK / 2)) * 2;
      int64_t gr = row0 + r, gk = k0 + kc;
      int bytes = 0;
      if (gr < nrows && gk < kend) bytes = (gk + 1 < kend) ? 16 : 8;
      const double* src = bytes ? g + gr * ld + gk : g;
      cp_async16(sm + r * LDS + kc, src, bytes);
    }
  } else {
    constexpr int CH = ROWS * BK;
#pragma unroll
    for (int c = tid; c < CH; c += THREADS) {
      int r = c / BK, kc = c % BK;
      int64_t gr = row0 + r, gk = k0 + kc;
      int bytes = (gr < nrows && gk < kend) ? 8 : 0;
      const double* src = bytes ? g + gr * ld + gk : g;
      cp_async8(sm + r * LDS + kc, src, bytes);
    }
  }
}

// Load a tile whose contiguous direction is the M/N direction: smem[k][c], c < COLS.
template <int COLS, int THREADS, int BK>
__device__ __forceinline__ void load_mnmajor(double* sm, const double* __restrict__ g, int64_t ld,
                                             int64_t col0, int64_t ncols, int64_t k0, int64_t kend,
                                             int vec, int tid) {
  constexpr int LDS = COLS + PAD;
  if (vec) {
    constexpr int CPR = COLS / 2;
    constexpr int CH = BK * CPR;
#pragma unroll
    for (int c = tid; c < CH; c += THREADS) {
      int kr = c / CPR, cc = (c % CPR) * 2;
      int64_t gk = k0 + kr, gc = col0 + cc;
      int bytes = 0;
      if (gk < kend && gc < ncols) bytes = (gc + 1 < ncols) ? 16 : 8;
      const double* src = bytes ? g + gk * ld + gc : g;
      cp_async16(sm + kr * LDS + cc, src, bytes);
    }
  } else {
    constexpr int CH = BK * COLS;
#pragma unroll
    for (int c = tid; c < CH; c += THREADS) {
      int kr = c / COLS, cc = c % COLS;
      int64_t gk = k0 + kr, gc = col0 + cc;
      int bytes = (gk < kend && gc < ncols) ? 8 : 0;
      const double* src = bytes ? g + gk * ld + gc : g;
      cp_async8(sm + kr * LDS + cc, src, bytes);
    }
  }
}

template <int BM, int BN, int BK, int TA, int TB>
struct SmemLayout {
  static constexpr int A_ELEMS = TA ? BK * (BM + PAD) : BM * (BK + PAD);
  static constexpr int B_ELEMS = TB ? BN * (BK + PAD) : BK * (BN + PAD);
  static constexpr int STAGE = A_ELEMS + B_ELEMS;
};

template <int BM, int BN, int WM, int WN, int BK, int TA, int TB, int STAGES, int ILV>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, (BM * BN <= 128 * 48) ? 2 : 1)
dgemm_kernel(GemmArgs p) {
  constexpr int THREADS = (BM / WM) * (BN / WN) * 32;
  constexpr int MI = WM / 8, NI = WN / 8;
  using L = SmemLayout<BM, BN, BK, TA, TB>;
  extern __shared__ __align__(16) double smem[];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int wm0 = (warp % (BM / WM)) * WM;
  const int wn0 = (warp / (BM / WM)) * WN;

  // ---- tile coordinates (grouped rasterisation: GROUP row-tiles share column sweeps)
  const int64_t tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  int64_t tile = blockIdx.x;
  int64_t tm, tn;
  {
    const int64_t GROUP = 8;
    int64_t per_group = GROUP * tiles_n;
    int64_t gid = tile / per_group;
    int64_t first_m = gid * GROUP;
    int64_t gsz = min(tiles_m - first_m, GROUP);
    tm = first_m + (tile % per_group) % gsz;
    tn = (tile % per_group) / gsz;
  }
  const int64_t m0 = tm * BM, n0 = tn * BN;

  // ---- batch / split-K slice
  const int64_t zb = blockIdx.y;
  const int64_t r = zb / p.splitk, s = zb % p.splitk;
  const int64_t kbeg = s * p.kchunk;
  const int64_t kend = min(p.K, kbeg + p.kchunk);
  const double* __restrict__ A = p.A + r * p.sA;
  const double* __restrict__ B = p.B + r * p.sB;
  double* __restrict__ C = p.C + zb * p.sC;

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int64_t ktiles = (kend - kbeg + BK - 1) / BK;
  static_assert(BK % 4 == 0, "BK");

  auto load_stage = [&](int stage, int64_t kt) {
    double* sa = smem + stage * L::STAGE;
    double* sb = sa + L::A_ELEMS;
    const int64_t k0 = kbeg + kt * BK;
    if (TA == 0) load_kmajor<BM, THREADS, BK>(sa, A, p.lda, m0, p.M, k0, kend, p.vecA, tid);
    else load_mnmajor<BM, THREADS, BK>(sa, A, p.lda, m0, p.M, k0, kend, p.vecA, tid);
    if (TB == 0) load_mnmajor<BN, THREADS, BK>(sb, B, p.ldb, n0, p.N, k0, kend, p.vecB, tid);
    else load_kmajor<BN, THREADS, BK>(sb, B, p.ldb, n0, p.N, k0, kend, p.vecB, tid);
  };

#pragma unroll
  for (int st = 0; st < STAGES - 1; ++st) {
    if (st < ktiles) load_stage(st, st);
    cp_async_commit();
  }

  for (int64_t kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (!ILV) {
      int64_t nk = kt + STAGES - 1;
      if (nk < ktiles) load_stage((int)(nk % STAGES), nk);
      cp_async_commit();
    }
    const double* sa = smem + (kt % STAGES) * L::STAGE;
    const double* sb = sa + L::A_ELEMS;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double af[MI], bf[NI];
      const int kk = k4 * 4 + tig;
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int m = wm0 + i * 8 + g;
        af[i] = TA ? sa[kk * (BM + PAD) + m] : sa[m * (BK + PAD) + kk];
      }
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int n = wn0 + j * 8 + g;
        bf[j] = TB ? sb[n * (BK + PAD) + kk] : sb[kk * (BN + PAD) + n];
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      if (ILV && k4 == 0) {
        // issue the next stage's copies in the shadow of the DMMAs just queued: the slot being
        // refilled was last read in iteration kt-1, i.e. before the barrier above
        int64_t nk = kt + STAGES - 1;
        if (nk < ktiles) load_stage((int)(nk % STAGES), nk);
        cp_async_commit();
      }
    }
  }
  cp_async_wait<0>();

  // ---- epilogue
  const double alpha = p.alpha, beta = p.beta;
  const bool vecC = p.vecC;
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int64_t m = m0 + wm0 + i * 8 + g;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int64_t n = n0 + wn0 + j * 8 + 2 * tig;
      if (n >= p.N) continue;
      double* c = C + m * p.ldc + n;
      double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
      if (vecC && n + 1 < p.N) {
        if (beta != 0.0) {
          double2 old = *reinterpret_cast<const double2*>(c);
          v0 += beta * old.x;
          v1 += beta * old.y;
        }
        *reinterpret_cast<double2*>(c) = make_double2(v0, v1);
      } else {
        if (beta != 0.0) v0 += beta * c[0];
        c[0] = v0;
        if (n + 1 < p.N) {
          if (beta != 0.0) v1 += beta * c[1];
          c[1] = v1;
        }
      }
    }
  }
}

template <int BM, int BN, int WM, int WN, int BK, int STAGES, int ILV>
cudaError_t launch_cfg(const GemmArgs& p, cudaStream_t st) {
  constexpr int THREADS = (BM / WM) * (BN / WN) * 32;
  int64_t tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  dim3 grid((unsigned)tiles, (unsigned)p.batch, 1);
#define ECW_LAUNCH(TA_, TB_)                                                                        \
  {                                                                                                 \
    auto kern = dgemm_kernel<BM, BN, WM, WN, BK, TA_, TB_, STAGES, ILV>;                                      \
    size_t smem = sizeof(double) * STAGES * SmemLayout<BM, BN, BK, TA_, TB_>::STAGE;                     \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                 \
    kern<<<grid, THREADS, smem, st>>>(p);                                                           \
    return cudaGetLastError();                                                                      \
  }
  if (p.ta == 0 && p.tb == 0) ECW_LAUNCH(0, 0)
  if (p.ta == 0 && p.tb == 1) ECW_LAUNCH(0, 1)
  if (p.ta == 1 && p.tb == 0) ECW_LAUNCH(1, 0)
  ECW_LAUNCH(1, 1)
#undef ECW_LAUNCH
}

}  // namespace

int gemm_pick_config(int64_t M, int64_t N) {
  // 1: 32x128  2: 128x32  3: 128x8  4: 64x64  8: 128x128 (16 warps, interleaved loads)
  // 10: 112x128  11: 96x128 (8 warps) — chosen when they cut the padded-row waste of the M dimension
  // 12: 48x128  13: 128x48 — one tile covers a whole occupied index of up to 48 (no operand re-read)
  if (N <= 8) return 3;
  if (M <= 32) return 1;
  if (M <= 48) return 12;
  if (N <= 32) return 2;
  if (N <= 48) return 13;
  if (M <= 96 || N <= 96) return 4;
  auto padded = [](int64_t x, int64_t b) { return (x + b - 1) / b * b; };
  double w128 = (double)padded(M, 128) * padded(N, 128);
  double w64 = (double)padded(M, 64) * padded(N, 64);
  if (w64 * 1.15 < w128) return 4;
  double p128 = (double)padded(M, 128), p112 = (double)padded(M, 112), p96 = (double)padded(M, 96);
  if (p112 * 1.04 < p128 && p112 <= p96) return 10;
  if (p96 * 1.06 < p128 && p96 < p112) return 11;
  return 8;
}

cudaError_t launch_gemm(const GemmArgs& args, cudaStream_t st, int force_cfg) {
  GemmArgs p = args;
  if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return cudaSuccess;
  if (p.splitk < 1) p.splitk = 1;
  if (p.splitk == 1) p.kchunk = p.K;
  auto aligned = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.vecA = aligned(p.A) && (p.lda % 2 == 0) && (p.sA % 2 == 0);
  p.vecB = aligned(p.B) && (p.ldb % 2 == 0) && (p.sB % 2 == 0);
  p.vecC = aligned(p.C) && (p.ldc % 2 == 0) && (p.sC % 2 == 0);
  if (p.batch > 65535) return cudaErrorInvalidValue;
  int cfg = force_cfg >= 0 ? force_cfg : gemm_pick_config(p.M, p.N);
  // large aligned problems: TMA producer + mbarrier ring + DMMA consumers (gemm_tma.cu)
  static const bool no_tma = getenv("ECW_NO_TMA") != nullptr;
  if (cfg >= 20 || (force_cfg < 0 && !no_tma && (cfg == 8 || cfg == 10 || cfg == 11))) {
    if (gemm_tma_eligible(p)) {
      int tcfg = cfg >= 20 ? cfg : ((cfg != 8 && p.ta == 0) ? 21 : 20);
      cudaError_t e = launch_gemm_tma(p, st, tcfg);
      if (e != cudaErrorNotSupported) return e;
      cudaGetLastError();
    }
    if (cfg >= 20) cfg = 8;
  }
  switch (cfg) {
    case 0: return launch_cfg<128, 128, 32, 32, 16, 4, 0>(p, st);
    case 1: return launch_cfg<32, 128, 16, 32, 16, 4, 0>(p, st);
    case 2: return launch_cfg<128, 32, 32, 16, 16, 4, 0>(p, st);
    case 3: return launch_cfg<128, 8, 16, 8, 16, 4, 0>(p, st);
    case 4: return launch_cfg<64, 64, 32, 16, 16, 4, 0>(p, st);
    case 5: return launch_cfg<128, 128, 64, 32, 16, 4, 0>(p, st);
    case 6: return launch_cfg<128, 128, 64, 32, 16, 4, 1>(p, st);
    case 7: return launch_cfg<128, 128, 64, 32, 32, 3, 1>(p, st);
    case 8: return launch_cfg<128, 128, 32, 32, 16, 4, 1>(p, st);
    case 9: return launch_cfg<128, 128, 32, 32, 32, 3, 1>(p, st);
    case 10: return launch_cfg<112, 128, 56, 32, 16, 4, 1>(p, st);
    case 12: return launch_cfg<48, 128, 48, 16, 16, 4, 1>(p, st);
    case 13: return launch_cfg<128, 48, 32, 24, 16, 4, 1>(p, st);
    case 11: return launch_cfg<96, 128, 48, 32, 16, 4, 1>(p, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace ecw
