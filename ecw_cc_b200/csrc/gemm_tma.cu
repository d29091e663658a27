// gemm_tma.cu — warp-specialised FP64 GEMM for sm_100a: TMA producer + mbarrier ring + DMMA consumers.
//
// Same contract as gemm.cu (C = alpha op(A) op(B) + beta C, batched / split-K), used when both operands
// are 16-byte aligned with even leading dimensions (every large GEMM of the CC path).  One producer
// warp issues cp.async.bulk.tensor (TMA) loads of whole 128-byte-swizzled tiles into a STAGES-deep
// shared-memory ring and signals `full` mbarriers by transaction bytes; eight consumer warps wait on
// `full`, feed DMMA.8x8x4 from LDS.64 fragment reads and release the slot through `empty` mbarriers.
// No __syncthreads and no address arithmetic in the math warps: the FP64 tensor pipe stays busy
// while the next tiles stream in.
//
// Bank conflicts.  64-bit fragments cannot use ldmatrix, and TMA writes dense tiles (no padding), so
// the 128B swizzle (16-byte chunk index XOR row%8) is combined with a permutation of which tile
// row each lane group owns:
//   K-contiguous tile  [rows][16 k]   : lane group g owns row  rho(g) = 2*(g%4) + g/4  of an 8-row group;
//   M/N-contiguous tile [16-col box][16 k][16 cols]: lane group g owns column mu(g) = {0,1,8,9,2,3,10,11}[g]
//                                        (+4 for the odd tile of a pair) of a 16-column box.
// With either map the 16 lanes of a half-warp hit 16 distinct 8-byte bank pairs at every k-step.
#include <cuda.h>

#include "kernels.h"

namespace ecw {

namespace {

constexpr int TBK = 16;   // k extent of a tile = one 128-byte swizzle row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
          "r"(smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ int rho(int g) { return 2 * (g & 3) + (g >> 2); }
__device__ __forceinline__ int mu(int g) { return ((g & 1) | ((g & 2) << 2) | ((g & 4) >> 1)); }   // {0,1,8,9,2,3,10,11}

// Thread layout: warps [0, NCW) are DMMA consumers (two warpgroups), warps [NCW, NCW+4) form the
// producer warpgroup (one elected lane issues TMA).  Registers are re-balanced with setmaxnreg
// (producer 40, consumers 232) so the 64x32 warp tiles keep their 128 accumulator registers unspilled.
template <int BM, int BN, int WM, int WN, int TA, int TB, int STAGES>
__global__ void __launch_bounds__(((BM / WM) * (BN / WN) + 4) * 32, 1)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs p) {
  constexpr int NCW = (BM / WM) * (BN / WN);
  constexpr int MI = WM / 8, NI = WN / 8;
  constexpr int A_BYTES = BM * TBK * 8, B_BYTES = BN * TBK * 8, STAGE_BYTES = A_BYTES + B_BYTES;
  static_assert(TA == 0 || WM % 16 == 0, "M-contiguous tiles pair two m8 tiles per 16-column box");
  static_assert(TB == 1 || WN % 16 == 0, "N-contiguous tiles pair two n8 tiles per 16-column box");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- tile coordinates (grouped rasterisation as in gemm.cu)
  const int64_t tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  int64_t tm, tn;
  {
    const int64_t GROUP = 8;
    int64_t tile = blockIdx.x, per_group = GROUP * tiles_n, gid = tile / per_group, first_m = gid * GROUP;
    int64_t gsz = min(tiles_m - first_m, GROUP);
    tm = first_m + (tile % per_group) % gsz;
    tn = (tile % per_group) / gsz;
  }
  const int64_t m0 = tm * BM, n0 = tn * BN;
  const int64_t zb = blockIdx.y;
  const int64_t r = zb / p.splitk, s = zb % p.splitk;
  const int64_t kbeg = s * p.kchunk, kend = min(p.K, kbeg + p.kchunk);
  const int ktiles = (int)((kend - kbeg + TBK - 1) / TBK);

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= NCW) {
    // =================== TMA producer warpgroup (one elected lane) ===================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    if (warp == NCW && lane == 0) {
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB) : "memory");
      const int ra = p.sA ? (int)r : 0, rb = p.sB ? (int)r : 0;
      for (int kt = 0; kt < ktiles; ++kt) {
        const int st = kt % STAGES, use = kt / STAGES;
        if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
        unsigned char* sa = smem + st * STAGE_BYTES;
        unsigned char* sb = sa + A_BYTES;
        const int k0 = (int)(kbeg + (int64_t)kt * TBK);
        mbar_expect_tx(&full[st], STAGE_BYTES);
        if (TA == 0) {
          tma_load_3d(sa, &tmA, &full[st], k0, (int)m0, ra);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 16; ++j) tma_load_3d(sa + j * 2048, &tmA, &full[st], (int)m0 + 16 * j, k0, ra);
        }
        if (TB == 1) {
          tma_load_3d(sb, &tmB, &full[st], k0, (int)n0, rb);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 16; ++j) tma_load_3d(sb + j * 2048, &tmB, &full[st], (int)n0 + 16 * j, k0, rb);
        }
      }
    }
    return;
  }

  // =================== DMMA consumers ===================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
  const int g = lane >> 2, tig = lane & 3;
  const int wm0 = (warp % (BM / WM)) * WM;
  const int wn0 = (warp / (BM / WM)) * WN;

  // per-lane byte offsets of the fragment elements inside a tile, for the 4 k4-steps of a k-tile
  int aoff[4], boff[4];
  int abase, bbase;
  if (TA == 0) {
    const int rr = rho(g);
    abase = (wm0 + rr) * 128;
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) aoff[s4] = (((2 * s4 + (tig >> 1)) ^ rr) << 4) + ((tig & 1) << 3);
  } else {
    const int mm = mu(g);
    abase = (wm0 >> 4) * 2048 + ((mm & 1) << 3);
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      const int k = 4 * s4 + tig;
      aoff[s4] = k * 128 + (((mm >> 1) ^ (k & 7)) << 4);   // even tile of a pair; the odd tile adds 4 columns: chunk ^= 2
    }
  }
  if (TB == 1) {
    const int rr = rho(g);
    bbase = (wn0 + rr) * 128;
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) boff[s4] = (((2 * s4 + (tig >> 1)) ^ rr) << 4) + ((tig & 1) << 3);
  } else {
    const int mm = mu(g);
    bbase = (wn0 >> 4) * 2048 + ((mm & 1) << 3);
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      const int k = 4 * s4 + tig;
      boff[s4] = k * 128 + (((mm >> 1) ^ (k & 7)) << 4);
    }
  }

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int kt = 0; kt < ktiles; ++kt) {
    const int st = kt % STAGES, use = kt / STAGES;
    mbar_wait(&full[st], use & 1);
    const unsigned char* sa = smem + st * STAGE_BYTES + abase;
    const unsigned char* sb = smem + st * STAGE_BYTES + A_BYTES + bbase;
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      double af[MI], bf[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        if (TA == 0) af[i] = *reinterpret_cast<const double*>(sa + i * 1024 + aoff[s4]);
        else af[i] = *reinterpret_cast<const double*>(sa + (i >> 1) * 2048 + (aoff[s4] ^ ((i & 1) << 5)));
      }
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        if (TB == 1) bf[j] = *reinterpret_cast<const double*>(sb + j * 1024 + boff[s4]);
        else bf[j] = *reinterpret_cast<const double*>(sb + (j >> 1) * 2048 + (boff[s4] ^ ((j & 1) << 5)));
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }

  // ---- epilogue: logical fragment coordinates -> tile coordinates through the same permutations.  With beta != 0 the
  // old values of two fragment rows (4 NI scalars) are fetched before the first store (see gemm.cu).
  const double alpha = p.alpha, beta = p.beta;
  double* __restrict__ C = p.C + zb * p.sC;
  auto row_of = [&](int i) { return m0 + wm0 + (TA == 0 ? i * 8 + rho(g) : (i >> 1) * 16 + mu(g) + (i & 1) * 4); };
  auto col_of = [&](int j, int e) {
    const int nl = 2 * tig + e;
    return n0 + wn0 + (TB == 1 ? j * 8 + rho(nl) : (j >> 1) * 16 + mu(nl) + (j & 1) * 4);
  };
#pragma unroll
  for (int i0 = 0; i0 < MI; i0 += 2) {
    double old[2][NI][2];
    if (beta != 0.0) {
#pragma unroll
      for (int ii = 0; ii < 2; ++ii) {
        if (i0 + ii >= MI) continue;
        const int64_t m = row_of(i0 + ii);
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int64_t n = col_of(j, e);
            old[ii][j][e] = (m < p.M && n < p.N) ? C[m * p.ldc + n] : 0.0;
          }
      }
    }
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      const int i = i0 + ii;
      if (i >= MI) continue;
      const int64_t m = row_of(i);
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < NI; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int64_t n = col_of(j, e);
          if (n >= p.N) continue;
          double v = alpha * acc[i][j][e];
          if (beta != 0.0) v += beta * old[ii][j][e];
          C[m * p.ldc + n] = v;
        }
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 3-D map over one operand: `kmajor` = the k index is contiguous ([rows][K], leading dim ld);
// otherwise the M/N index is contiguous ([K][cols], leading dim ld).  Third dim = batch.
bool make_map(CUtensorMap* tm, const double* base, bool kmajor, int64_t rows, int64_t K, int64_t ld, int64_t nb,
              int64_t bstride, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3], strides[2];
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  const int64_t nbatch = bstride ? nb : 1;
  if (kmajor) {
    dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows; box[0] = TBK; box[1] = (cuuint32_t)box_rows;
  } else {
    dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K; box[0] = 16; box[1] = TBK;
  }
  dims[2] = (cuuint64_t)nbatch;
  box[2] = 1;
  strides[0] = (cuuint64_t)ld * 8;
  strides[1] = (cuuint64_t)(bstride ? bstride : (kmajor ? rows : K) * ld) * 8;
  if (strides[0] % 16 || strides[1] % 16) return false;
  CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS;
}

template <int BM, int BN, int WM, int WN, int STAGES>
cudaError_t launch_tma_cfg(const GemmArgs& p, cudaStream_t st) {
  constexpr int THREADS = ((BM / WM) * (BN / WN) + 4) * 32;
  static_assert((BM / WM) * (BN / WN) == 8, "two consumer warpgroups");
  constexpr size_t SMEM = (size_t)STAGES * (BM + BN) * TBK * 8 + 2 * STAGES * 8 + 1024;
  CUtensorMap tmA, tmB;
  const int64_t nb = p.batch / p.splitk;
  if (!make_map(&tmA, p.A, p.ta == 0, p.M, p.K, p.lda, nb, p.sA, BM)) return cudaErrorNotSupported;
  if (!make_map(&tmB, p.B, p.tb == 1, p.N, p.K, p.ldb, nb, p.sB, BN)) return cudaErrorNotSupported;
  int64_t tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  dim3 grid((unsigned)tiles, (unsigned)p.batch, 1);
#define ECW_TMA_LAUNCH(TA_, TB_)                                                                         \
  {                                                                                                      \
    auto kern = dgemm_tma_kernel<BM, BN, WM, WN, TA_, TB_, STAGES>;                                       \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);   \
    if (e != cudaSuccess) return e;                                                                      \
    kern<<<grid, THREADS, SMEM, st>>>(tmA, tmB, p);                                                       \
    return cudaGetLastError();                                                                           \
  }
  if (p.ta == 0 && p.tb == 1) ECW_TMA_LAUNCH(0, 1)
  if constexpr (WN % 16 == 0) {
    if (p.ta == 0 && p.tb == 0) ECW_TMA_LAUNCH(0, 0)
  }
  if constexpr (WM % 16 == 0) {
    if (p.ta == 1 && p.tb == 1) ECW_TMA_LAUNCH(1, 1)
    if constexpr (WN % 16 == 0) {
      if (p.ta == 1 && p.tb == 0) ECW_TMA_LAUNCH(1, 0)
    }
  }
#undef ECW_TMA_LAUNCH
  return cudaErrorNotSupported;
}

}  // namespace

bool gemm_tma_eligible(const GemmArgs& p) {
  auto aligned = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!aligned(p.A) || !aligned(p.B)) return false;
  if ((p.lda & 1) || (p.ldb & 1) || (p.sA & 1) || (p.sB & 1)) return false;
  if (p.M > 0x7fffffff || p.N > 0x7fffffff || p.K > 0x7fffffff) return false;
  if (p.splitk > 1 && p.kchunk % TBK) return false;
  return get_encode() != nullptr;
}

// cfg 20: 128x128 (8 consumer warps 64x32), cfg 21: 112x128 (56x32, K-contiguous A only)
cudaError_t launch_gemm_tma(const GemmArgs& p, cudaStream_t st, int cfg) {
  if (cfg == 21 && p.ta == 0) return launch_tma_cfg<112, 128, 56, 32, 6>(p, st);
  return launch_tma_cfg<128, 128, 64, 32, 6>(p, st);
}

}  // namespace ecw
