// ozaki.cu — FP64 GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
//
// Blackwell's tcgen05 path has no FP64 kind; the native FP64 tensor instruction is the legacy DMMA
// (gemm_tma.cu, 37 TFLOP/s pipe peak).  The large contractions of the CC residual (CCSD.py:305 ladder,
// :411 ring, :602 Lambda intermediates) are instead evaluated with an error-free splitting (Ozaki
// scheme I) on the INT8 tensor pipe (4.5 POP/s):
//
//   row r of an operand X[R,K] is scaled by 2^e_r (|X[r,:]| < 2^e_r) and cut into NS signed digits
//   of 7 bits (the first of 6):  X[r,k] = s_r * sum_p d_p[r,k] * 128^-p,  s_r = 2^(e_r-6),
//   d_p in [-64, 64] (int8).  Every step of the cut is exact in FP64.
//   C[m,n] = sA_m sB_n * sum_{p+q<NS} 128^-(p+q) * (A_p . B_q^T)[m,n]      (triangular truncation)
//   Each int8 product A_p . B_q^T is exact in the int32 accumulator (|d d'| <= 2^12, flushed to FP64
//   every 32768 k, at most NS products per accumulator: < 2^31).  The only approximation is the
//   truncation p+q >= NS: |dC| <= (NS+1)/4 * 2^-(7 NS - 2) * K * 2^(e_m + f_n)  — for NS = 7 that is
//   2^-47 relative to K*rowmax*colmax, i.e. at the FP64 rounding level of the DMMA result itself.
//
// Data layout (memory laid out for the MMA, not for the host): an operand is stored as *digit planes*
//   planes[kb][p][rg][j][ri][16]   kb = k/32, p = digit, rg = r/8, j = (k%32)/16, ri = r%8
// i.e. for a fixed (k-block, digit) all rows are contiguous in the tcgen05 "no-swizzle K-major" core
// matrix order (8 rows x 16 bytes = 128 contiguous bytes; SBO = 256 B between 8-row groups, LBO =
// 128 B between the two 16-byte k-chunks).  A 128-row A tile of one digit is ONE contiguous 4 KB span,
// a 64-row B tile a 2 KB span: the producer streams them with linear bulk TMA (cp.async.bulk), no
// tensor maps, no swizzle, and any operand can serve on either side.  Rows are padded to 128, k to 32.
//
// Kernel (one CTA per SM, persistent over 128x64 output tiles, 192 threads):
//   warp 0   lane 0: TMA producer — per k-block one stage = NS A-planes + NS B-planes (42 KB at NS=7),
//                    5-stage mbarrier ring;
//   warp 1   lane 0: MMA issuer — tcgen05.mma kind::i8 (s8 x s8 -> s32, M128 K32); product (p,q)
//                    accumulates into TMEM columns [64(p+q), 64(p+q)+64): all products of one weight share
//                    an accumulator and all NS digits of A and B are loaded once per k-block.  Digits
//                    q = 0..NS-1-p of B are contiguous in shared memory, so they are issued as one MMA of
//                    N = 64 (NS-p) columns (split at 256): 10 instructions per stage at NS = 7;
//   warps 2-5      : epilogue — tcgen05.ld the NS accumulators, Horner in FP64
//                    (acc_d + 2^-7 (acc_{d+1} + ...)), scale by sA sB alpha, add beta C, store.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>

#include "kernels.h"

namespace ecw {

namespace {

constexpr int OZ_BK = 32;        // k per stage = K of one kind::i8 MMA
constexpr int OZ_TM = 128;       // tile rows  (MMA M)
constexpr int OZ_TN = 64;        // tile cols  (MMA N); NS accumulators of 64 columns fit TMEM's 512
constexpr int OZ_KFLUSH = 32768; // int32 accumulators are drained to FP64 at least every OZ_KFLUSH k
constexpr int OZ_SMEM_MAX = 227 * 1024;

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
// linear bulk TMA: global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// D[tmem] (+)= A[smem] . B[smem]^T, signed 8-bit operands, int32 accumulate; one thread issues for the CTA
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
// 8 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle (core matrix = 8 rows x 16 bytes, 128 B contiguous):
// bits [0,14) start>>4, [16,30) leading byte offset>>4 (between the two 16-byte k-chunks),
// [32,46) stride byte offset>>4 (between 8-row groups), [46,48) version = 1 (sm_100), [61,64) layout = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}

struct OzGemmParams {
  const int8_t* pa;      // digit planes of the M-side operand
  const int8_t* pb;      // digit planes of the N-side operand
  const double* sa;      // row scales 2^(e-6), length Mp
  const double* sb;
  double* C;
  int64_t M, N, K;       // logical extents
  int64_t Mp, Np;        // padded row counts of the two plane sets (multiples of 128)
  int64_t crs, ccs;      // C[m*crs + n*ccs]
  double alpha, beta;
  uint32_t lbo, sbo;     // descriptor strides (bytes)
};

template <int NS>
struct OzCfg {
  static constexpr int A_PLANE = OZ_TM * OZ_BK;           // 4096
  static constexpr int B_PLANE = OZ_TN * OZ_BK;           // 2048
  static constexpr int STAGE = NS * (A_PLANE + B_PLANE);
  static constexpr int STAGES = (OZ_SMEM_MAX - 2048) / STAGE > 8 ? 8 : (OZ_SMEM_MAX - 2048) / STAGE;
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 512;
};

// tile index -> (m block, n block); groups of 8 m-blocks are swept along n so that the CTAs running
// together share A and B panels in L2
__device__ __forceinline__ void tile_coord(int64_t tile, int64_t tiles_m, int64_t tiles_n, int64_t& tm, int64_t& tn) {
  const int64_t GROUP = 8;
  const int64_t per_group = GROUP * tiles_n, gid = tile / per_group, first_m = gid * GROUP;
  const int64_t gsz = min(tiles_m - first_m, GROUP);
  tm = first_m + (tile % per_group) % gsz;
  tn = (tile % per_group) / gsz;
}

template <int NS>
__global__ void __launch_bounds__(192, 1) ozaki_gemm_kernel(OzGemmParams p) {
  using Cfg = OzCfg<NS>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles_m = (p.M + OZ_TM - 1) / OZ_TM, tiles_n = (p.N + OZ_TN - 1) / OZ_TN;
  const int64_t ntiles = tiles_m * tiles_n;
  const int nkb = (int)((p.K + OZ_BK - 1) / OZ_BK);
  constexpr int KB_FLUSH = OZ_KFLUSH / OZ_BK;
  const int nchunk = (nkb + KB_FLUSH - 1) / KB_FLUSH;

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =================== TMA producer ===================
    if (lane == 0) {
      const int64_t a_slab = p.Mp * OZ_BK;   // bytes of one (k-block, digit) slab
      const int64_t b_slab = p.Np * OZ_BK;
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int64_t tm, tn;
        tile_coord(tile, tiles_m, tiles_n, tm, tn);
        const int8_t* ga = p.pa + tm * Cfg::A_PLANE;
        const int8_t* gb = p.pb + tn * Cfg::B_PLANE;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t st = it % STAGES, use = it / STAGES;
          mbar_wait(&empty[st], (use & 1) ^ 1);
          unsigned char* s = smem + st * Cfg::STAGE;
          mbar_expect_tx(&full[st], Cfg::STAGE);
#pragma unroll
          for (int d = 0; d < NS; ++d) {
            bulk_load(s + d * Cfg::A_PLANE, ga + ((int64_t)kb * NS + d) * a_slab, Cfg::A_PLANE, &full[st]);
            bulk_load(s + NS * Cfg::A_PLANE + d * Cfg::B_PLANE, gb + ((int64_t)kb * NS + d) * b_slab, Cfg::B_PLANE,
                      &full[st]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =================== MMA issuer ===================
    if (lane == 0) {
      // instruction descriptor: D = s32 (2 @bit4), A = B = signed 8-bit (1 @bit7, 1 @bit10), both K-major,
      // N>>3 @bit17, M>>4 @bit24
      const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_TM >> 4) << 24);
      const uint64_t desc0 = make_desc(0, p.lbo, p.sbo);
      const uint32_t sbase = smem_u32(smem);
      uint32_t it = 0, acc_it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int c = 0; c < nchunk; ++c, ++acc_it) {
          mbar_wait(tmem_empty, (acc_it & 1) ^ 1);
          tc_fence_after();
          const int kb0 = c * KB_FLUSH, kb1 = min(nkb, kb0 + KB_FLUSH);
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const uint32_t st = it % STAGES, use = it / STAGES;
            mbar_wait(&full[st], use & 1);
            tc_fence_after();
            const uint32_t sa = sbase + st * Cfg::STAGE, sb = sa + NS * Cfg::A_PLANE;
            // The B digit planes of a stage are contiguous, so digits q = 0..NS-1-pd form ONE operand of
            // 64 (NS-pd) rows: A_pd . [B_0; ..; B_{NS-1-pd}]^T lands in the TMEM columns of the weights
            // pd .. NS-1, which are contiguous too.  One MMA per <= 256 columns: A is read from shared
            // memory once per 256 output columns instead of once per 64.
#pragma unroll
            for (int pd = 0; pd < NS; ++pd) {
              const uint64_t adesc = desc0 | (uint64_t)(((sa + pd * Cfg::A_PLANE) >> 4) & 0x3fff);
              const int ncols = OZ_TN * (NS - pd);
#pragma unroll
              for (int n0 = 0; n0 < ncols; n0 += 256) {
                const int nlen = ncols - n0 < 256 ? ncols - n0 : 256;
                const uint64_t bdesc = desc0 | (uint64_t)(((sb + n0 * OZ_BK) >> 4) & 0x3fff);
                mma_i8(tmem_base + (uint32_t)(pd * OZ_TN + n0), adesc, bdesc, idesc0 | ((uint32_t)(nlen >> 3) << 17),
                       (kb > kb0 || pd > 0) ? 1u : 0u);
              }
            }
            mma_commit(&empty[st]);
          }
          mma_commit(tmem_full);
        }
      }
    }
  } else {
    // =================== epilogue (4 warps = 128 TMEM lanes = 128 tile rows) ===================
    const int quad = warp & 3;                 // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int row = quad * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const double w = 0.0078125;                // 2^-7
    uint32_t acc_it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int64_t tm, tn;
      tile_coord(tile, tiles_m, tiles_n, tm, tn);
      const int64_t m = tm * OZ_TM + row, n0 = tn * OZ_TN;
      double run[OZ_TN];
#pragma unroll
      for (int j = 0; j < OZ_TN; ++j) run[j] = 0.0;
      for (int c = 0; c < nchunk; ++c, ++acc_it) {
        mbar_wait(tmem_full, acc_it & 1);
        tc_fence_after();
#pragma unroll
        for (int cb = 0; cb < OZ_TN / 8; ++cb) {
          int32_t v[NS][8];
#pragma unroll
          for (int d = 0; d < NS; ++d) tmem_ld8(tlane + (uint32_t)(d * OZ_TN + cb * 8), v[d]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            double s = (double)v[NS - 1][j];
#pragma unroll
            for (int d = NS - 2; d >= 0; --d) s = fma(s, w, (double)v[d][j]);
            run[cb * 8 + j] += s;
          }
        }
        tc_fence_before();
        mbar_arrive(tmem_empty);
      }
      if (m < p.M) {
        const double fa = p.alpha * p.sa[m];
        double* crow = p.C + m * p.crs + n0 * p.ccs;
        const int nn = (int)min((int64_t)OZ_TN, p.N - n0);
        if (p.ccs == 1 && nn == OZ_TN && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < OZ_TN; j += 2) {
            double2 o;
            o.x = fa * p.sb[n0 + j] * run[j];
            o.y = fa * p.sb[n0 + j + 1] * run[j + 1];
            if (p.beta != 0.0) {
              const double2 old = *reinterpret_cast<const double2*>(crow + j);
              o.x += p.beta * old.x;
              o.y += p.beta * old.y;
            }
            *reinterpret_cast<double2*>(crow + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < OZ_TN; ++j) {
            if (j < nn) {
              double o = fa * p.sb[n0 + j] * run[j];
              if (p.beta != 0.0) o += p.beta * crow[j * p.ccs];
              crow[j * p.ccs] = o;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// row scales: s_r = 2^(e_r - 6) with |X[r,:]| < 2^e_r  (s_r = 1 for an all-zero or padded row)
// X[r*rs + k*ks]; one of rs, ks is 1.
__global__ void ozaki_rowmax_kcontig(const double* __restrict__ X, int64_t R, int64_t K, int64_t rs, int64_t Rp,
                                     double* __restrict__ scale) {   // scale already offset to the chunk's first row
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= Rp) return;
  double mx = 0.0;
  if (r < R) {
    const double* x = X + r * rs;
    for (int64_t k = lane; k < K; k += 32) mx = fmax(mx, fabs(x[k]));
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) {
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);          // mx = f 2^e, f in [0.5, 1)
    scale[r] = mx > 0.0 ? ldexp(1.0, e - 6) : 1.0;
  }
}
__global__ void ozaki_rowmax_rcontig(const double* __restrict__ X, int64_t R, int64_t K, int64_t ks, int64_t Rp,
                                     double* __restrict__ scale) {
  __shared__ double red[8][33];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * 32 + lane;
  double mx = 0.0;
  if (r < R)
    for (int64_t k = wy; k < K; k += 8) mx = fmax(mx, fabs(X[r + k * ks]));
  red[wy][lane] = mx;
  __syncthreads();
  if (wy == 0 && r < Rp) {
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmax(mx, red[i][lane]);
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);
    scale[r] = mx > 0.0 ? ldexp(1.0, e - 6) : 1.0;
  }
}

// cut X into NS digit planes (layout in the header comment).  Block = 128 rows x one k-block (32 k),
// thread t: row t%128, 16-byte k-chunk t/128.
template <int NS>
__global__ void __launch_bounds__(256) ozaki_split_kernel(const double* __restrict__ X, int64_t R, int64_t K, int64_t rs,
                                                          int64_t ks, int64_t Rp, int64_t row0,
                                                          const double* __restrict__ scale,
                                                          int8_t* __restrict__ planes) {
  // X, scale: the chunk (local rows 0..R); planes: the whole set of Rp padded rows, chunk rows start at row0
  const int t = threadIdx.x;
  const int64_t r = (int64_t)blockIdx.x * 128 + (t & 127);
  const int j = t >> 7;
  const int64_t kb = blockIdx.y;
  const int64_t k0 = kb * OZ_BK + j * 16;
  double x[16];
  if (r < R) {
    const double inv = 1.0 / scale[r];     // exact: a power of two
    const double* src = X + r * rs + k0 * ks;
    if (ks == 1 && k0 + 16 <= K && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const double2 v = *reinterpret_cast<const double2*>(src + i);
        x[i] = v.x * inv;
        x[i + 1] = v.y * inv;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = (k0 + i < K) ? src[i * ks] * inv : 0.0;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = 0.0;
  }
  const int64_t slab = Rp * OZ_BK;
  const int64_t rg = row0 + r;
  int8_t* dst = planes + (kb * NS) * slab + (rg >> 3) * 256 + j * 128 + (rg & 7) * 16;
#pragma unroll
  for (int p = 0; p < NS; ++p) {
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const double d = rint(x[i]);
      x[i] = (x[i] - d) * 128.0;           // exact
      w[i >> 2] |= ((uint32_t)(__double2int_rn(d)) & 0xffu) << ((i & 3) * 8);
    }
    *reinterpret_cast<uint4*>(dst + p * slab) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

template <int NS>
cudaError_t launch_gemm_ns(const OzGemmParams& p, cudaStream_t st, int sm_count) {
  using Cfg = OzCfg<NS>;
  auto kern = ozaki_gemm_kernel<NS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  if (e != cudaSuccess) return e;
  const int64_t tiles = ((p.M + OZ_TM - 1) / OZ_TM) * ((p.N + OZ_TN - 1) / OZ_TN);
  const unsigned grid = (unsigned)(tiles < sm_count ? tiles : sm_count);
  kern<<<grid, 192, Cfg::SMEM, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

int64_t ozaki_padded_rows(int64_t R) { return (R + 127) / 128 * 128; }
int64_t ozaki_plane_bytes(int64_t R, int64_t K, int ns) {
  return ozaki_padded_rows(R) * ((K + OZ_BK - 1) / OZ_BK * OZ_BK) * ns;
}

cudaError_t launch_ozaki_split(const double* X, int64_t R, int64_t K, int64_t rs, int64_t ks, int ns, int8_t* planes,
                               double* scale, cudaStream_t st, int64_t row0, int64_t total_rows) {
  if (ns < 3 || ns > 8 || (rs != 1 && ks != 1) || (row0 & 127)) return cudaErrorInvalidValue;
  if (total_rows <= 0) total_rows = row0 + R;
  if (row0 + R > total_rows || (row0 + R < total_rows && (R & 127))) return cudaErrorInvalidValue;
  const int64_t Rp = ozaki_padded_rows(R);              // rows this launch writes (incl. zero padding)
  const int64_t Rp_total = ozaki_padded_rows(total_rows);
  double* sc = scale + row0;
  if (ks == 1) {
    ozaki_rowmax_kcontig<<<(unsigned)((Rp + 7) / 8), 256, 0, st>>>(X, R, K, rs, Rp, sc);
  } else {
    ozaki_rowmax_rcontig<<<(unsigned)((Rp + 31) / 32), 256, 0, st>>>(X, R, K, ks, Rp, sc);
  }
  const int64_t nkb = (K + OZ_BK - 1) / OZ_BK;
  if (nkb > 65535) return cudaErrorInvalidValue;
  dim3 grid((unsigned)(Rp / 128), (unsigned)nkb, 1);
  switch (ns) {
#define ECW_OZ_SPLIT(NS_) \
  case NS_: ozaki_split_kernel<NS_><<<grid, 256, 0, st>>>(X, R, K, rs, ks, Rp_total, row0, sc, planes); break;
    ECW_OZ_SPLIT(3) ECW_OZ_SPLIT(4) ECW_OZ_SPLIT(5) ECW_OZ_SPLIT(6) ECW_OZ_SPLIT(7) ECW_OZ_SPLIT(8)
#undef ECW_OZ_SPLIT
  }
  return cudaGetLastError();
}

cudaError_t launch_ozaki_gemm(const int8_t* pa, const double* sa, const int8_t* pb, const double* sb, int64_t M, int64_t N,
                              int64_t K, double* C, int64_t crs, int64_t ccs, double alpha, double beta, int ns,
                              cudaStream_t st, int sm_count) {
  OzGemmParams p{};
  p.pa = pa; p.pb = pb; p.sa = sa; p.sb = sb; p.C = C;
  p.M = M; p.N = N; p.K = K;
  p.Mp = ozaki_padded_rows(M); p.Np = ozaki_padded_rows(N);
  p.crs = crs; p.ccs = ccs; p.alpha = alpha; p.beta = beta;
  p.lbo = 128; p.sbo = 256;
  // bring-up overrides (tools/ozaki_check.py): descriptor strides
  if (const char* s = getenv("ECW_OZ_LBO")) p.lbo = (uint32_t)atoi(s);
  if (const char* s = getenv("ECW_OZ_SBO")) p.sbo = (uint32_t)atoi(s);
  if (sm_count <= 0) sm_count = 148;
  switch (ns) {
    case 3: return launch_gemm_ns<3>(p, st, sm_count);
    case 4: return launch_gemm_ns<4>(p, st, sm_count);
    case 5: return launch_gemm_ns<5>(p, st, sm_count);
    case 6: return launch_gemm_ns<6>(p, st, sm_count);
    case 7: return launch_gemm_ns<7>(p, st, sm_count);
    case 8: return launch_gemm_ns<8>(p, st, sm_count);
  }
  return cudaErrorInvalidValue;
}

}  // namespace ecw
