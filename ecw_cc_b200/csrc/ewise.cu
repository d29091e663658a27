// ewise.cu — bandwidth-bound kernels of the CC residual path (sm_100a):
// strided N-d permute/axpby (tiled transpose when the contiguous axis moves),
// split-K reduction, tau, antisymmetric pair pack/unpack, the residual->update
// "finish" kernel (soft-threshold + denominators, CCSD.py:316-338 and
// utilities.py:26-73), deterministic dot products, rdm1 assembly.
#include "kernels.h"

#include <algorithm>
#include <type_traits>

namespace ecw {

namespace {

constexpr int EW_THREADS = 256;

inline unsigned grid_for(int64_t n, int per_block, int64_t cap = 148LL * 32) {
  int64_t b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return (unsigned)std::min<int64_t>(b, cap);
}

// ------------------------------------------------------------------ permute
struct PermK {
  const double* in;
  double* out;
  int nd;                       // number of "rest" dims (linear kernel: all dims)
  int64_t dim[KMAXD], sin[KMAXD], sout[KMAXD];
  int64_t d_fi, d_fo;           // tiled kernel: extents of the two fast axes
  int64_t fi_sin, fi_sout, fo_sin, fo_sout;
  int64_t total;
  double alpha, beta;
};

// Same contiguous axis on both sides (or no unit stride at all): plain strided copy,
// dims ordered so the last one is the fast axis.  IT = index type of the element counter (32-bit divisions when the
// tensor has fewer than 2^31 elements), VEC = 2: the fast axis has unit stride on both sides and even extent, so a
// thread moves 16 bytes per access.  Four independent elements per thread and iteration keep enough loads in flight
// to reach HBM speed (the one-element version sat at ~2 TB/s, latency bound).
template <typename IT, int VEC>
__global__ void __launch_bounds__(EW_THREADS) permute_linear_kernel(PermK p) {
  using V = typename std::conditional<VEC == 2, double2, double>::type;
  const IT total = (IT)(p.total / VEC), stride = (IT)gridDim.x * (IT)blockDim.x;
  for (IT base = (IT)blockIdx.x * (IT)blockDim.x + (IT)threadIdx.x; base < total; base += 4 * stride) {
    int64_t oi[4], oo[4];
    bool ok[4];
    V vi[4], vo[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      IT rem = base + (IT)u * stride;
      ok[u] = rem < total;
      int64_t a = 0, b = 0;
#pragma unroll
      for (int d = KMAXD - 1; d >= 0; --d) {
        if (d < p.nd) {
          const IT dd = (IT)p.dim[d];
          const IT q = rem / dd, c = rem - q * dd;
          rem = q;
          a += (int64_t)c * p.sin[d];
          b += (int64_t)c * p.sout[d];
        }
      }
      oi[u] = a; oo[u] = b;
      if (ok[u]) {
        vi[u] = *reinterpret_cast<const V*>(p.in + a);
        if (p.beta != 0.0) vo[u] = *reinterpret_cast<const V*>(p.out + b);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      if constexpr (VEC == 2) {
        double2 r = make_double2(p.alpha * vi[u].x, p.alpha * vi[u].y);
        if (p.beta != 0.0) { r.x += p.beta * vo[u].x; r.y += p.beta * vo[u].y; }
        *reinterpret_cast<double2*>(p.out + oo[u]) = r;
      } else {
        double r = p.alpha * vi[u];
        if (p.beta != 0.0) r += p.beta * vo[u];
        p.out[oo[u]] = r;
      }
    }
  }
}

// The contiguous axis differs: 32x32 shared-memory transpose over (fi, fo); the
// remaining axes are enumerated by the block index.
__global__ void __launch_bounds__(256) permute_tiled_kernel(PermK p) {
  __shared__ double tile[32][33];
  const int64_t tiles_i = (p.d_fi + 31) / 32, tiles_o = (p.d_fo + 31) / 32;
  int64_t b = blockIdx.x;
  const int64_t ti = b % tiles_i; b /= tiles_i;
  const int64_t to = b % tiles_o; b /= tiles_o;
  int64_t oi = 0, oo = 0;
#pragma unroll
  for (int d = KMAXD - 1; d >= 0; --d) {
    if (d < p.nd) {
      int64_t q = b / p.dim[d];
      int64_t c = b - q * p.dim[d];
      b = q;
      oi += c * p.sin[d];
      oo += c * p.sout[d];
    }
  }
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int64_t i0 = ti * 32, o0 = to * 32;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    int64_t ii = i0 + tx, jo = o0 + r;
    if (ii < p.d_fi && jo < p.d_fo) tile[r][tx] = p.in[oi + ii * p.fi_sin + jo * p.fo_sin];
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    int64_t ii = i0 + r, jo = o0 + tx;
    if (ii < p.d_fi && jo < p.d_fo) {
      int64_t off = oo + ii * p.fi_sout + jo * p.fo_sout;
      double v = p.alpha * tile[tx][r];
      if (p.beta != 0.0) v += p.beta * p.out[off];
      p.out[off] = v;
    }
  }
}

__global__ void __launch_bounds__(EW_THREADS) ew2_kernel(Ew2Args p, int64_t total) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t rem = idx, oa = 0, ob = 0, oo = 0;
#pragma unroll
    for (int d = KMAXD - 1; d >= 0; --d) {
      if (d < p.nd) {
        int64_t q = rem / p.dim[d];
        int64_t c = rem - q * p.dim[d];
        rem = q;
        oa += c * p.sa[d];
        ob += c * p.sb[d];
        oo += c * p.so[d];
      }
    }
    double v = p.alpha;
    if (p.a) v *= p.a[oa];
    if (p.b) v *= p.b[ob];
    if (p.beta != 0.0) v += p.beta * p.out[oo];
    p.out[oo] = v;
  }
}

__global__ void __launch_bounds__(EW_THREADS) fill_kernel(double* c, int64_t n, double v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    c[i] = v;
}

__global__ void __launch_bounds__(EW_THREADS)
reduce_kernel(const double* __restrict__ part, int64_t nz, int64_t M, int64_t N, double* C, int64_t sr,
              int64_t sc, double alpha, double beta) {
  const int64_t mn = M * N;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < mn; e += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int64_t z = 0; z < nz; ++z) s += part[z * mn + e];
    int64_t m = e / N, n = e - m * N;
    int64_t off = m * sr + n * sc;
    double v = alpha * s;
    if (beta != 0.0) v += beta * C[off];
    C[off] = v;
  }
}

// two consecutive b per thread (v even) or one; the (i,j) pair comes from blockIdx.y so that only one 32-bit
// division per thread remains
template <int VEC>
__global__ void __launch_bounds__(EW_THREADS)
tau_kernel(const double* __restrict__ t2, const double* __restrict__ t1, double* __restrict__ out, int o, int v,
           double c1, double c2, int i0, int ni) {
  // t2 / out hold the rows i0 .. i0+ni-1 of the leading occupied index (a slab; i0 = 0, ni = o: everything)
  const uint32_t vv = (uint32_t)v * (uint32_t)v;
  for (int64_t ij = blockIdx.y; ij < (int64_t)ni * o; ij += gridDim.y) {     // grid-stride over the (i,j) rows
  const int i = (int)(ij / o), j = (int)(ij - (int64_t)i * o);
  const double* __restrict__ ti = t1 + (int64_t)(i + i0) * v;
  const double* __restrict__ tj = t1 + (int64_t)j * v;
  const int64_t row = ij * vv;
  for (uint32_t ab = (blockIdx.x * blockDim.x + threadIdx.x) * VEC; ab < vv; ab += gridDim.x * blockDim.x * VEC) {
    const uint32_t a = ab / (uint32_t)v, b = ab - a * (uint32_t)v;
    if constexpr (VEC == 2) {
      const double2 t = *reinterpret_cast<const double2*>(t2 + row + ab);
      const double ia = ti[a], ja = tj[a];
      double2 r;
      r.x = t.x + (c1 * (ia * tj[b]) - c2 * (ti[b] * ja));
      r.y = t.y + (c1 * (ia * tj[b + 1]) - c2 * (ti[b + 1] * ja));
      *reinterpret_cast<double2*>(out + row + ab) = r;
    } else {
      out[row + ab] = t2[row + ab] + (c1 * (ti[a] * tj[b]) - c2 * (ti[b] * tj[a]));
    }
  }
  }
}

// out[0] = max asymmetry of a doubles amplitude, out[1] = max |x| (non-negative doubles order like their bit patterns)
__global__ void __launch_bounds__(EW_THREADS)
defect_kernel(const double* __restrict__ x, int o, int v, double* out) {
  __shared__ double sh[EW_THREADS], sa[EW_THREADS];
  const uint32_t uv = (uint32_t)v, vv = uv * uv;
  double m = 0.0, am = 0.0;
  for (int64_t row = blockIdx.y; row < (int64_t)o * o; row += gridDim.y) {
    const int i = (int)(row / o), j = (int)(row - (int64_t)i * o);
    const double* __restrict__ xr = x + row * vv;
    const double* __restrict__ xt = x + ((int64_t)j * o + i) * vv;
    for (uint32_t ab = blockIdx.x * blockDim.x + threadIdx.x; ab < vv; ab += gridDim.x * blockDim.x) {
      const uint32_t a = ab / uv, b = ab - a * uv;
      const double val = xr[ab];
      am = fmax(am, fabs(val));
      m = fmax(m, fmax(fabs(val + xt[ab]), fabs(val + xr[b * uv + a])));
    }
  }
  sh[threadIdx.x] = m;
  sa[threadIdx.x] = am;
  __syncthreads();
  for (int w = EW_THREADS / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
      sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + w]);
      sa[threadIdx.x] = fmax(sa[threadIdx.x], sa[threadIdx.x + w]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(sh[0]));
    atomicMax(reinterpret_cast<unsigned long long*>(out + 1), (unsigned long long)__double_as_longlong(sa[0]));
  }
}

// packed pair index k = hi(hi-1)/2 + lo  (lo < hi)
__device__ __forceinline__ void pair_decode(int64_t k, int64_t& lo, int64_t& hi) {
  int64_t h = (int64_t)((1.0 + sqrt(1.0 + 8.0 * (double)k)) * 0.5);
  while (h * (h - 1) / 2 > k) --h;
  while ((h + 1) * h / 2 <= k) ++h;
  hi = h;
  lo = k - h * (h - 1) / 2;
}

// blockIdx.y enumerates the rows (first pair, decoded once per block), threads the columns with 32-bit arithmetic
__global__ void __launch_bounds__(EW_THREADS) pack_kernel(PackArgs p, int64_t rows, int64_t cols) {
  const uint32_t ncols = (uint32_t)cols, d3 = (uint32_t)p.d3;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    int64_t i0, i1;
    if (p.flags & 1) pair_decode(r, i0, i1);
    else { i0 = r / p.d1; i1 = r - i0 * p.d1; }
    const double* __restrict__ s = p.src + i0 * p.s0 + i1 * p.s1;
    const double* __restrict__ s2 = p.src + i1 * p.s0 + i0 * p.s1;
    double* drow = p.dst + r * p.ld;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < ncols; c += gridDim.x * blockDim.x) {
      int64_t i2, i3;
      if (p.flags & 2) pair_decode((int64_t)c, i2, i3);
      else { const uint32_t q = c / d3; i2 = q; i3 = c - q * d3; }
      double val = s[i2 * p.s2 + i3 * p.s3];
      if (p.flags & 4) val -= s[i3 * p.s2 + i2 * p.s3];
      if (p.flags & 8) {
        double w = s2[i2 * p.s2 + i3 * p.s3];
        if (p.flags & 4) w -= s2[i3 * p.s2 + i2 * p.s3];
        val -= w;
      }
      double out = p.alpha * val;
      if (p.beta != 0.0) out += p.beta * drow[c];
      drow[c] = out;
    }
  }
}

// blockIdx.y enumerates the first pair (i0,i1), threads the second pair (i2,i3) with 32-bit arithmetic, two
// consecutive i3 per thread and 16-byte accesses to the destination when its last axis is contiguous (VEC = 2)
template <int VEC>
__global__ void __launch_bounds__(EW_THREADS) unpack_kernel(PackArgs p) {
  const uint32_t d3 = (uint32_t)p.d3, n23 = (uint32_t)p.d2 * d3;
  for (int64_t row = blockIdx.y; row < p.d0 * p.d1; row += gridDim.y) {
  const int64_t i0 = row / p.d1, i1 = row - i0 * p.d1;
  double sign0 = 1.0;
  bool zero0 = false;
  int64_t r;
  if (p.flags & 1) {
    zero0 = i0 == i1;
    const int64_t lo = i0 < i1 ? i0 : i1, hi = i0 < i1 ? i1 : i0;
    if (i0 > i1) sign0 = -1.0;
    r = hi * (hi - 1) / 2 + lo;
  } else r = i0 * p.d1 + i1;
  const double* __restrict__ src = p.src + r * p.ld;
  double* dst = p.dst + i0 * p.s0 + i1 * p.s1;
  auto value = [&](uint32_t i2, uint32_t i3) {
    double sign = sign0;
    bool zero = zero0;
    int64_t c;
    if (p.flags & 2) {
      zero = zero || i2 == i3;
      const uint32_t lo = i2 < i3 ? i2 : i3, hi = i2 < i3 ? i3 : i2;
      if (i2 > i3) sign = -sign;
      c = (int64_t)hi * (hi - 1) / 2 + lo;
    } else c = (int64_t)i2 * d3 + i3;
    return zero ? 0.0 : p.alpha * (sign * src[c]);
  };
  for (uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) * VEC; e < n23; e += gridDim.x * blockDim.x * VEC) {
    const uint32_t i2 = e / d3, i3 = e - i2 * d3;
    double* d = dst + (int64_t)i2 * p.s2 + (int64_t)i3 * p.s3;
    if constexpr (VEC == 2) {
      double2 out = make_double2(value(i2, i3), value(i2, i3 + 1));
      if (p.beta != 0.0) {
        const double2 old = *reinterpret_cast<const double2*>(d);
        out.x += p.beta * old.x;
        out.y += p.beta * old.y;
      }
      *reinterpret_cast<double2*>(d) = out;
    } else {
      double out = value(i2, i3);
      if (p.beta != 0.0) out += p.beta * *d;
      *d = out;
    }
  }
  }
}

// utilities.subdiff (utilities.py:53-67): v > 0 -> e + alpha; otherwise soft threshold (Q1)
__device__ __forceinline__ double subdiff(double e, double v, double alpha) {
  if (v > 0.0) return e + alpha;
  if (e < -alpha) return e + alpha;
  if (e > alpha) return e - alpha;
  return 0.0;
}

__global__ void __launch_bounds__(EW_THREADS)
subdiff_kernel(const double* __restrict__ e, const double* __restrict__ v, double alpha, double* out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = subdiff(e[i], v[i], alpha);
}

// rank 4: blockIdx.y enumerates (i,j), threads (a,b) with 32-bit arithmetic; rank 2: one row of blocks
__global__ void __launch_bounds__(EW_THREADS)
finish_kernel(const double* r, const double* __restrict__ amp, const double* __restrict__ fock, int64_t ldf,
              double* out, int o, int v, int rank, int has_alpha, int equation, double alpha, double shift,
              int sub_singles) {
  extern __shared__ double eps[];
  const int n = o + v;
  for (int i = threadIdx.x; i < n; i += blockDim.x) eps[i] = fock[(int64_t)i * ldf + i];
  __syncthreads();
  const uint32_t uv = (uint32_t)v;
  const uint32_t per = rank == 2 ? (uint32_t)o * uv : uv * uv;
  const int64_t nrow = rank == 2 ? 1 : (int64_t)o * o;
  for (int64_t row = blockIdx.y; row < nrow; row += gridDim.y) {
    const int i = (int)(row / o), j = (int)(row - (int64_t)i * o);
    const int64_t base = row * (int64_t)per;
    for (uint32_t e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < per; e0 += gridDim.x * blockDim.x) {
      const uint32_t a = e0 / uv, b = e0 - a * uv;
      double d;
      if (rank == 2) d = shift + (eps[a] - eps[o + b]);            // here (a, b) = (i, a)
      else {
        d = (eps[i] - eps[o + a]) + (eps[j] - eps[o + b]);
        if (shift != 0.0) d += shift;
      }
      const int64_t idx = base + e0;
      double e = r[idx], res;
      if (has_alpha) {
        double t = amp[idx];
        double w = (rank == 4 || sub_singles) ? subdiff(e, t, alpha) : e;   // CCSD: L1 only on doubles (Q3)
        res = equation ? w : (w + t * d) / d;
      } else {
        res = equation ? e : e / d;
      }
      out[idx] = res;
    }
  }
}

__global__ void __launch_bounds__(EW_THREADS)
dot_partial_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* partial) {
  __shared__ double sh[EW_THREADS];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += a[i] * b[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = EW_THREADS / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// convergence vector of the solver loop (Solver_GS.py:598-612): conv = |a| + |b| (b == nullptr: conv = a) and the
// per-block partial sums of (conv - prev)^2 (prev == nullptr: skipped); fixed two-stage reduction order
__global__ void __launch_bounds__(EW_THREADS)
conv_partial_kernel(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ prev,
                    double* __restrict__ conv, int64_t n, double* partial) {
  __shared__ double sh[EW_THREADS];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double c = b ? fabs(a[i]) + fabs(b[i]) : a[i];
    if (prev) {
      const double d = c - prev[i];
      s += d * d;
    }
    conv[i] = c;
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = EW_THREADS / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// experimental potential of the density-matrix target (exp_pot.py:185-195): diff = target - rdm1,
// vexp = L diff, fsp = fock - vexp, stats = {sum |diff|, max |diff|}.  n x n is tiny: one block, fixed reduction order.
__global__ void __launch_bounds__(EW_THREADS)
vexp_mat_kernel(const double* __restrict__ rdm1, const double* __restrict__ target, const double* __restrict__ fock,
                double L, double* __restrict__ vexp, double* __restrict__ fsp, double* __restrict__ stats, int64_t n) {
  __shared__ double sh_sum[EW_THREADS];
  __shared__ double sh_max[EW_THREADS];
  double s = 0.0, m = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = target[i] - rdm1[i];
    const double w = L * d;
    vexp[i] = w;
    fsp[i] = fock[i] - w;
    s += fabs(d);
    m = fmax(m, fabs(d));
  }
  sh_sum[threadIdx.x] = s;
  sh_max[threadIdx.x] = m;
  __syncthreads();
  for (int w = EW_THREADS / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
      sh_sum[threadIdx.x] += sh_sum[threadIdx.x + w];
      sh_max[threadIdx.x] = fmax(sh_max[threadIdx.x], sh_max[threadIdx.x + w]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    stats[0] = sh_sum[0];
    stats[1] = sh_max[0];
  }
}

__global__ void __launch_bounds__(EW_THREADS)
dot_final_kernel(const double* partial, int nb, double* scal, double alpha, double beta) {
  __shared__ double sh[EW_THREADS];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = EW_THREADS / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *scal = (beta != 0.0 ? beta * *scal : 0.0) + alpha * sh[0];
}

__global__ void __launch_bounds__(EW_THREADS)
scale_dev_kernel(double* c, int64_t n, const double* scal, double d0, double d1) {
  const double f = d0 + d1 * *scal;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    c[i] *= f;
}

__global__ void diag_add_kernel(double* c, int64_t ldc, int64_t m, const double* fock, int64_t ldf, int64_t foff,
                                double alpha) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
    c[i * ldc + i] += alpha * fock[(foff + i) * ldf + foff + i];
}

// CCSD.py:154-160
__global__ void __launch_bounds__(EW_THREADS)
rdm1_kernel(const double* __restrict__ doo, const double* __restrict__ dvoT, const double* __restrict__ l1,
            const double* __restrict__ dvv, double* out, int o, int v) {
  const int n = o + v;
  const int64_t total = (int64_t)n * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int p = (int)(idx / n), q = (int)(idx - (int64_t)p * n);
    double val;
    if (p < o && q < o) {
      val = (doo[p * o + q] + doo[q * o + p]) * 0.5;
      if (p == q) val += 1.0;
    } else if (p < o) {
      val = (l1[p * v + (q - o)] + dvoT[p * v + (q - o)]) * 0.5;
    } else if (q < o) {
      val = (l1[q * v + (p - o)] + dvoT[q * v + (p - o)]) * 0.5;
    } else {
      int a = p - o, b = q - o;
      val = (dvv[(int64_t)a * v + b] + dvv[(int64_t)b * v + a]) * 0.5;
    }
    out[idx] = val;
  }
}

}  // namespace

cudaError_t launch_permute(const PermArgs& a, cudaStream_t st) {
  // squeeze unit dims and merge adjacent dims that are contiguous on both sides
  int64_t dim[KMAXD], si[KMAXD], so[KMAXD];
  int nd = 0;
  int64_t total = 1;
  for (int d = 0; d < a.nd; ++d) {
    total *= a.dim[d];
    if (a.dim[d] == 1) continue;
    if (nd > 0 && si[nd - 1] == a.sin[d] * a.dim[d] && so[nd - 1] == a.sout[d] * a.dim[d]) {
      dim[nd - 1] *= a.dim[d];
      si[nd - 1] = a.sin[d];
      so[nd - 1] = a.sout[d];
    } else {
      dim[nd] = a.dim[d]; si[nd] = a.sin[d]; so[nd] = a.sout[d];
      ++nd;
    }
  }
  if (total == 0) return cudaSuccess;
  if (nd == 0) { dim[0] = 1; si[0] = 1; so[0] = 1; nd = 1; }
  int fi = 0, fo = 0;
  for (int d = 1; d < nd; ++d) {
    if (si[d] < si[fi]) fi = d;
    if (so[d] < so[fo]) fo = d;
  }
  PermK k{};
  k.in = a.in; k.out = a.out; k.alpha = a.alpha; k.beta = a.beta; k.total = total;
  if (fi == fo) {
    int n = 0;
    for (int d = 0; d < nd; ++d)
      if (d != fo) { k.dim[n] = dim[d]; k.sin[n] = si[d]; k.sout[n] = so[d]; ++n; }
    k.dim[n] = dim[fo]; k.sin[n] = si[fo]; k.sout[n] = so[fo]; ++n;
    k.nd = n;
    const bool small = total + 4LL * 148 * 32 * EW_THREADS < (1LL << 31);
    auto even = [](int64_t x) { return (x & 1) == 0; };
    bool vec = k.sin[n - 1] == 1 && k.sout[n - 1] == 1 && even(k.dim[n - 1]) &&
               (reinterpret_cast<uintptr_t>(a.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
    for (int d = 0; d + 1 < n; ++d) vec = vec && even(k.sin[d]) && even(k.sout[d]);
    if (vec) { k.dim[n - 1] /= 2; k.sin[n - 1] = 2; k.sout[n - 1] = 2; }
    const unsigned grid = grid_for(total / (vec ? 2 : 1), EW_THREADS * 4);
    if (vec && small) permute_linear_kernel<uint32_t, 2><<<grid, EW_THREADS, 0, st>>>(k);
    else if (vec) permute_linear_kernel<int64_t, 2><<<grid, EW_THREADS, 0, st>>>(k);
    else if (small) permute_linear_kernel<uint32_t, 1><<<grid, EW_THREADS, 0, st>>>(k);
    else permute_linear_kernel<int64_t, 1><<<grid, EW_THREADS, 0, st>>>(k);
  } else {
    int n = 0;
    int64_t rest = 1;
    for (int d = 0; d < nd; ++d)
      if (d != fo && d != fi) { k.dim[n] = dim[d]; k.sin[n] = si[d]; k.sout[n] = so[d]; rest *= dim[d]; ++n; }
    k.nd = n;
    k.d_fi = dim[fi]; k.fi_sin = si[fi]; k.fi_sout = so[fi];
    k.d_fo = dim[fo]; k.fo_sin = si[fo]; k.fo_sout = so[fo];
    int64_t blocks = ((k.d_fi + 31) / 32) * ((k.d_fo + 31) / 32) * rest;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    permute_tiled_kernel<<<(unsigned)blocks, 256, 0, st>>>(k);
  }
  return cudaGetLastError();
}

cudaError_t launch_ew2(const Ew2Args& a, cudaStream_t st) {
  int64_t total = 1;
  for (int d = 0; d < a.nd; ++d) total *= a.dim[d];
  if (total <= 0) return cudaSuccess;
  ew2_kernel<<<grid_for(total, EW_THREADS * 2), EW_THREADS, 0, st>>>(a, total);
  return cudaGetLastError();
}

cudaError_t launch_fill(double* c, int64_t n, double value, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  fill_kernel<<<grid_for(n, EW_THREADS * 4), EW_THREADS, 0, st>>>(c, n, value);
  return cudaGetLastError();
}

cudaError_t launch_reduce(const double* part, int64_t nz, int64_t M, int64_t N, double* C, int64_t sr, int64_t sc,
                          double alpha, double beta, cudaStream_t st) {
  if (M * N <= 0) return cudaSuccess;
  reduce_kernel<<<grid_for(M * N, EW_THREADS), EW_THREADS, 0, st>>>(part, nz, M, N, C, sr, sc, alpha, beta);
  return cudaGetLastError();
}

// out[ijab] = c0 base[ijab] + z[ijab] - z[jiab] - z[ijba] + z[jiba] + beta out[ijab]   (P(ij)P(ab) z, CCSD.py:306-310,
// 475-477: the antisymmetriser of the doubles residuals).  One block per (pair i <= j, 32x32 tile pair A <= B of (a,b)):
// with D = z_ij - z_ji the four images are out_ij[A,B] = +D[A,B] - D[B,A]^T, out_ij[B,A] = -(that)^T, out_ji = -out_ij,
// so every element of z is read once and every element of out written once.
__global__ void __launch_bounds__(256) asym4_kernel(const double* __restrict__ base, double c0, const double* __restrict__ z,
                                                    double* __restrict__ out, double beta, int o, int v) {
  __shared__ double tAB[32][33], tBA[32][33];
  const int nt = (v + 31) / 32;
  // blockIdx.x -> tile pair (A <= B); blockIdx.y -> occupied pair (i <= j), grid-stride
  int tA = 0, rem = blockIdx.x;
  while (rem >= nt - tA) { rem -= nt - tA; ++tA; }
  const int tB = tA + rem;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
  const int64_t vv = (int64_t)v * v;
  const int64_t npair = (int64_t)o * (o + 1) / 2;
  for (int64_t pr = blockIdx.y; pr < npair; pr += gridDim.y) {
    int j = (int)((sqrt(8.0 * (double)pr + 1.0) - 1.0) * 0.5);
    while ((int64_t)(j + 1) * (j + 2) / 2 <= pr) ++j;
    while ((int64_t)j * (j + 1) / 2 > pr) --j;
    const int i = (int)(pr - (int64_t)j * (j + 1) / 2);             // i <= j
    const int64_t rij = ((int64_t)i * o + j) * vv, rji = ((int64_t)j * o + i) * vv;
    // D tiles: tAB[r][c] = D[A*32+r, B*32+c], tBA[r][c] = D[B*32+r, A*32+c]
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int a = tA * 32 + r, b = tB * 32 + tx;
      tAB[r][tx] = (a < v && b < v) ? z[rij + (int64_t)a * v + b] - z[rji + (int64_t)a * v + b] : 0.0;
      const int a2 = tB * 32 + r, b2 = tA * 32 + tx;
      tBA[r][tx] = (a2 < v && b2 < v) ? z[rij + (int64_t)a2 * v + b2] - z[rji + (int64_t)a2 * v + b2] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      {  // tile (A,B)
        const int a = tA * 32 + r, b = tB * 32 + tx;
        if (a < v && b < v) {
          const double d = (i == j) ? 0.0 : tAB[r][tx] - tBA[tx][r];
          const int64_t e1 = rij + (int64_t)a * v + b, e2 = rji + (int64_t)a * v + b;
          double w1 = d, w2 = -d;
          if (c0 != 0.0) { w1 += c0 * base[e1]; w2 += c0 * base[e2]; }
          if (beta != 0.0) { w1 += beta * out[e1]; w2 += beta * out[e2]; }
          out[e1] = w1;
          if (i != j) out[e2] = w2;
        }
      }
      if (tA != tB) {  // mirror tile (B,A): D[B,A] - D[A,B]^T
        const int a = tB * 32 + r, b = tA * 32 + tx;
        if (a < v && b < v) {
          const double d = (i == j) ? 0.0 : tBA[r][tx] - tAB[tx][r];
          const int64_t e1 = rij + (int64_t)a * v + b, e2 = rji + (int64_t)a * v + b;
          double w1 = d, w2 = -d;
          if (c0 != 0.0) { w1 += c0 * base[e1]; w2 += c0 * base[e2]; }
          if (beta != 0.0) { w1 += beta * out[e1]; w2 += beta * out[e2]; }
          out[e1] = w1;
          if (i != j) out[e2] = w2;
        }
      }
    }
    __syncthreads();
  }
}

cudaError_t launch_asym4(const double* base, double c0, const double* z, double* out, double beta, int o, int v,
                         cudaStream_t st) {
  if ((int64_t)o * o * v * v <= 0) return cudaSuccess;
  if (!base) c0 = 0.0;
  const int nt = (v + 31) / 32;
  const int64_t npair = (int64_t)o * (o + 1) / 2;
  dim3 grid((unsigned)(nt * (nt + 1) / 2), (unsigned)std::min<int64_t>(npair, 65535));
  asym4_kernel<<<grid, 256, 0, st>>>(base, c0, z, out, beta, o, v);
  return cudaGetLastError();
}

cudaError_t launch_tau(const double* t2, const double* t1, double* out, int o, int v, double c1, double c2,
                       cudaStream_t st, int i0, int ni) {
  if (ni < 0) ni = o - i0;
  int64_t total = (int64_t)ni * o * v * v;
  if (total <= 0) return cudaSuccess;
  const bool vec = (v % 2 == 0) && (reinterpret_cast<uintptr_t>(t2) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const int64_t per = (int64_t)v * v / (vec ? 2 : 1);
  dim3 grid((unsigned)std::min<int64_t>((per + EW_THREADS - 1) / EW_THREADS, 64),
            (unsigned)std::min<int64_t>((int64_t)ni * o, 65535));
  if (vec) tau_kernel<2><<<grid, EW_THREADS, 0, st>>>(t2, t1, out, o, v, c1, c2, i0, ni);
  else tau_kernel<1><<<grid, EW_THREADS, 0, st>>>(t2, t1, out, o, v, c1, c2, i0, ni);
  return cudaGetLastError();
}

cudaError_t launch_antisym_defect(const double* x, int o, int v, double* out, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, 2 * sizeof(double), st);
  if (e != cudaSuccess) return e;
  int64_t total = (int64_t)o * o * v * v;
  if (total <= 0) return cudaSuccess;
  dim3 grid((unsigned)std::min<int64_t>(((int64_t)v * v + EW_THREADS - 1) / EW_THREADS, 16),
            (unsigned)std::min<int64_t>((int64_t)o * o, 65535));
  defect_kernel<<<grid, EW_THREADS, 0, st>>>(x, o, v, out);
  return cudaGetLastError();
}

cudaError_t launch_pack(const PackArgs& a, cudaStream_t st) {
  const int64_t rows = (a.flags & 1) ? a.d0 * (a.d0 - 1) / 2 : a.d0 * a.d1;
  const int64_t cols = (a.flags & 2) ? a.d2 * (a.d2 - 1) / 2 : a.d2 * a.d3;
  if (rows * cols <= 0) return cudaSuccess;
  if (cols >= (1LL << 31)) return cudaErrorInvalidValue;
  dim3 grid((unsigned)std::min<int64_t>((cols + EW_THREADS - 1) / EW_THREADS, 64), (unsigned)std::min<int64_t>(rows, 65535));
  pack_kernel<<<grid, EW_THREADS, 0, st>>>(a, rows, cols);
  return cudaGetLastError();
}

cudaError_t launch_unpack(const PackArgs& a, cudaStream_t st) {
  const int64_t total = a.d0 * a.d1 * a.d2 * a.d3;
  if (total <= 0) return cudaSuccess;
  if (a.d2 * a.d3 >= (1LL << 31)) return cudaErrorInvalidValue;
  auto even = [](int64_t x) { return (x & 1) == 0; };
  const bool vec = a.s3 == 1 && even(a.d3) && even(a.s0) && even(a.s1) && even(a.s2) &&
                   (reinterpret_cast<uintptr_t>(a.dst) & 15) == 0;
  const int64_t per = a.d2 * a.d3 / (vec ? 2 : 1);
  dim3 grid((unsigned)std::min<int64_t>((per + EW_THREADS - 1) / EW_THREADS, 64),
            (unsigned)std::min<int64_t>(a.d0 * a.d1, 65535));
  if (vec) unpack_kernel<2><<<grid, EW_THREADS, 0, st>>>(a);
  else unpack_kernel<1><<<grid, EW_THREADS, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_finish(const double* r, const double* amp, const double* fock, int64_t ldf, double* out, int o,
                          int v, int rank, int has_alpha, int equation, double alpha, double shift, int sub_singles,
                          cudaStream_t st) {
  int64_t total = rank == 2 ? (int64_t)o * v : (int64_t)o * o * v * v;
  if (total <= 0) return cudaSuccess;
  size_t sm = sizeof(double) * (size_t)(o + v);
  const int64_t per = rank == 2 ? (int64_t)o * v : (int64_t)v * v, rows = rank == 2 ? 1 : (int64_t)o * o;
  if (per >= (1LL << 31)) return cudaErrorInvalidValue;
  dim3 grid((unsigned)std::min<int64_t>((per + EW_THREADS - 1) / EW_THREADS, rank == 2 ? 1024 : 64),
            (unsigned)std::min<int64_t>(rows, 65535));
  finish_kernel<<<grid, EW_THREADS, sm, st>>>(r, amp, fock, ldf, out, o, v, rank, has_alpha, equation, alpha, shift,
                                              sub_singles);
  return cudaGetLastError();
}

cudaError_t launch_subdiff(const double* e, const double* v, double alpha, double* out, int64_t n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  subdiff_kernel<<<grid_for(n, EW_THREADS * 4), EW_THREADS, 0, st>>>(e, v, alpha, out, n);
  return cudaGetLastError();
}

cudaError_t launch_conv(const double* a, const double* b, const double* prev, double* conv, int64_t n, double* partial,
                        int nblocks, double* scal, double beta, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  int nb = (int)std::min<int64_t>(nblocks, std::max<int64_t>(1, (n + EW_THREADS - 1) / EW_THREADS));
  conv_partial_kernel<<<nb, EW_THREADS, 0, st>>>(a, b, prev, conv, n, partial);
  dot_final_kernel<<<1, EW_THREADS, 0, st>>>(partial, nb, scal, 1.0, beta);
  return cudaGetLastError();
}

cudaError_t launch_vexp_mat(const double* rdm1, const double* target, const double* fock, double L, double* vexp,
                            double* fsp, double* stats, int64_t n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  vexp_mat_kernel<<<1, EW_THREADS, 0, st>>>(rdm1, target, fock, L, vexp, fsp, stats, n);
  return cudaGetLastError();
}

cudaError_t launch_dot(const double* a, const double* b, int64_t n, double* partial, int nblocks, double* scal,
                       double alpha, double beta, cudaStream_t st) {
  int nb = (int)std::min<int64_t>(nblocks, std::max<int64_t>(1, (n + EW_THREADS - 1) / EW_THREADS));
  dot_partial_kernel<<<nb, EW_THREADS, 0, st>>>(a, b, n, partial);
  dot_final_kernel<<<1, EW_THREADS, 0, st>>>(partial, nb, scal, alpha, beta);
  return cudaGetLastError();
}

cudaError_t launch_scale_dev(double* c, int64_t n, const double* scal, double d0, double d1, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  scale_dev_kernel<<<grid_for(n, EW_THREADS * 4), EW_THREADS, 0, st>>>(c, n, scal, d0, d1);
  return cudaGetLastError();
}

cudaError_t launch_diag_add(double* c, int64_t ldc, int64_t m, const double* fock, int64_t ldf, int64_t foff,
                            double alpha, cudaStream_t st) {
  if (m <= 0) return cudaSuccess;
  diag_add_kernel<<<grid_for(m, 128), 128, 0, st>>>(c, ldc, m, fock, ldf, foff, alpha);
  return cudaGetLastError();
}

cudaError_t launch_rdm1(const double* doo, const double* dvoT, const double* l1, const double* dvv, double* out,
                        int o, int v, cudaStream_t st) {
  int64_t total = (int64_t)(o + v) * (o + v);
  rdm1_kernel<<<grid_for(total, EW_THREADS), EW_THREADS, 0, st>>>(doo, dvoT, l1, dvv, out, o, v);
  return cudaGetLastError();
}

}  // namespace ecw
