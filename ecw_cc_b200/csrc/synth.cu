// synth.cu — function-defined synthetic inputs generated directly in their
// device layouts (packed vvvv never exists densely).  Bit-identical to the
// numpy definition in oracle/synth.py; the formula is specified in DESIGN.md
// "Synthetic inputs".  Integral symmetries follow Eris.py:128.
#include "kernels.h"

namespace ecw {

namespace {

constexpr unsigned long long GOLDEN = 0x9E3779B97F4A7C15ULL;

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long key, unsigned long long seed) {
  unsigned long long x = key + seed * GOLDEN;
  x += GOLDEN;
  unsigned long long z = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

__device__ __forceinline__ double unit(unsigned long long z) {
  return __dadd_rn(__dmul_rn((double)(z >> 11), 0x1.0p-52), -1.0);
}

// <pq||rs>, absolute spin-orbital indices
__device__ __forceinline__ double eri(long long n, long long p, long long q, long long r, long long s, double scale) {
  if (p == q || r == s) return 0.0;
  double sgn = ((p < q) ? 1.0 : -1.0) * ((r < s) ? 1.0 : -1.0);
  long long bra = (p < q ? p : q) * n + (p < q ? q : p);
  long long ket = (r < s ? r : s) * n + (r < s ? s : r);
  long long lo = bra < ket ? bra : ket, hi = bra < ket ? ket : bra;
  unsigned long long key = (unsigned long long)(lo * (n * n) + hi);
  return sgn * __dmul_rn(scale, unit(splitmix64(key, 1ULL)));
}

__device__ __forceinline__ void pair_decode(long long k, long long& lo, long long& hi) {
  long long h = (long long)((1.0 + sqrt(1.0 + 8.0 * (double)k)) * 0.5);
  while (h * (h - 1) / 2 > k) --h;
  while ((h + 1) * h / 2 <= k) ++h;
  hi = h;
  lo = k - h * (h - 1) / 2;
}

__device__ __forceinline__ double eps(int o, int v, long long p) {
  if (p < o) return -2.0 + __ddiv_rn(__dmul_rn(1.5, (double)p), (double)(o > 1 ? o - 1 : 1));
  return 0.5 + __ddiv_rn(__dmul_rn(2.5, (double)(p - o)), (double)(v > 1 ? v - 1 : 1));
}

__device__ __forceinline__ double amp2(int o, int v, long long i, long long j, long long a, long long b,
                                       unsigned long long seed) {
  if (i == j || a == b) return 0.0;
  long long li = i < j ? i : j, hi = i < j ? j : i, la = a < b ? a : b, ha = a < b ? b : a;
  unsigned long long key = (unsigned long long)(((li * o + hi) * v + la) * v + ha);
  unsigned long long z = splitmix64(key, seed);
  if ((z & 7ULL) == 0ULL) return 0.0;
  double sgn = ((i < j) ? 1.0 : -1.0) * ((a < b) ? 1.0 : -1.0);
  return sgn * __dmul_rn(0.02, unit(z));
}

__global__ void __launch_bounds__(256)
synth_kernel(int kind, double* out, int o, int v, long long row0, long long nrows, double scale) {
  const long long n = o + v, po = (long long)o * (o - 1) / 2, pv = (long long)v * (v - 1) / 2;
  long long cols;
  switch (kind) {
    case SY_OOOO: cols = (long long)o * o * o; break;
    case SY_OOOV: cols = (long long)o * o * v; break;
    case SY_OOVV: cols = (long long)o * v * v; break;
    case SY_OOVV_PH: case SY_OVOV_PH: cols = (long long)v * o * v; break;
    case SY_OVVV: cols = (long long)v * v * v; break;
    case SY_OOOO_P: cols = po; break;
    case SY_OOVV_P: cols = pv; break;
    case SY_OVVV_P: cols = (long long)v * pv; break;
    case SY_VVVV_P: cols = pv; break;
    case SY_FOCK: case SY_FSP: cols = n; break;
    case SY_T1: case SY_L1: cols = v; break;
    default: cols = (long long)o * v * v; break;   // T2 / L2
  }
  const long long total = nrows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long r = row0 + idx / cols, c = idx % cols;
    double val = 0.0;
    long long x, y, z, w, lo, hi, lo2, hi2;
    switch (kind) {
      case SY_OOOO: x = c / ((long long)o * o); y = (c / o) % o; z = c % o; val = eri(n, r, x, y, z, scale); break;
      case SY_OOOV: x = c / ((long long)o * v); y = (c / v) % o; z = c % v; val = eri(n, r, x, y, o + z, scale); break;
      case SY_OOVV: x = c / ((long long)v * v); y = (c / v) % v; z = c % v; val = eri(n, r, x, o + y, o + z, scale); break;
      case SY_OOVV_PH:   // [(m,e),(n,f)] = <mn||ef>
        x = c / ((long long)o * v); y = (c / v) % o; z = c % v; val = eri(n, r, y, o + x, o + z, scale); break;
      case SY_OVOV_PH:   // [(i,a),(n,f)] = <na||if>
        x = c / ((long long)o * v); y = (c / v) % o; z = c % v; val = eri(n, y, o + x, r, o + z, scale); break;
      case SY_OVVV: x = c / ((long long)v * v); y = (c / v) % v; z = c % v; val = eri(n, r, o + x, o + y, o + z, scale); break;
      case SY_OOOO_P: pair_decode(r, lo, hi); pair_decode(c, lo2, hi2); val = eri(n, lo, hi, lo2, hi2, scale); break;
      case SY_OOVV_P: pair_decode(r, lo, hi); pair_decode(c, lo2, hi2); val = eri(n, lo, hi, o + lo2, o + hi2, scale); break;
      case SY_OVVV_P: x = c / pv; w = c % pv; pair_decode(w, lo2, hi2); val = eri(n, r, o + x, o + lo2, o + hi2, scale); break;
      case SY_VVVV_P: pair_decode(r, lo, hi); pair_decode(c, lo2, hi2); val = eri(n, o + lo, o + hi, o + lo2, o + hi2, scale); break;
      case SY_FOCK: val = (r == c) ? eps(o, v, r) : 0.0; break;
      case SY_FSP: {
        long long a = r < c ? r : c, b = r < c ? c : r;
        double vx = __dmul_rn(0.05, unit(splitmix64((unsigned long long)(a * n + b), 2ULL)));
        val = ((r == c) ? eps(o, v, r) : 0.0) - vx;
        break;
      }
      case SY_T1: val = __dmul_rn(0.05, unit(splitmix64((unsigned long long)(r * v + c), 3ULL))); break;
      case SY_L1: val = __dmul_rn(0.05, unit(splitmix64((unsigned long long)(r * v + c), 4ULL))); break;
      case SY_T2: x = c / ((long long)v * v); y = (c / v) % v; z = c % v; val = amp2(o, v, r, x, y, z, 5ULL); break;
      case SY_L2: x = c / ((long long)v * v); y = (c / v) % v; z = c % v; val = amp2(o, v, r, x, y, z, 6ULL); break;
    }
    out[idx] = val;
  }
}

// oovv_ph[(m,e),(n,f)] = oovv[m,n,e,f];  ovov_ph[(i,a),(n,f)] = ovov[n,a,i,f]
__global__ void __launch_bounds__(256)
ph_layout_kernel(const double* __restrict__ oovv, const double* __restrict__ ovov, double* oovv_ph, double* ovov_ph,
                 int o, int v) {
  const long long total = (long long)o * v * o * v;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long f = idx % v, nn = (idx / v) % o, e = (idx / ((long long)v * o)) % v, m = idx / ((long long)v * o * v);
    oovv_ph[idx] = oovv[((m * o + nn) * v + e) * v + f];
    ovov_ph[idx] = ovov[((nn * v + e) * o + m) * v + f];
  }
}

}  // namespace

cudaError_t launch_synth(int kind, double* out, int o, int v, int64_t row0, int64_t nrows, double scale,
                         cudaStream_t st) {
  if (nrows <= 0) return cudaSuccess;
  synth_kernel<<<148 * 16, 256, 0, st>>>(kind, out, o, v, (long long)row0, (long long)nrows, scale);
  return cudaGetLastError();
}

cudaError_t launch_eris_layouts_from_dense(const double* oovv, const double* ovov, double* oovv_ph, double* ovov_ph,
                                           int o, int v, cudaStream_t st) {
  ph_layout_kernel<<<148 * 16, 256, 0, st>>>(oovv, ovov, oovv_ph, ovov_ph, o, v);
  return cudaGetLastError();
}

}  // namespace ecw
