/* synth_c.c — TEST INFRASTRUCTURE ONLY.  C restatement of oracle/synth.py (function-defined synthetic integrals and
 * amplitudes: splitmix64 of the index tuple) so that the sampled-element oracle at the benchmark shape (40,400),
 * tests/test_gpu_bench_shape.py, gets its o v^3 / v^3 slices in seconds instead of minutes.  Bit-identical to
 * oracle/synth.py (tests/test_oracle_columns_cpu.py); compiled by oracle/synth_fast.py with `gcc -O2 -fopenmp`.
 * Block definitions: Eris.py:128-150 of the reference (antisymmetrised <pq||rs>). */
#include <stdint.h>

static inline uint64_t splitmix64(uint64_t key, uint64_t seed) {
  uint64_t x = key + seed * 0x9E3779B97F4A7C15ull;
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static inline double unit(uint64_t z) { return (double)(z >> 11) * 0x1p-52 - 1.0; }

static inline double eri(int64_t n, double scale, int64_t p, int64_t q, int64_t r, int64_t s) {
  if (p == q || r == s) return 0.0;
  double sgn = ((p < q) ? 1.0 : -1.0) * ((r < s) ? 1.0 : -1.0);
  int64_t bra = (p < q ? p : q) * n + (p < q ? q : p);
  int64_t ket = (r < s ? r : s) * n + (r < s ? s : r);
  int64_t lo = bra < ket ? bra : ket, hi = bra < ket ? ket : bra;
  return sgn * (scale * unit(splitmix64((uint64_t)(lo * (n * n) + hi), 1)));
}

/* out[ip, iq, ir, is] = <ps[ip] qs[iq] || rs[ir] ss[is]>  (absolute spin-orbital indices) */
void ecw_oracle_eri_block(int64_t n, double scale, const int64_t* ps, int64_t np, const int64_t* qs, int64_t nq,
                          const int64_t* rs, int64_t nr, const int64_t* ss, int64_t ns, double* out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t a = 0; a < np; ++a)
    for (int64_t b = 0; b < nq; ++b) {
      double* o = out + (a * nq + b) * nr * ns;
      for (int64_t c = 0; c < nr; ++c)
        for (int64_t d = 0; d < ns; ++d) o[c * ns + d] = eri(n, scale, ps[a], qs[b], rs[c], ss[d]);
    }
}

/* doubles amplitudes t2 / l2 (oracle/synth.py:doubles) */
void ecw_oracle_doubles(int64_t o, int64_t v, uint64_t seed, double scale, double* out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t i = 0; i < o; ++i)
    for (int64_t j = 0; j < o; ++j) {
      double* dst = out + (i * o + j) * v * v;
      for (int64_t a = 0; a < v; ++a)
        for (int64_t b = 0; b < v; ++b) {
          double val = 0.0;
          if (i != j && a != b) {
            int64_t li = i < j ? i : j, hi = i < j ? j : i, la = a < b ? a : b, ha = a < b ? b : a;
            uint64_t z = splitmix64((uint64_t)(((li * o + hi) * v + la) * v + ha), seed);
            double sgn = ((i < j) ? 1.0 : -1.0) * ((a < b) ? 1.0 : -1.0);
            val = (z & 7ull) == 0 ? 0.0 : sgn * (scale * unit(z));
          }
          dst[a * v + b] = val;
        }
    }
}
