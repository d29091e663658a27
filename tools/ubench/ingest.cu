// ingest.cu — per-SM ingest rate from L2 into shared memory on B200: bulk TMA (cp.async.bulk), cp.async (LDGSTS)
// and both at once.  Question behind it (csrc/ozaki.cu): the INT8 GEMM needs 47 B/clk/SM; is the ~32 B/clk/SM it
// gets a limit of the TMA unit or of the SM's L2 port?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ingest ingest.cu && ./ingest
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 1) ingest(const char* __restrict__ src, size_t region, int iters, int mode,
                                                 unsigned long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int CH = 4096, NCH = 32;                 // TMA ring: 4 stages of 8 chunks of 4 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NCH * CH + 32 * 2048);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < NCH; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t base = ((size_t)blockIdx.x * 7919 * CH) % region;
  const unsigned long long t0 = clock64();
  if ((mode & 1) && tid == 0) {
    // bulk TMA as the GEMM producer issues it: NST stages of PER chunks (4 KB each), one mbarrier per stage
    constexpr int PER = 8, NST = NCH / PER;
    const int nstage = iters / PER;
    for (int it = 0; it < nstage + NST; ++it) {
      const int slot = it % NST;
      if (it >= NST) {
        const uint32_t par = ((it / NST) - 1) & 1;
        asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra W; }" ::"r"(s32(&bar[slot])), "r"(par) : "memory");
      }
      if (it < nstage) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[slot])), "r"(CH * PER) : "memory");
#pragma unroll
        for (int c = 0; c < PER; ++c) {
          const size_t off = (base + ((size_t)it * PER + c) * CH) % region;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + (slot * PER + c) * CH)), "l"(src + off), "r"(CH), "r"(s32(&bar[slot])) : "memory");
        }
      }
    }
  }
  if ((mode & 2) && tid >= 128) {
    // cp.async 16 B per thread per step, 128 threads -> 2 KB per step, 4 steps per group, 8 groups in flight (64 KB)
    const int t = tid - 128;
    unsigned char* dst = smem + NCH * CH;
    const int steps = iters * 2;                     // same byte count as the TMA side (iters * 4 KB)
    for (int s = 0; s < steps; ++s) {
      const size_t off = (base + region / 2 + (size_t)s * 2048 + t * 16) % region;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst + (s % 32) * 2048 + t * 16)), "l"(src + off) : "memory");
      if ((s & 3) == 3) {
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 7;" ::: "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  const size_t region = 16ull << 20;                 // 16 MB: stays in L2 (each die caches its own copy)
  char* src;
  cudaMalloc(&src, region);
  cudaMemset(src, 1, region);
  unsigned long long* cyc;
  cudaMalloc(&cyc, 148 * sizeof(unsigned long long));
  const int smem = 32 * 4096 + 32 * 2048 + 1024;
  cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 40000;                           // 80 MB per CTA per path
  for (int grid : {1, 74, 148}) {
    for (int mode : {1, 2, 3}) {
      ingest<<<grid, 256, smem>>>(src, region, iters, mode, cyc);      // warm (L2)
      ingest<<<grid, 256, smem>>>(src, region, iters, mode, cyc);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      unsigned long long h[148];
      cudaMemcpy(h, cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int i = 0; i < grid; ++i) mx = mx > (double)h[i] ? mx : (double)h[i];
      const double bytes = (double)iters * 4096 * ((mode & 1 ? 1 : 0) + (mode & 2 ? 1 : 0));
      printf("grid %3d mode %d (%s): %.1f B/clk/SM (%.0f kclk)\n", grid, mode,
             mode == 1 ? "bulk TMA" : mode == 2 ? "cp.async" : "both", bytes / mx, mx / 1e3);
    }
  }
  return 0;
}
