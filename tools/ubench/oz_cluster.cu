// oz_cluster.cu — load pattern of the INT8 GEMM main loop (csrc/ozaki.cu, NS = 6, TN = 80: per stage 24.6 KB of A planes
// and 15.4 KB of B planes per CTA) with and without sharing the B panel inside a thread-block cluster:
//   mode 0: every CTA fetches its A stage and the whole B stage with bulk TMA (what the kernel does today);
//   mode 1: clusters of CS CTAs (same B panel, CS different A panels): each CTA fetches 1/CS of the B stage and
//           multicasts it to all CTAs of the cluster (cp.async.bulk ... .multicast::cluster).
// No MMA is issued: the question is only how many clocks a stage of loads takes when all SMs load (DESIGN.md §10.1;
// the GEMM needs <= 840 clk per stage to become MMA bound, today's loads take ~1224).
// Protocol as the GEMM would use it: per stage slot a `full` mbarrier (expect_tx = bytes landing in THIS CTA's shared
// memory, A + whole B) and an `empty` mbarrier that needs one arrival from the consumer of EVERY CTA of the cluster
// (a peer's multicast writes into my slot, so the slot is free only when all CS consumers are done with it).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o oz_cluster oz_cluster.cu && ./oz_cluster
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int A_BYTES = 128 * 32 * 6;   // 24576: 128 rows x 32 k x 6 digit planes
constexpr int B_BYTES = 80 * 32 * 6;    // 15360
constexpr int NST = 4;                  // stages in flight (the kernel has 4-5)
constexpr int SLOT = A_BYTES + B_BYTES; // 39936

__device__ __forceinline__ void wait_parity(uint32_t bar, uint32_t par) {
  asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra W; }" ::"r"(bar), "r"(par) : "memory");
}

template <int CS>
__global__ void __launch_bounds__(64, 1) stage_loads(const char* __restrict__ a, const char* __restrict__ b, size_t a_region,
                                                      size_t b_region, int nstage, int mode, unsigned long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NST * SLOT);
  uint64_t* empty = full + NST;
  uint32_t rank = 0;
  if (CS > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[i])), "r"(mode ? CS : 1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CS > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  const size_t a0 = ((size_t)blockIdx.x * 7919 * 4096) % a_region;
  const size_t b0 = ((size_t)(blockIdx.x / CS) * 104729 * 512) % b_region;     // one B panel per cluster
  const unsigned long long t0 = clock64();
  if (threadIdx.x == 0) {                                                       // producer
    for (int it = 0; it < nstage; ++it) {
      const int slot = it % NST;
      if (it >= NST) wait_parity(s32(&empty[slot]), ((it / NST) - 1) & 1);
      unsigned char* dst = smem + slot * SLOT;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[slot])), "r"(SLOT) : "memory");
      const size_t ao = (a0 + (size_t)it * A_BYTES) % (a_region - A_BYTES);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(s32(dst)), "l"(a + (ao & ~(size_t)15)), "r"(A_BYTES), "r"(s32(&full[slot])) : "memory");
      const size_t bo = ((b0 + (size_t)it * B_BYTES) % (b_region - B_BYTES)) & ~(size_t)15;
      if (mode == 0) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(dst + A_BYTES)), "l"(b + bo), "r"(B_BYTES), "r"(s32(&full[slot])) : "memory");
      } else {
        constexpr int PART = B_BYTES / CS;                                      // 1920 B for CS = 8
        const uint16_t mask = (uint16_t)((1u << CS) - 1);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                     ::"r"(s32(dst + A_BYTES + rank * PART)), "l"(b + bo + rank * PART), "r"(PART), "r"(s32(&full[slot])), "h"(mask)
                     : "memory");
      }
    }
  } else if (threadIdx.x == 32) {                                               // consumer: wait, then free the slot
    for (int it = 0; it < nstage; ++it) {
      const int slot = it % NST;
      wait_parity(s32(&full[slot]), (it / NST) & 1);
      if (mode == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[slot])) : "memory");
      } else {
#pragma unroll
        for (int c = 0; c < CS; ++c) {                                          // tell every CTA of the cluster
          uint32_t remote;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(s32(&empty[slot])), "r"(c));
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  if (CS > 1) {                                                                 // nobody exits while peers may still write here
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

template <int CS>
static double run(const char* a, const char* b, size_t ar, size_t br, int grid, int nstage, int mode, unsigned long long* cyc) {
  const int smem = NST * SLOT + 1024;
  cudaFuncSetAttribute(stage_loads<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, stage_loads<CS>, a, b, ar, br, nstage, mode, cyc);
    if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
      return -1;
    }
  }
  unsigned long long h[148];
  cudaMemcpy(h, cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < grid; ++i) mx = mx > (double)h[i] ? mx : (double)h[i];
  return mx / nstage;
}

int main() {
  const size_t ar = 1ull << 30, br = 64ull << 20;    // A streams from DRAM/L2 like the vvvv planes; B panels stay in L2
  char *a, *b;
  cudaMalloc(&a, ar);
  cudaMalloc(&b, br);
  cudaMemset(a, 1, ar);
  cudaMemset(b, 1, br);
  unsigned long long* cyc;
  cudaMalloc(&cyc, 148 * sizeof(unsigned long long));
  const int nstage = 20000;
  printf("clocks per stage (A %d B + B %d B), max over CTAs; the MMAs of a stage take 840 clk\n", A_BYTES, B_BYTES);
  printf("  grid 148, no cluster, unicast B        : %.0f\n", run<1>(a, b, ar, br, 148, nstage, 0, cyc));
  printf("  grid 144, cluster 2,  multicast B      : %.0f\n", run<2>(a, b, ar, br, 144, nstage, 1, cyc));
  printf("  grid 144, cluster 4,  multicast B      : %.0f\n", run<4>(a, b, ar, br, 144, nstage, 1, cyc));
  printf("  grid 144, cluster 8,  unicast B        : %.0f\n", run<8>(a, b, ar, br, 144, nstage, 0, cyc));
  printf("  grid 144, cluster 8,  multicast B      : %.0f\n", run<8>(a, b, ar, br, 144, nstage, 1, cyc));
  return 0;
}
