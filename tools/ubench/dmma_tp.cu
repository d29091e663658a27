// DMMA.8x8x4 issue-throughput microbenchmark: TFLOP/s vs warps per SM and independent accumulators per warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k(double* out, int iters, double a0, double b0) {
  double acc[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = a0 + threadIdx.x, b = b0 + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run(int warps_per_sm, double* out) {
  int iters = 20000;
  dim3 grid(148), block(warps_per_sm * 32);
  k<NACC><<<grid, block>>>(out, 100, 1.0, 2.0);
  cudaDeviceSynchronize();
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  cudaEventRecord(s);
  k<NACC><<<grid, block>>>(out, iters, 1.0, 2.0);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  double fl = 148.0 * warps_per_sm * (double)iters * NACC * 512.0;
  printf("warps/SM %2d  acc/warp %2d : %7.2f TFLOP/s\n", warps_per_sm, NACC, fl / ms / 1e9);
}
int main() {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  for (int w : {4, 8, 16, 32}) { run<4>(w, out); run<8>(w, out); run<16>(w, out); run<32>(w, out); }
  return 0;
}
