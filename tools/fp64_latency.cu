// tools/fp64_latency.cu -- dependent-issue latency and per-scheduler throughput of the FP64 pipe (DFMA/DMUL/DADD), and how
// throughput scales with warps per scheduler x independent chains per warp. One block per SM, WARPS warps.
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(int iters, double x, double y, double* out, long long* cyc) {
  double a[CHAINS];
#pragma unroll
  for (int j = 0; j < CHAINS; j++) a[j] = 1.0 + j + threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++)
#pragma unroll
      for (int j = 0; j < CHAINS; j++) a[j] = __fma_rn(a[j], x, y);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; j++) s += a[j];
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CHAINS>
void run(int warps) {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<CHAINS><<<148, warps * 32>>>(iters, 1.0000001, 1e-9, out, cyc);
  k<CHAINS><<<148, warps * 32>>>(iters, 1.0000001, 1e-9, out, cyc);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double inst_per_warp = (double)iters * 16 * CHAINS;
  double per_sched_warps = warps / 4.0;
  printf("warps/SM %2d chains %d: %.2f cycles per dependent DFMA step, %.3f DFMA/cycle/scheduler\n", warps, CHAINS,
         (double)c / (iters * 16.0), inst_per_warp * per_sched_warps / c);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1>(4); run<2>(4); run<4>(4); run<8>(4);
  run<1>(8); run<1>(16); run<2>(16); run<4>(16); run<1>(32); run<2>(32); run<1>(64);
  return 0;
}
