// Microbenchmark: FP64 add/mul/fma latency (dependent chain) and throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k(double* out, long long* cyc, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x;
  const double s = seed * 0.999;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) { a[0] = __dadd_rn(a[0], s); }                                  // 1 dependent DADD
    else if (MODE == 1) { a[0] = __dadd_rn(__dmul_rn(a[0], s), s); }               // DMUL -> DADD chain
    else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = __dadd_rn(a[i], s);                       // 8 independent DADD
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], s, s);                          // 8 independent DFMA
    } else if (MODE == 4) { a[0] = a[0] / s; }                                     // 1 dependent division
  }
  long long t1 = clock64();
  double r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int threads, int ops) {
  double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  k<MODE><<<148, threads>>>(out, cyc, 1.0); k<MODE><<<148, threads>>>(out, cyc, 1.0);
  long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %4d threads/SM: %.1f cycles per iteration (%d fp64 ops per thread per iteration)\n", name, threads, (double)h / ITERS, ops);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("1 dependent DADD", 32, 1); run<1>("DMUL->DADD chain", 32, 2); run<4>("1 dependent division", 32, 1);
  run<2>("8 independent DADD", 32, 8); run<2>("8 independent DADD", 128, 8); run<2>("8 independent DADD", 512, 8); run<2>("8 independent DADD", 1024, 8);
  run<3>("8 independent DFMA", 128, 8); run<3>("8 independent DFMA", 1024, 8);
  return 0;
}
