// Microbenchmark: warp-level mma.sync (legacy tensor path) issue rate on sm_100a, tf32 m16n8k8 and bf16 m16n8k16,
// alone and interleaved with FFMA, per SM sub-partition (threads/SM = 128 -> one warp per sub-partition).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int MODE, int NCH>
__global__ void k(float* out, long long* cyc, unsigned seed) {
  float c[NCH][4];
  float f[8];
  unsigned a[4] = {seed, seed + 1, seed + 2, seed + 3}, b[2] = {seed + 4, seed + 5};
#pragma unroll
  for (int i = 0; i < NCH; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = threadIdx.x + i;
  const float s = 1.0001f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (MODE == 0 || MODE == 2) mma_tf32(c[i], a, b); else mma_bf16(c[i], a, b);
      if (MODE >= 2) {
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] = fmaf(f[q], s, s);       // 8 FFMA per MMA
      }
    }
  }
  long long t1 = clock64();
  float r = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) for (int j = 0; j < 4; ++j) r += c[i][j];
#pragma unroll
  for (int i = 0; i < 8; ++i) r += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE, int NCH> void run(const char* name, int threads) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  k<MODE, NCH><<<148, threads>>>(out, cyc, 1u); k<MODE, NCH><<<148, threads>>>(out, cyc, 1u);
  long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / ITERS / NCH;
  printf("%-40s %4d thr/SM, %d chains: %.2f cycles per MMA per warp; %.2f cycles per MMA per sub-partition\n", name, threads, NCH, per,
         per / (threads / 128.0));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 1>("tf32 m16n8k8 dependent", 128); run<0, 8>("tf32 m16n8k8", 128); run<0, 8>("tf32 m16n8k8", 512); run<0, 8>("tf32 m16n8k8", 1024);
  run<1, 1>("bf16 m16n8k16 dependent", 128); run<1, 8>("bf16 m16n8k16", 128); run<1, 8>("bf16 m16n8k16", 512); run<1, 8>("bf16 m16n8k16", 1024);
  run<2, 8>("tf32 m16n8k8 + 8 FFMA each", 512); run<3, 8>("bf16 m16n8k16 + 8 FFMA each", 512);
  cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
  return 0;
}
