// Microbenchmark: streaming-write patterns for the [B,64,600,600] canvas (369 MB).
#include <cstdio>
#include <cuda_runtime.h>
#define HW 360000
#define C 64
#define B 4
__global__ void k_linear(float4* o, size_t n4) {
  const float4 z = make_float4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) __stcs(o + i, z);
}
// warp owns CELLS consecutive cells for all channels (CELLS/128 float4 per lane per channel)
template <int CELLS, int CS>
__global__ void __launch_bounds__(256) k_warp(float* canvas) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ngroups = HW / CELLS;
  float* cb = canvas + (size_t)b * C * HW;
  const float4 z = make_float4(0, 0, 0, 0);
  for (int grp = blockIdx.x * 8 + warp; grp < ngroups; grp += gridDim.x * 8) {
    float4* o = reinterpret_cast<float4*>(cb + (size_t)grp * CELLS) + lane;
#pragma unroll 8
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int j = 0; j < CELLS / 128; ++j) {
        if (CS) __stcs(o + (size_t)c * (HW / 4) + j * 32, z); else o[(size_t)c * (HW / 4) + j * 32] = z;
      }
  }
}
// CTA owns (channel, chunk of 4096 cells): long contiguous runs per plane
__global__ void __launch_bounds__(256) k_plane(float* canvas) {
  const float4 z = make_float4(0, 0, 0, 0);
  const int chunks = (HW + 4095) / 4096;
  const int total = B * C * chunks;
  for (int u = blockIdx.x; u < total; u += gridDim.x) {
    const int chunk = u % chunks, plane = u / chunks;
    float4* o = reinterpret_cast<float4*>(canvas + (size_t)plane * HW + (size_t)chunk * 4096);
    const int n4 = min(4096, HW - chunk * 4096) / 4;
    for (int i = threadIdx.x; i < n4; i += 256) __stcs(o + i, z);
  }
}
template <class F> void timeit(const char* name, F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a); for (int i = 0; i < 10; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("%-40s %.1f us  %.0f GB/s  (%s)\n", name, ms * 100, (double)B * C * HW * 4 / (ms / 10 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float* canvas; size_t bytes = (size_t)B * C * HW * 4; cudaMalloc(&canvas, bytes);
  float* other; cudaMalloc(&other, (size_t)1 << 30);   // L2 flush target not used; canvas >> L2
  timeit("cudaMemsetAsync", [&] { cudaMemsetAsync(canvas, 0, bytes); });
  timeit("linear float4 grid-stride (148*8 CTAs)", [&] { k_linear<<<148 * 8, 256>>>((float4*)canvas, bytes / 16); });
  timeit("linear float4, one pass (n/256 CTAs)", [&] { k_linear<<<(unsigned)(bytes / 16 / 256), 256>>>((float4*)canvas, bytes / 16); });
  timeit("warp 128 cells x 64 planes, .cs", [&] { k_warp<128, 1><<<dim3(352, B), 256>>>(canvas); });
  timeit("warp 128 cells x 64 planes, default st", [&] { k_warp<128, 0><<<dim3(352, B), 256>>>(canvas); });
  timeit("warp 256 cells x 64 planes, .cs", [&] { k_warp<256, 1><<<dim3(176, B), 256>>>(canvas); });
  timeit("warp 512 cells x 64 planes, .cs", [&] { k_warp<512, 1><<<dim3(88, B), 256>>>(canvas); });
  timeit("CTA per (plane, 4096-cell chunk), .cs", [&] { k_plane<<<148 * 8, 256>>>(canvas); });
  timeit("CTA per (plane, 4096-cell chunk) 148*16", [&] { k_plane<<<148 * 16, 256>>>(canvas); });
  return 0;
}
