// Microbenchmark: epilogue structure variants WITH TMEM loads. 16 warps/SM (4 per SMSP), each warp reduces
// COLS columns of its lane quarter per iteration.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define ITERS 300
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void acc_pair(unsigned long long& S, unsigned long long& Q, float t0, float t1) {
  asm("{\n.reg .b64 tp;\nmov.b64 tp, {%2, %3};\nadd.rn.f32x2 %0, %0, tp;\nfma.rn.f32x2 %1, tp, tp, %1;\n}\n" : "+l"(S), "+l"(Q) : "f"(t0), "f"(t1));
}
#define LD16(taddr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
  : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr))
#define LD32(taddr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
  : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
    "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr))
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// V: 0 = x16 double-buffered in-place packed (current kernel), 1 = x32 single buffer packed, 2 = x16 single buffer packed,
//    3 = x32 single buffer scalar sums, 4 = x16 double-buffered, loads only + max (eval), 5 = x32 single, max only
template <int V, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(float* out, long long* cyc, int cols, float sgn_in) {
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = holder;
  const int q = warp & 3, k4 = warp >> 2;
  const uint32_t taddr = base + ((uint32_t)(32 * q) << 16) + (k4 & 3) * 96;
  const float sgn = sgn_in;
  float mx[2] = {-1e30f, -1e30f};
  unsigned long long S[4] = {0, 0, 0, 0}, Q[4] = {0, 0, 0, 0};
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto consume = [&](uint32_t* v, const int n, const bool train, const bool packed) {
#pragma unroll
    for (int i = 0; i < n; i += 2) mx[(i >> 1) & 1] = fmaxf(mx[(i >> 1) & 1], fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
    if (train) {
#pragma unroll
      for (int i = 0; i < n; ++i) { const float y = __uint_as_float(v[i]); v[i] = __float_as_uint(fmaf(sgn, y, fabsf(y))); }
      if (packed) {
#pragma unroll
        for (int i = 0; i < n; i += 2) acc_pair(S[(i >> 1) & 3], Q[(i >> 1) & 3], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      } else {
#pragma unroll
        for (int i = 0; i < n; ++i) { const float t = __uint_as_float(v[i]); s1[i & 7] += t; q1[i & 7] = fmaf(t, t, q1[i & 7]); }
      }
    }
  };
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (V == 0 || V == 4) {
      uint32_t va[16], vb[16];
      const int n16 = cols >> 4;
      LD16(taddr, va);
      for (int kk = 0; kk < n16; kk += 2) {
        ld_wait();
        if (kk + 1 < n16) LD16(taddr + 16 * (kk + 1), vb);
        consume(va, 16, V == 0, true);
        if (kk + 1 < n16) {
          ld_wait();
          if (kk + 2 < n16) LD16(taddr + 16 * (kk + 2), va);
          consume(vb, 16, V == 0, true);
        }
      }
    } else if (V == 1 || V == 3 || V == 5) {
      uint32_t v[32];
      for (int c0 = 0; c0 < cols; c0 += 32) { LD32(taddr + c0, v); ld_wait(); consume(v, 32, V != 5, V == 1); }
    } else if (V == 2) {
      uint32_t v[16];
      for (int c0 = 0; c0 < cols; c0 += 16) { LD16(taddr + c0, v); ld_wait(); consume(v, 16, true, true); }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  float r = mx[0] + mx[1];
  for (int i = 0; i < 4; ++i) r += __uint_as_float((unsigned)S[i]) + __uint_as_float((unsigned)(Q[i] >> 32));
  for (int i = 0; i < 8; ++i) r += s1[i] + q1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}
template <int V, int WARPS> void run(const char* name) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int cols = 96;
  k<V, WARPS><<<148, WARPS * 32>>>(out, cyc, cols, 1.0f); k<V, WARPS><<<148, WARPS * 32>>>(out, cyc, cols, 1.0f);
  long long h = 0; cudaError_t e = cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-52s %2d warps/SMSP %s: %.2f cycles per warp-value per SMSP\n", name, WARPS / 4, e == cudaSuccess ? "" : cudaGetErrorString(e), (double)h / ITERS / cols / (WARPS / 4));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 16>("x16 double-buffered, packed (current)"); run<1, 16>("x32 single buffer, packed"); run<2, 16>("x16 single buffer, packed");
  run<3, 16>("x32 single buffer, scalar sums"); run<4, 16>("x16 double-buffered, max only"); run<5, 16>("x32 single, max only");
  run<0, 8>("x16 double-buffered, packed (current)"); run<1, 8>("x32 single buffer, packed"); run<3, 8>("x32 single buffer, scalar sums"); run<5, 8>("x32 single, max only");
  run<1, 4>("x32 single buffer, packed"); run<3, 4>("x32 single buffer, scalar sums");
  return 0;
}
