// Microbenchmark: TMEM-load + reduce epilogue in isolation (no MMA, no barriers): cycles per 200-column pass.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define ITERS 200
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void acc_pair(unsigned long long& S, unsigned long long& Q, float t0, float t1) {
  asm("{\n.reg .b64 tp;\nmov.b64 tp, {%2, %3};\nadd.rn.f32x2 %0, %0, tp;\nfma.rn.f32x2 %1, tp, tp, %1;\n}\n" : "+l"(S), "+l"(Q) : "f"(t0), "f"(t1));
}
#define LD16(taddr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
  : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr))
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: full train epilogue; 1: loads + max only; 2: arithmetic only on stale registers (no loads)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int ncols, float sgn_in) {
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = holder;
  const int q = warp & 3, k4 = warp >> 2;           // 4 sets of 4 warps
  const int e = k4 >> 1, j = k4 & 1;
  const int n0 = j ? 104 : 0, n1 = j ? ncols : 104;
  const uint32_t taddr = base + ((uint32_t)(32 * q) << 16) + e * 256 + n0;
  const int n16 = (n1 - n0) >> 4;
  const float sgn = sgn_in;
  float mx[2] = {-1e30f, -1e30f};
  unsigned long long S[4] = {0, 0, 0, 0}, Q[4] = {0, 0, 0, 0};
  uint32_t va[16], vb[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { va[i] = __float_as_uint(1.0f + i + lane); vb[i] = __float_as_uint(2.0f + i); }
  auto consume = [&](uint32_t* v) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) mx[(i >> 1) & 1] = fmaxf(mx[(i >> 1) & 1], fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
    if (MODE != 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float y = __uint_as_float(v[i]); v[i] = __float_as_uint(fmaf(sgn, y, fabsf(y))); }
#pragma unroll
      for (int i = 0; i < 16; i += 2) acc_pair(S[(i >> 1) & 3], Q[(i >> 1) & 3], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
    }
  };
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 2) {
      for (int kk = 0; kk < n16; ++kk) { consume(va); asm volatile("" : "+r"(va[0]), "+r"(va[1])); }
    } else {
      if (n16 > 0) LD16(taddr, va);
      for (int kk = 0; kk < n16; kk += 2) {
        ld_wait();
        if (kk + 1 < n16) LD16(taddr + 16 * (kk + 1), vb);
        consume(va);
        if (kk + 1 < n16) {
          ld_wait();
          if (kk + 2 < n16) LD16(taddr + 16 * (kk + 2), va);
          consume(vb);
        }
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  float r = mx[0] + mx[1];
  for (int i = 0; i < 4; ++i) r += __uint_as_float((unsigned)S[i]) + __uint_as_float((unsigned)(Q[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}
template <int MODE> void run(const char* name) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  k<MODE><<<148, 512>>>(out, cyc, 200, 1.0f); k<MODE><<<148, 512>>>(out, cyc, 200, 1.0f);
  long long h = 0; cudaError_t e = cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-40s %s: %.1f cycles per iteration (16 warps: 2 pairs x 200 columns)\n", name, cudaGetErrorString(e), (double)h / ITERS);
  cudaFree(out); cudaFree(cyc);
}
int main() { run<0>("loads + max + relu sums (train)"); run<1>("loads + max (eval)"); run<2>("arithmetic only, no TMEM loads"); return 0; }
