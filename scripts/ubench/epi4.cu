// Microbenchmark (round 2): what bounds the tensor-core epilogue of k_pfn_stats_tc / k_pfn_pad_tc?
//   * pure tcgen05.ld throughput per SM sub-partition (x32 / x16 / x64 shapes, 1..4 warps per sub-partition)
//   * tcgen05.ld + the arithmetic mixes under consideration
// Prints cycles per warp-value (32 lanes x 1 column) per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define ITERS 400
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void acc_pair(unsigned long long& S, unsigned long long& Q, float t0, float t1) {
  asm("{\n.reg .b64 tp;\nmov.b64 tp, {%2, %3};\nadd.rn.f32x2 %0, %0, tp;\nfma.rn.f32x2 %1, tp, tp, %1;\n}\n" : "+l"(S), "+l"(Q) : "f"(t0), "f"(t1));
}
#define R4(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3])
#define LD8(taddr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : R4(v, 0), R4(v, 4) : "r"(taddr))
#define LD16(taddr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
  : R4(v, 0), R4(v, 4), R4(v, 8), R4(v, 12) : "r"(taddr))
#define LD32(taddr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
  : R4(v, 0), R4(v, 4), R4(v, 8), R4(v, 12), R4(v, 16), R4(v, 20), R4(v, 24), R4(v, 28) : "r"(taddr))
#define LD64(taddr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31," \
  "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];" \
  : R4(v, 0), R4(v, 4), R4(v, 8), R4(v, 12), R4(v, 16), R4(v, 20), R4(v, 24), R4(v, 28), R4(v, 32), R4(v, 36), R4(v, 40), R4(v, 44), R4(v, 48), R4(v, 52), R4(v, 56), R4(v, 60) : "r"(taddr))
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MIX: 0 none (one register touched per load), 1 max only (FMNMX3), 2 current train mix (FFMA relu2 + packed add/fma + max),
//      3 abs mix: S += |v| (FADD), Q = fma(v, |v|, Q) (FFMA), max (sum v and sum v^2 come from the input moments),
//      4 abs mix without max, 5 current train mix without max
template <int MIX, int NV>
__device__ __forceinline__ void consume(uint32_t* v, float sgn, float* mx, unsigned long long* S, unsigned long long* Q, float* s1, float* q1) {
  if (MIX == 0) { mx[0] = fmaxf(mx[0], __uint_as_float(v[0])); return; }
  if (MIX == 1 || MIX == 2 || MIX == 3) {
#pragma unroll
    for (int i = 0; i < NV; i += 2) mx[(i >> 1) & 1] = fmaxf(mx[(i >> 1) & 1], fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
  }
  if (MIX == 2 || MIX == 5) {
#pragma unroll
    for (int i = 0; i < NV; ++i) { const float y = __uint_as_float(v[i]); v[i] = __float_as_uint(fmaf(sgn, y, fabsf(y))); }
#pragma unroll
    for (int i = 0; i < NV; i += 2) acc_pair(S[(i >> 1) & 3], Q[(i >> 1) & 3], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
  }
  if (MIX == 6 || MIX == 7) {   // integer max on the bit patterns (VIMNMX3) + abs mix
    int* im = reinterpret_cast<int*>(mx);
#pragma unroll
    for (int i = 0; i < NV; i += 2) im[(i >> 1) & 1] = max(im[(i >> 1) & 1], max((int)v[i], (int)v[i + 1]));
  }
  if (MIX == 8) {   // 2-input FMNMX, one per value
#pragma unroll
    for (int i = 0; i < NV; ++i) mx[i & 1] = fmaxf(mx[i & 1], __uint_as_float(v[i]));
  }
  if (MIX == 9) {   // unsigned min + signed max (mixed-sign rows)
    int* im = reinterpret_cast<int*>(mx);
    unsigned* um = reinterpret_cast<unsigned*>(mx) + 1;
#pragma unroll
    for (int i = 0; i < NV; i += 2) { im[0] = max(im[0], max((int)v[i], (int)v[i + 1])); um[0] = min(um[0], min(v[i], v[i + 1])); }
  }
  if (MIX == 3 || MIX == 4 || MIX == 6 || MIX == 9) {
#pragma unroll
    for (int i = 0; i < NV; ++i) { const float y = __uint_as_float(v[i]); s1[i & 7] += fabsf(y); q1[i & 7] = fmaf(y, fabsf(y), q1[i & 7]); }
  }
}

// SHAPE: 8/16/32/64 columns per tcgen05.ld; DB: 0 = load, wait, consume; 1 = software-pipelined (next load issued before consuming)
template <int MIX, int SHAPE, int DB, int WARPS>
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
  const uint32_t taddr = base + ((uint32_t)(32 * q) << 16) + (k4 & 3) * 128;
  const float sgn = sgn_in;
  float mx[2] = {-1e30f, -1e30f};
  unsigned long long S[4] = {0, 0, 0, 0}, Q[4] = {0, 0, 0, 0};
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (DB == 0) {
      uint32_t v[SHAPE];
      for (int c0 = 0; c0 < cols; c0 += SHAPE) {
        if (SHAPE == 8) LD8(taddr + c0, v); else if (SHAPE == 16) LD16(taddr + c0, v); else if (SHAPE == 32) LD32(taddr + c0, v); else LD64(taddr + c0, v);
        ld_wait();
        consume<MIX, SHAPE>(v, sgn, mx, S, Q, s1, q1);
      }
    } else {
      uint32_t va[SHAPE], vb[SHAPE];
      if (SHAPE == 16) LD16(taddr, va); else LD32(taddr, va);
      for (int c0 = 0; c0 < cols; c0 += 2 * SHAPE) {
        ld_wait();
        if (SHAPE == 16) LD16(taddr + c0 + SHAPE, vb); else LD32(taddr + c0 + SHAPE, vb);
        consume<MIX, SHAPE>(va, sgn, mx, S, Q, s1, q1);
        ld_wait();
        if (c0 + 2 * SHAPE < cols) { if (SHAPE == 16) LD16(taddr + c0 + 2 * SHAPE, va); else LD32(taddr + c0 + 2 * SHAPE, va); }
        consume<MIX, SHAPE>(vb, sgn, mx, S, Q, s1, q1);
      }
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
template <int MIX, int SHAPE, int DB, int WARPS> void run(const char* name) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int cols = 128;
  k<MIX, SHAPE, DB, WARPS><<<148, WARPS * 32>>>(out, cyc, cols, 1.0f); k<MIX, SHAPE, DB, WARPS><<<148, WARPS * 32>>>(out, cyc, cols, 1.0f);
  long long h = 0; cudaError_t e = cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s x%-2d %s %d warps/SMSP %s: %.2f cycles per warp-value per SMSP\n", name, SHAPE, DB ? "pipelined" : "ld-wait  ", WARPS / 4,
         e == cudaSuccess ? "" : cudaGetErrorString(e), (double)h / ITERS / cols / (WARPS / 4));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 32, 0, 4>("load only"); run<0, 32, 0, 8>("load only"); run<0, 32, 0, 16>("load only");
  run<0, 16, 0, 16>("load only"); run<0, 64, 0, 16>("load only"); run<0, 64, 0, 8>("load only"); run<0, 8, 0, 16>("load only");
  run<0, 32, 1, 4>("load only"); run<0, 32, 1, 16>("load only");
  run<1, 32, 0, 16>("max only"); run<1, 32, 0, 8>("max only"); run<1, 64, 0, 8>("max only"); run<1, 32, 1, 8>("max only"); run<1, 32, 1, 16>("max only");
  run<2, 32, 0, 16>("train mix (current)"); run<2, 32, 0, 8>("train mix (current)"); run<2, 32, 1, 8>("train mix (current)"); run<2, 16, 1, 16>("train mix (current)");
  run<2, 64, 0, 8>("train mix (current)");
  run<5, 32, 0, 16>("train mix, no max");
  run<3, 32, 0, 16>("abs mix"); run<3, 32, 0, 8>("abs mix"); run<3, 32, 1, 8>("abs mix"); run<3, 64, 0, 8>("abs mix"); run<3, 16, 1, 16>("abs mix");
  run<4, 32, 0, 16>("abs mix, no max");
  run<6, 32, 0, 16>("abs mix + int max3"); run<6, 32, 0, 8>("abs mix + int max3"); run<6, 64, 0, 8>("abs mix + int max3");
  run<7, 32, 0, 16>("int max3 only"); run<8, 32, 0, 16>("fmnmx 2-input only");
  run<9, 32, 0, 16>("abs mix + int max3 + uint min3");
  return 0;
}
