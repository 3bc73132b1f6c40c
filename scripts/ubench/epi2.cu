// Microbenchmark: variants of the epilogue arithmetic (no TMEM), 16 warps/SM, 16 values per step.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define ITERS 4000
__device__ __forceinline__ void acc_pair(unsigned long long& S, unsigned long long& Q, float t0, float t1) {
  asm("{\n.reg .b64 tp;\nmov.b64 tp, {%2, %3};\nadd.rn.f32x2 %0, %0, tp;\nfma.rn.f32x2 %1, tp, tp, %1;\n}\n" : "+l"(S), "+l"(Q) : "f"(t0), "f"(t1));
}
template <int V>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, float sgn_in, float decay) {
  const int lane = threadIdx.x & 31;
  const float sgn = sgn_in;
  float mx[2] = {-1e30f, -1e30f};
  unsigned long long S[4] = {0, 0, 0, 0}, Q[4] = {0, 0, 0, 0};
  float s1[8] = {0}, q1[8] = {0};
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 1.0f + i + lane;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    float t[16];
    if (V != 3) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) mx[(i >> 1) & 1] = fmaxf(mx[(i >> 1) & 1], fmaxf(v[i], v[i + 1]));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = fmaf(sgn, v[i], fabsf(v[i]));
    if (V == 0 || V == 3) {            // packed sums, 4+4 chains
#pragma unroll
      for (int i = 0; i < 16; i += 2) acc_pair(S[(i >> 1) & 3], Q[(i >> 1) & 3], t[i], t[i + 1]);
    } else if (V == 1) {               // scalar sums, 8+8 chains
#pragma unroll
      for (int i = 0; i < 16; ++i) { s1[i & 7] += t[i]; q1[i & 7] = fmaf(t[i], t[i], q1[i & 7]); }
    } else if (V == 2) {               // packed sum, scalar squares
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        asm("{\n.reg .b64 tp;\nmov.b64 tp, {%1, %2};\nadd.rn.f32x2 %0, %0, tp;\n}\n" : "+l"(S[(i >> 1) & 3]) : "f"(t[i]), "f"(t[i + 1]));
        q1[i & 7] = fmaf(t[i], t[i], q1[i & 7]); q1[(i + 1) & 7] = fmaf(t[i + 1], t[i + 1], q1[(i + 1) & 7]);
      }
    }
    // keep v alive and changing without extra FP work in the measured mix: rotate registers
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                     "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]));
  }
  const long long t1 = clock64();
  float r = mx[0] + mx[1];
  for (int i = 0; i < 4; ++i) r += __uint_as_float((unsigned)S[i]) + __uint_as_float((unsigned)(Q[i] >> 32)) + __uint_as_float((unsigned)(S[i] >> 32)) + __uint_as_float((unsigned)Q[i]);
  for (int i = 0; i < 8; ++i) r += s1[i] + q1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int V> void run(const char* name, int threads) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  k<V><<<148, threads>>>(out, cyc, 1.0f, 0.5f); k<V><<<148, threads>>>(out, cyc, 1.0f, 0.5f);
  long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %2d warps/SMSP: %.2f cycles per warp-value per SMSP\n", name, threads / 128, (double)h / ITERS / 16.0 / (threads / 128));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int th : {128, 256, 512}) {
    if (th == 128) { run<0>("max3 + relu-FFMA + packed S,Q", 128); run<1>("max3 + relu-FFMA + scalar S,Q", 128); run<2>("max3 + relu-FFMA + packed S, scalar Q", 128); run<3>("relu-FFMA + packed S,Q (no max)", 128); }
    if (th == 256) { run<0>("max3 + relu-FFMA + packed S,Q", 256); run<1>("max3 + relu-FFMA + scalar S,Q", 256); run<2>("max3 + relu-FFMA + packed S, scalar Q", 256); run<3>("relu-FFMA + packed S,Q (no max)", 256); }
    if (th == 512) { run<0>("max3 + relu-FFMA + packed S,Q", 512); run<1>("max3 + relu-FFMA + scalar S,Q", 512); run<2>("max3 + relu-FFMA + packed S, scalar Q", 512); run<3>("relu-FFMA + packed S,Q (no max)", 512); }
  }
  return 0;
}
