// Microbenchmark: issue throughput of FFMA, packed FFMA2/FADD2, FMNMX3 and the epilogue mix on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[16];
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2*i]) << 32) | __float_as_uint(a[2*i+1]);
  float s = seed * 0.5f, mx0 = -1e30f, mx1 = -1e30f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {          // 16 independent scalar FFMA
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, a[i]);
    } else if (MODE == 1) {   // 8 independent packed FFMA2 (= 16 flops-lanes)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(p[i]) : "l"(p[(i+1)&7]));
    } else if (MODE == 2) {   // 8 packed FADD2
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(p[(i+1)&7]));
    } else if (MODE == 3) {   // 8 FMNMX3 (2 chains)
#pragma unroll
      for (int i = 0; i < 16; i += 2) { if (i & 2) mx0 = fmaxf(mx0, fmaxf(a[i], a[i+1])); else mx1 = fmaxf(mx1, fmaxf(a[i], a[i+1])); a[i] += 0.f; }
    } else if (MODE == 4) {   // epilogue mix per 16 values: 16 FFMA(abs) + 8 FADD2 + 8 FFMA2 + 8 FMNMX3
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float y0 = a[i], y1 = a[i+1];
        if (i & 2) mx0 = fmaxf(mx0, fmaxf(y0, y1)); else mx1 = fmaxf(mx1, fmaxf(y0, y1));
        float t0 = fmaf(s, y0, fabsf(y0)), t1 = fmaf(s, y1, fabsf(y1));
        asm volatile("{\n.reg .b64 tp;\nmov.b64 tp, {%2, %3};\nadd.rn.f32x2 %0, %0, tp;\nfma.rn.f32x2 %1, tp, tp, %1;\n}\n"
            : "+l"(p[(i>>1)&3]), "+l"(p[4+((i>>1)&3)]) : "f"(t0), "f"(t1));
      }
    } else if (MODE == 5) {   // scalar-only variant of the mix: 16 FFMA(abs) + 16 FADD + 16 FFMA + 8 FMNMX3
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float y0 = a[i], y1 = a[i+1];
        if (i & 2) mx0 = fmaxf(mx0, fmaxf(y0, y1)); else mx1 = fmaxf(mx1, fmaxf(y0, y1));
        float t0 = fmaf(s, y0, fabsf(y0)), t1 = fmaf(s, y1, fabsf(y1));
        float* f = reinterpret_cast<float*>(p);
        f[i & 7] += t0; f[(i + 1) & 7] += t1;
        f[8 + (i & 7)] = fmaf(t0, t0, f[8 + (i & 7)]); f[8 + ((i + 1) & 7)] = fmaf(t1, t1, f[8 + ((i + 1) & 7)]);
      }
    }
  }
  long long t1 = clock64();
  float r = s + mx0 + mx1;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int warps, int per_iter) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  k<MODE><<<148, warps * 32>>>(out, cyc, 1.0f); k<MODE><<<148, warps * 32>>>(out, cyc, 1.0f);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps/SM %2d: %.2f cycles per warp-instr-group of %d (%.3f cyc/instr/SMSP)\n", name, warps, (double)h / ITERS, per_iter,
         (double)h / ITERS / per_iter * (warps >= 4 ? 1.0 : 1.0) / ((warps + 3) / 4));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16}) {
    if (w == 4) { run<0>("16 FFMA", 4, 16); run<1>("8 FFMA2", 4, 8); run<2>("8 FADD2", 4, 8); run<3>("8 FMNMX3(+8 FADD)", 4, 16); run<4>("mix packed (40 instr)", 4, 40); run<5>("mix scalar (56 instr)", 4, 56); }
    if (w == 8) { run<0>("16 FFMA", 8, 16); run<1>("8 FFMA2", 8, 8); run<2>("8 FADD2", 8, 8); run<3>("8 FMNMX3(+8 FADD)", 8, 16); run<4>("mix packed (40 instr)", 8, 40); run<5>("mix scalar (56 instr)", 8, 56); }
    if (w == 16) { run<0>("16 FFMA", 16, 16); run<1>("8 FFMA2", 16, 8); run<2>("8 FADD2", 16, 8); run<3>("8 FMNMX3(+8 FADD)", 16, 16); run<4>("mix packed (40 instr)", 16, 40); run<5>("mix scalar (56 instr)", 16, 56); }
  }
  return 0;
}
