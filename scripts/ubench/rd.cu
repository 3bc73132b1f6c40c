// rd.cu -- pure-read HBM ceiling: LDG.128 grid-stride vs cp.async.bulk into shared-memory stages.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench/rd scripts/ubench/rd.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k_ldg(const uint4* __restrict__ p, size_t n, unsigned* out) {
  unsigned acc = 0;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * st < n; i += 4 * st) {
    uint4 a = p[i], b = p[i + st], c = p[i + 2 * st], d = p[i + 3 * st];
    acc += a.x ^ b.y ^ c.z ^ d.w;
  }
  for (; i < n; i += st) acc += p[i].x;
  if (acc == 0x12345678u) *out = acc;
}
// one producer lane, one consumer lane; chunk-sized stages
__global__ void __launch_bounds__(64, 1) k_bulk(const unsigned char* __restrict__ p, size_t total, int chunk, int stages, int split) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = (uint64_t*)sm;
  uint64_t* empty = full + 16;
  unsigned char* buf = sm + 256;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const size_t nch = total / chunk;
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) != 0) return;
  int s = 0; uint32_t ph = 0;
  for (size_t c = blockIdx.x; c < nch; c += gridDim.x) {
    if (warp == 0) {
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\nselp.u32 %0,1,0,q;\n}" : "=r"(ok) : "r"(s32(&empty[s])), "r"(ph ^ 1u) : "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(chunk) : "memory");
      const int part = chunk / split;
      for (int k = 0; k < split; ++k)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(buf + (size_t)s * chunk + k * part)), "l"(p + c * chunk + (size_t)k * part), "r"(part), "r"(s32(&full[s])) : "memory");
    } else {
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\nselp.u32 %0,1,0,q;\n}" : "=r"(ok) : "r"(s32(&full[s])), "r"(ph) : "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
    }
    if (++s == stages) { s = 0; ph ^= 1u; }
  }
}
int main() {
  const size_t total = (size_t)24000 * 200 * 48;   // the prepared operand of the padding pass
  unsigned char *d, *fl; unsigned* out;
  cudaMalloc(&d, total); cudaMalloc(&fl, 400u << 20); cudaMalloc(&out, 4);
  cudaMemset(d, 1, total);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto launch) {
    float best = 1e9f, sum = 0;
    for (int r = 0; r < 6; ++r) {
      cudaMemsetAsync(fl, r, 400u << 20);
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r) { sum += ms; if (ms < best) best = ms; }
    }
    printf("%-44s avg %7.2f us best %7.2f us  %6.0f GB/s (%s)\n", name, sum / 5 * 1e3, best * 1e3, total / (sum / 5 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  for (int bpsm : {4, 8, 16}) {
    char nm[64]; snprintf(nm, 64, "ldg128 x4, 256 thr, %d blocks/SM", bpsm);
    timeit(nm, [&] { k_ldg<<<148 * bpsm, 256>>>((const uint4*)d, total / 16, out); });
  }
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int chunk, stages, split, grid; };
  for (Cfg c : {Cfg{19200, 8, 1, 148}, Cfg{19200, 10, 1, 148}, Cfg{19200, 8, 6, 148}, Cfg{9600, 16, 1, 148}, Cfg{38400, 5, 1, 148},
                Cfg{19200, 4, 1, 148}, Cfg{19200, 2, 1, 148}, Cfg{4800, 16, 1, 148}, Cfg{19200, 8, 2, 148}}) {
    char nm[64]; snprintf(nm, 64, "bulk chunk %d stages %d split %d grid %d", c.chunk, c.stages, c.split, c.grid);
    timeit(nm, [&] { k_bulk<<<c.grid, 64, 256 + c.chunk * c.stages>>>(d, total, c.chunk, c.stages, c.split); });
  }
  return 0;
}
