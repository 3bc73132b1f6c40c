// Probe 2: A = all ones (64x8, K-major). B region = 16 KB filled by pattern; descriptor fields from argv.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// pattern: 0 = all ones; 1 = value is (16B-chunk index within region) for element 0 of chunk, 0 otherwise
//          2 = value = element index within chunk + 1 (1..4) only in chunk `sel`, zero elsewhere
__global__ void probe(float* out, int N, int lbo, int sbo, int bmajor, int pattern, int sel, int layout) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  float* A = (float*)smem;                 // 2048 B all ones
  unsigned char* Bt = smem + 2048;         // 16 KB
  for (int i = threadIdx.x; i < 512; i += blockDim.x) A[i] = 1.0f;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
    int chunk = i / 4, e = i % 4;
    float v = 1.0f;
    if (pattern == 1) v = (e == 0) ? (float)chunk : 0.f;
    if (pattern == 2) v = (chunk == sel) ? (float)(e + 1) : 0.f;
    ((float*)Bt)[i] = v;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = holder;
  if (threadIdx.x == 0) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | ((uint32_t)bmajor << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
    uint64_t ad = smem_desc(smem_u32(A), 128, 256, 0);
    uint64_t bd = smem_desc(smem_u32(Bt), lbo, sbo, layout);
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  { uint32_t ok = 0; while (!ok) { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory"); } }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int col = 0; col < N; col += 8) {
    uint32_t v[8];
    uint32_t ta = tb + ((uint32_t)(32 * warp) << 16) + col;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(ta));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[(32 * warp + lane) * N + col + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory");
}
int main(int argc, char** argv) {
  int N = atoi(argv[1]), lbo = atoi(argv[2]), sbo = atoi(argv[3]), bmajor = atoi(argv[4]), pattern = atoi(argv[5]);
  int sel = argc > 6 ? atoi(argv[6]) : 0, layout = argc > 7 ? atoi(argv[7]) : 0;
  float* d; CK(cudaMalloc(&d, 128 * N * 4)); CK(cudaMemset(d, 0, 128 * N * 4));
  size_t smem = 2048 + 16384 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe<<<1, 128, smem>>>(d, N, lbo, sbo, bmajor, pattern, sel, layout);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  float* h = (float*)malloc(128 * N * 4); CK(cudaMemcpy(h, d, 128 * N * 4, cudaMemcpyDeviceToHost));
  printf("N=%d lbo=%d sbo=%d bmajor=%d pattern=%d sel=%d layout=%d : lane0 =", N, lbo, sbo, bmajor, pattern, sel, layout);
  for (int c = 0; c < N; ++c) printf(" %g", h[c]);
  printf("\n");
  return 0;
}
