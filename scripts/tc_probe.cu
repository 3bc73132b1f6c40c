// Standalone probe: one tcgen05.mma (kind::tf32, M=64, N, K=8), dump all of TMEM to see the layout.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;
  return d;
}

__global__ void probe(float* out, int N, int sbo_b, int variant, int lane_off) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  float* A = (float*)smem;                 // 2048 B
  unsigned char* Bt = smem + 2048;
  const int G4 = N / 4;
  const int lbo_b = G4 * sbo_b;
  // A[m][k]: K-major no swizzle
  for (int idx = threadIdx.x; idx < 64 * 8; idx += blockDim.x) {
    int m = idx / 8, k = idx % 8;
    float v = (k == 0) ? (float)(m + 1) : ((k == 1) ? 0.5f : 0.f);
    if (variant >= 2 && variant <= 5) v = 1.0f;
    *(float*)((unsigned char*)A + (m >> 3) * 256 + (k >> 2) * 128 + (m & 7) * 16 + (k & 3) * 4) = v;
  }
  // B[k][n]: MN-major no swizzle: chunk (k, g) of 4 slots at g*sbo + k*16
  for (int idx = threadIdx.x; idx < 8 * N; idx += blockDim.x) {
    int k = idx / N, n = idx % N;
    float v = (k == 0) ? (float)(n + 1) : ((k == 1) ? 1000.f : 0.f);
    if (variant >= 2 && variant <= 5) v = 1.0f;
    *(float*)(Bt + (n >> 2) * sbo_b + k * 16 + (n & 3) * 4) = v;
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
  {
    // pre-fill TMEM columns [0,N) of every lane with 7.0 through tcgen05.st
    const int warp0 = threadIdx.x >> 5;
    for (int col = 0; col < N; col += 8) {
      uint32_t seven = __float_as_uint(7.0f);
      uint32_t ta = tb + ((uint32_t)(32 * warp0) << 16) + col;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(ta), "r"(seven) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (threadIdx.x == 0) out[128 * N] = __uint_as_float(tb);
  if (threadIdx.x == 0 && variant != 9) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
    if (variant == 3) idesc &= ~(1u << 16);            // B K-major
    if (variant == 4) idesc = (idesc & ~((7u << 7) | (7u << 10))) | (0u << 7) | (0u << 10);   // claim f16 formats
    if (variant == 1) idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint64_t ad = smem_desc(smem_u32(A), 128, 256);
    uint64_t bd = smem_desc(smem_u32(Bt), lbo_b, sbo_b);
    uint32_t d = tb + ((uint32_t)lane_off << 16);
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMA
  if (variant != 9) {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
  }
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
  int N = argc > 1 ? atoi(argv[1]) : 16;
  int sbo = argc > 2 ? atoi(argv[2]) : 128;
  int variant = argc > 3 ? atoi(argv[3]) : 0;
  int lane_off = argc > 4 ? atoi(argv[4]) : 0;
  float* d; CK(cudaMalloc(&d, 128 * N * 4 + 32)); CK(cudaMemset(d, 0, 128 * N * 4 + 32));
  size_t smem = 2048 + (N / 4) * sbo + 256;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe<<<1, 128, smem>>>(d, N, sbo, variant, lane_off);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  float* h = (float*)malloc(128 * N * 4 + 32); CK(cudaMemcpy(h, d, 128 * N * 4 + 32, cudaMemcpyDeviceToHost)); { unsigned* u = (unsigned*)&h[128 * N]; printf("tmem base=0x%08x idesc=0x%08x adesc=0x%08x_%08x bdesc=0x%08x_%08x\n", u[0], u[1], u[3], u[2], u[5], u[4]); }
  printf("N=%d sbo=%d variant=%d lane_off=%d  expect D[m][n] = (m+1)(n+1) + 500\n", N, sbo, variant, lane_off);
  for (int l = 0; l < 128; ++l) {
    bool any = false; for (int c = 0; c < N; ++c) any |= h[l * N + c] != 7.f;
    if (!any) continue;
    if (l > 3) continue;
    printf("lane %3d:", l); for (int c = 0; c < (N < 12 ? N : 12); ++c) printf(" %8.1f", h[l * N + c]); printf("\n");
  }
  return 0;
}
