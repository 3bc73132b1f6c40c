// tc_common.cuh -- PTX wrappers shared by the tcgen05 PFN kernels (mbarrier, bulk/tensor TMA, TMEM,
// UMMA descriptors) for sm_100a.
#pragma once
#include "common.cuh"

namespace pp {
namespace tcx {

constexpr unsigned long long kSpinLimit = 4000000000ull;   // ~2 s at 1.9 GHz: trap instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// try_wait with an explicit suspend-time hint: the thread sleeps in hardware until the phase completes
// or the hint (ns) expires.  Without the hint try_wait came back after ~18 cycles and the 16 waiting
// epilogue warps burnt more issue slots polling than the whole kernel spends computing (ncu r1j:
// 19.7 M TRYWAIT, 426 M instructions executed vs 173 M useful).
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  // a protocol bug traps after ~2 s instead of wedging the device.  The watchdog costs nothing that
  // shows in the kernel time (measured against a bare three-instruction retry loop: no difference).
  const unsigned long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, 100000u)) {
    if ((++spins & 255u) == 0u && clock64() - t0 > kSpinLimit) __trap();
  }
}
// Latency-critical hand-offs: poll with plain try_wait.  A try_wait with a suspend-time hint parks the thread and
// it wakes ~0.5 us after the phase completes (measured: k_pfn_pad_tc with neither MMAs nor epilogue arithmetic
// took 56 us for 81 barrier round trips per CTA), which serialised every TMEM hand-off.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 4095u) == 0u && clock64() - t0 > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_spin_t(uint64_t* bar, uint32_t parity, bool on, long long& acc) {
  if (!on) { mbar_wait_spin(bar, parity); return; }
  const long long t = clock64();
  mbar_wait_spin(bar, parity);
  acc += clock64() - t;
}
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, bool on, long long& acc) {
  if (!on) { mbar_wait(bar, parity); return; }
  const long long t = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// one lane of the (converged) warp, chosen by the hardware
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// Blackwell packed fp32 pairs (add/fma .f32x2): S += (t0,t1), Q += (t0*t0, t1*t1) in two instructions
__device__ __forceinline__ void acc_pair(unsigned long long& S, unsigned long long& Q, float t0, float t1) {
  asm("{\n.reg .b64 tp;\nmov.b64 tp, {%2, %3};\nadd.rn.f32x2 %0, %0, tp;\nfma.rn.f32x2 %1, tp, tp, %1;\n}\n"
      : "+l"(S), "+l"(Q) : "f"(t0), "f"(t1));
}
// weighted form: S += (u0,u1), Q += (u0*t0, u1*t1)
__device__ __forceinline__ void acc_pair_w(unsigned long long& S, unsigned long long& Q, float u0, float u1, float t0, float t1) {
  asm("{\n.reg .b64 up, tp;\nmov.b64 up, {%2, %3};\nmov.b64 tp, {%4, %5};\nadd.rn.f32x2 %0, %0, up;\nfma.rn.f32x2 %1, up, tp, %1;\n}\n"
      : "+l"(S), "+l"(Q) : "f"(u0), "f"(u1), "f"(t0), "f"(t1));
}
__device__ __forceinline__ float pair_sum(unsigned long long v) {
  return __uint_as_float((unsigned)(v & 0xffffffffull)) + __uint_as_float((unsigned)(v >> 32));
}

// shared-memory matrix descriptor, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;   // descriptor version 1 (Blackwell)
  return d;
}

#define PP_TMEM_LD32(taddr, v)                                                                        \
  asm volatile(                                                                                       \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                       \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"   \
      "%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                          \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),          \
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),     \
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),  \
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),  \
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                            \
      : "r"(taddr))
#define PP_TMEM_LD8(taddr, v)                                                       \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) \
               : "r"(taddr))
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// optional role timing (development): prof[role*8 + k] accumulates clock cycles of CTA 0

}  // namespace tcx
}  // namespace pp
