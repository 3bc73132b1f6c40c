// pfn_pad.cu -- the padding pass of the fused input path (pp_input_path) on tcgen05 tensor cores, sm_100a.
//
// A padding slot (p, n) of the network input holds x = 0 - data_mean[:, p, n] in every sweep
// (data/dataset.py:99-105), so y_pad[c, p, n] = W x + b is evaluated ONCE per (p, n), whatever the batch
// size and whatever the pillars' point counts.  This pass produces
//   * padtab[p][5][64]: the sign-selected extreme of y_pad over the slot suffixes n >= 2, 4, 8, 16, 48
//     (a pillar with cnt points needs the extreme over n >= cnt; k_pfn_real evaluates the few slots between
//     cnt and the next ladder boundary itself and takes the rest from this table), and
//   * per-channel sums of |y_pad| and y_pad |y_pad| over all (p, n).  With the first and second moments of
//     the input (a constant of the dataset, pp_mean_prepare) they give the BatchNorm statistics of the
//     padding slots: sum relu(y) = (sum y + sum |y|) / 2, sum relu(y)^2 = (sum y^2 + sum y |y|) / 2, where
//     sum y and sum y^2 are linear / quadratic forms of the moments (k_bn_finalize).  That is one FADD and one
//     FFMA per accumulator value instead of three FMA-pipe operations for relu, sum and sum of squares: the
//     epilogue is bound by the FMA pipe (scripts/ubench/epi4.cu: 4.2 -> 3.15 cycles per warp-value).
//
// Operand (pp_mean_prepare, once per data_mean): x = -mean split into fp16 pieces x = xh + xl, stored in the
// exact shared-memory image of the UMMA B operand (K-major, no swizzle): per pillar three planes of
// [N slots][8 halves = 16 B], in memory order P2, P0, P1:
//     P0: xh_0 .. xh_7      P1: xh_8, xl_0 .. xl_6      P2: xl_7, xl_8, 1, 1, xh_8, 0, 0, 0
// so ONE contiguous cp.async.bulk per pillar pair fills a stage and no thread touches the data before the
// tensor core does (the round-1 kernel spent ~235 issue cycles per pair and sub-partition converting fp32 rows).
// 48 bytes per slot instead of 36 for the raw fp32 means.
//
// Contraction (fp32 accumulation in TMEM, M = 64 channels, N = the pillar's slots, K = 32 = two k-steps),
// W' = 256 W = Wh + Wl:
//     k-step 0 = planes P0, P1:  k 0..8 Wh_d x xh_d | k 9..15 Wh_0..6 x xl_0..6
//     k-step 1 = planes P2, P0:  k 16,17 Wh_7, Wh_8 x xl_7, xl_8 | k 18,19 bh, bl x 1, 1 | k 20 Wl_8 x xh_8 |
//                                k 24..31 Wl_0..7 x xh_0..7      (P0 is read by both k-steps: no duplicate in HBM)
// = 256 (W x + b) up to the dropped Wl*xl term (~2^-22 |w||x|); rows pre-multiplied by sign(gamma) so that only
// a maximum is tracked.  Four tcgen05.mma per pillar pair (a first version with a separate Wl operand needed six
// and was bound by the tensor pipe: M = 64 runs at half rate, ~150 cycles per instruction).
//
// Roles (576 threads, one persistent CTA per SM, work unit = pair of adjacent pillars at TMEM lane offset 16):
//   warps 0-15  epilogue in two groups: warps 0-7 own accumulator buffer 0 (even pairs of the CTA), warps 8-15
//               buffer 1 (odd pairs); lane quarter q = warp % 4, column half = (warp / 4) % 2 ([0,104) | [104,N)).
//               The groups run half a period apart, so one group's per-visit latencies (barrier wake-up, TMEM
//               load, hand-over) hide under the other group's arithmetic: 73 -> 62 us against all 16 warps on
//               the same pair.  The two halves of a row meet in shared memory; the upper-half warp writes the
//               table rows.  A pure read of the 230 MB operand alone takes 50 us here (scripts/ubench/rd.cu).
//   warp 16     producer (one bulk copy per pair into an 8-stage ring)
//   warp 17     TMEM allocation + MMA issuer (4 tcgen05.mma per pair)
#include "tc_common.cuh"
#include "internal.cuh"

namespace pp {
namespace padk {

using namespace tcx;

constexpr int kThreads = 18 * 32;
constexpr int kProducerWarp = 16, kMmaWarp = 17;
constexpr int kMaxStages = 8;
constexpr int kAccCols = 256;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kABytes = 2 * 2048;          // two k-steps of 64 rows x 16 k fp16
constexpr int kRows = 5;                   // table rows: suffix extremes from slot 2, 4, 8, 16, 48
constexpr int kPartSlot = 5 * 128;         // floats per hand-over slot: five partial maxima of the lower column half x 128 rows

struct Smem {
  int a_off, stage_off, stage_bytes, stages, part_off, stat_off, bar_off, total;
};

__host__ __device__ inline Smem smem_plan(int N) {
  Smem s;
  s.a_off = 0;
  s.stage_off = kABytes;                               // 4096: 128-byte aligned
  s.stage_bytes = 2 * 3 * N * 16;                      // pair of pillars x 3 planes x N x 16 B
  const int fixed = kABytes + 4 * kPartSlot * 4 + 4 * 2 * 64 * 8 + 512 + 128;
  int st = (kSmemBudget - fixed) / s.stage_bytes;
  s.stages = st > kMaxStages ? kMaxStages : st;
  s.part_off = s.stage_off + s.stages * s.stage_bytes;
  s.stat_off = s.part_off + 4 * kPartSlot * 4;
  s.bar_off = s.stat_off + 4 * 2 * 64 * 8;
  s.total = s.bar_off + 512 + 128;                     // +128: manual base alignment
  return s;
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ unsigned short h_bits(float v) {
  unsigned short r;
  asm("cvt.rn.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float h_value(unsigned short h) {
  float r;
  asm("cvt.f32.f16 %0, %1;" : "=f"(r) : "h"(h));
  return r;
}

#define PP_TMEM_LD16(taddr, v)                                                                      \
  asm volatile(                                                                                     \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                     \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                             \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),        \
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),   \
        "=r"(v[14]), "=r"(v[15])                                                                    \
      : "r"(taddr))

// max over v[FROM, CNT) into (m0, m1); |v| and v|v| of all CNT values into four fp32 chains each
template <int CNT, int FROM, bool TRAIN>
__device__ __forceinline__ void consume(const uint32_t* v, float& m0, float& m1, float* S, float* Q) {
#pragma unroll
  for (int i = FROM; i < CNT; i += 2) {
    const float a = __uint_as_float(v[i]), b = __uint_as_float(v[i + 1]);
    if (((i >> 1) & 1) == 0) m0 = fmaxf(m0, fmaxf(a, b)); else m1 = fmaxf(m1, fmaxf(a, b));
  }
  if (TRAIN) {
#pragma unroll
    for (int i = 0; i < CNT; ++i) {
      const float y = __uint_as_float(v[i]);
      S[i & 3] += fabsf(y);
      Q[i & 3] = fmaf(y, fabsf(y), Q[i & 3]);
    }
  }
}

// columns [a, b) of this thread's TMEM row, widest loads first; b - a is a multiple of 8
template <bool TRAIN>
__device__ __forceinline__ void consume_range(uint32_t taddr, int a, int b, float& m0, float& m1, float* S, float* Q) {
  uint32_t v[32];
  int col = a;
#pragma unroll
  for (; col + 32 <= b; col += 32) {
    PP_TMEM_LD32(taddr + col, v);
    tmem_ld_wait();
    consume<32, 0, TRAIN>(v, m0, m1, S, Q);
  }
  if (b - col >= 16) {
    PP_TMEM_LD16(taddr + col, v);
    tmem_ld_wait();
    consume<16, 0, TRAIN>(v, m0, m1, S, Q);
    col += 16;
  }
  if (b - col >= 8) {
    PP_TMEM_LD8(taddr + col, v);
    tmem_ld_wait();
    consume<8, 0, TRAIN>(v, m0, m1, S, Q);
  }
}

// NT: compile-time copy of N for the reference shape (the column loops unroll completely), 0 = any supported N
template <bool TRAIN, int NT>
__global__ void __launch_bounds__(kThreads, 1)
k_pfn_pad_tc(const unsigned char* __restrict__ hl, int P, int N,
             const float* __restrict__ conv_w, const float* __restrict__ conv_b, const float* __restrict__ bn_w,
             float* __restrict__ padtab, double* __restrict__ partials, int* __restrict__ range_flag,
             long long* __restrict__ prof, int dbg) {
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((128u - (smem_u32(smem_unaligned) & 127u)) & 127u);
  const Smem sp = smem_plan(N);
  const int warp = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31u;
  const int R = sp.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bar_off);
  uint64_t* full = bars;                            // [kMaxStages] stage filled by the bulk copy
  uint64_t* empty = bars + kMaxStages;              // [kMaxStages] stage read by the MMAs
  uint64_t* acc_full = bars + 2 * kMaxStages;       // [2]
  uint64_t* acc_empty = acc_full + 2;               // [2]
  uint64_t* pbar = acc_empty + 2;                   // [group][visit parity][quarter] lower column half published
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(pbar + 16);
  float* s_part = reinterpret_cast<float*>(smem + sp.part_off);      // [group][visit parity][5][128]
  double* s_stat = reinterpret_cast<double*>(smem + sp.stat_off);    // [4 ranges][sum |y|, sum y|y|][64]

  const int pairs = P >> 1;                          // host guarantees P even
  const int my_pairs = (int)blockIdx.x < pairs ? (pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  // ---- one-time setup -------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < R; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    for (int i = 0; i < 16; ++i) mbar_init(&pbar[i], 1);
    fence_barrier_init();
  }
  // A tile: 64 x 32 fp16, K-major, no swizzle: core matrix = 8 rows x 16 B (8 k), k-chunk stride 128 B,
  // row-group stride 256 B, k-step stride 2048 B
  {
    bool bad = false;
    for (int idx = threadIdx.x; idx < 64 * 32; idx += kThreads) {
      const int m = idx >> 5, K = idx & 31;
      const float sgn = bn_w[m] < 0.f ? -1.f : 1.f;
      int d = -1;
      bool lo = false;
      if (K < 9) d = K;                           // Wh_d x xh_d
      else if (K < 16) d = K - 9;                 // Wh_d x xl_d, d = 0..6
      else if (K < 18) d = K - 9;                 // Wh_7, Wh_8 x xl_7, xl_8
      else if (K == 20) { d = 8; lo = true; }     // Wl_8 x xh_8
      else if (K >= 24) { d = K - 24; lo = true; }  // Wl_d x xh_d, d = 0..7
      float v = 0.f;
      if (d >= 0) {
        const float w = 256.f * conv_w[m * 9 + d];
        bad |= !(fabsf(w) < 32768.f);
        const float wh = h_value(h_bits(w));
        v = lo ? (w - wh) : wh;
      } else if (K == 18 || K == 19) {
        const float bb = 256.f * conv_b[m];
        bad |= !(fabsf(bb) < 32768.f);
        const float bh = h_value(h_bits(bb));
        v = K == 18 ? bh : (bb - bh);
      }
      const int j = K >> 4, kk = K & 15;
      *reinterpret_cast<unsigned short*>(smem + sp.a_off + j * 2048 + (m >> 3) * 256 + (kk >> 3) * 128 + (m & 7) * 16 + (kk & 7) * 2) =
          h_bits(sgn * v);
    }
    if (bad) atomicOr(range_flag, 1);
  }
  fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
  if (warp == kMmaWarp) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // development timing (pp_debug.h): per-role wait cycles of CTA 0
  const bool pon = prof != nullptr && blockIdx.x == 0;
  long long pacc[4] = {0, 0, 0, 0};
  const long long prole0 = pon ? clock64() : 0;

  if (warp == kProducerWarp) {
    // ===== producer: one contiguous bulk copy per pair.  The whole warp runs the loop converged and one elected
    // lane issues (elect.sync): the loop state stays in uniform registers.  A lane-0-only role compiled into
    // ELECT / BRA.U.ANY retry loops around every uniform-datapath instruction.
    int s = 0;
    uint32_t ph = 1;                                  // parity to wait on empty[] (first pass falls through)
    const uint32_t bytes = (uint32_t)sp.stage_bytes;
    for (int it = 0; it < my_pairs; ++it) {
      const size_t pair = (size_t)blockIdx.x + (size_t)it * gridDim.x;
      mbar_wait_t(&empty[s], ph, pon, pacc[0]);
      if (elect_one()) {
        if (dbg & 16) { mbar_arrive(&full[s]); }
        else
        mbar_expect_tx(&full[s], bytes);
        if (dbg & 16) {
        } else if (dbg & 8) {
          const uint32_t part = bytes / 6;
          for (int k = 0; k < 6; ++k)
            bulk_g2s(smem + sp.stage_off + s * sp.stage_bytes + k * part, hl + pair * bytes + k * part, part, &full[s]);
        } else {
          bulk_g2s(smem + sp.stage_off + s * sp.stage_bytes, hl + pair * bytes, bytes, &full[s]);
        }
      }
      __syncwarp();
      if (++s == R) { s = 0; ph ^= 1u; }
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer (converged warp, one elected lane issues) =====
    // instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): fp32 accumulate, A = B = f16,
    // both K-major, N>>3, M>>4
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
    const uint32_t a_addr = smem_u32(smem + sp.a_off);
    const uint32_t plane = (uint32_t)N * 16u;
    const uint64_t a0 = smem_desc(a_addr, 128, 256), a1 = smem_desc(a_addr + 2048, 128, 256);
    // B descriptors differ only in the start-address field (bits 0..13, units of 16 bytes)
    const uint64_t b_base = smem_desc(smem_u32(smem + sp.stage_off), plane, 128);
    const uint32_t plane16 = plane >> 4, stage16 = (uint32_t)sp.stage_bytes >> 4;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < my_pairs; ++it) {
      const int e = it & 1;
      mbar_wait_spin_t(&full[s], ph, pon, pacc[0]);
      mbar_wait_spin_t(&acc_empty[e], ((uint32_t)(it >> 1) & 1u) ^ 1u, pon, pacc[1]);
      tc_fence_after();
      const long long tq0 = pon ? clock64() : 0;
      if (elect_one()) {
        const uint64_t b0 = b_base + (uint64_t)((uint32_t)s * stage16);       // pillar h = 0: P2 | P0 | P1
        const uint32_t d0 = tmem_base + (uint32_t)(e * kAccCols);
        const uint64_t b1 = b0 + 3u * plane16;                                // pillar h = 1
        const uint32_t d1 = d0 + (16u << 16);
        if (!(dbg & 2)) {
          umma_f16(d0, a0, b0 + plane16, idesc, 0u);                          // k-step 0: P0, P1
          umma_f16(d0, a1, b0, idesc, 1u);                                    // k-step 1: P2, P0
          umma_f16(d1, a0, b1 + plane16, idesc, 0u);
          umma_f16(d1, a1, b1, idesc, 1u);
        }
        umma_commit(&empty[s]);       // stage free once these MMAs have read it
        umma_commit(&acc_full[e]);    // accumulators ready for the epilogue
      }
      __syncwarp();
      if (pon) pacc[2] += clock64() - tq0;
      if (++s == R) { s = 0; ph ^= 1u; }
    }
  } else {
    // ===== epilogue, two groups: warps 0-7 own accumulator buffer 0 (even pairs), warps 8-15 buffer 1 (odd pairs).
    // The groups run half a period apart, so on every SM sub-partition one group's fixed per-visit latencies
    // (barrier wake-up, TMEM load latency, the hand-over through shared memory) overlap the other group's
    // arithmetic.  With all 16 warps on the same pair those latencies were exposed on every visit.
    const int q = warp & 3;                // TMEM lane quarter (must equal warp % 4)
    const int e = warp >> 3;               // accumulator buffer = pair parity
    const int jh = (warp >> 2) & 1;        // column half: [0, hs) | [hs, N)
    const int h = (int)(lane >> 4);
    const int c = 16 * q + (int)(lane & 15u);
    const int rowid = 32 * q + (int)lane;
    const float sgn = bn_w[c] < 0.f ? -1.f : 1.f;
    const int hs = N < 104 ? N : 104;
    float sh = 0.f, sl = 0.f, qh = 0.f, ql = 0.f;
    auto two_sum = [](float& hi, float& lo, float b) {
      const float s = __fadd_rn(hi, b);
      const float bb = __fsub_rn(s, hi);
      lo = __fadd_rn(lo, __fadd_rn(__fsub_rn(hi, __fsub_rn(s, bb)), __fsub_rn(b, bb)));
      hi = s;
    };
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(e * kAccCols);
    for (int it = e; it < my_pairs; it += 2) {
      const uint32_t n = (uint32_t)(it >> 1);
      mbar_wait_spin_t(&acc_full[e], n & 1u, pon, pacc[0]);
      tc_fence_after();
      const long long tq0 = pon ? clock64() : 0;
      float S[4] = {0.f, 0.f, 0.f, 0.f}, Q[4] = {0.f, 0.f, 0.f, 0.f};
      float m0 = -INFINITY, m1 = -INFINITY, lo2 = -INFINITY, lo4 = -INFINITY, lo8 = -INFINITY, m16 = -INFINITY;
      auto release = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[e]);
      };
      auto head = [&](const uint32_t* va) {
        float a0 = -INFINITY, a1 = -INFINITY;
        consume<16, 8, TRAIN>(va, a0, a1, S, Q);
        lo8 = fmaxf(a0, a1);                                                       // slots 8..15
        lo4 = fmaxf(fmaxf(__uint_as_float(va[4]), __uint_as_float(va[5])), fmaxf(__uint_as_float(va[6]), __uint_as_float(va[7])));
        lo2 = fmaxf(__uint_as_float(va[2]), __uint_as_float(va[3]));
      };
      if (dbg & 4) {
        release();
      } else if (NT == 200) {
        if (jh == 0) {
          uint32_t va[16], vb[32], vc[32], vd[16], ve[8];
          PP_TMEM_LD16(taddr, va);
          PP_TMEM_LD32(taddr + 16, vb);
          tmem_ld_wait();
          head(va);
          PP_TMEM_LD32(taddr + 48, vc);
          PP_TMEM_LD16(taddr + 80, vd);
          PP_TMEM_LD8(taddr + 96, ve);
          consume<32, 0, TRAIN>(vb, m0, m1, S, Q);
          m16 = fmaxf(m0, m1);                                                     // slots 16..47
          m0 = m1 = -INFINITY;
          tmem_ld_wait();
          release();
          consume<32, 0, TRAIN>(vc, m0, m1, S, Q);
          consume<16, 0, TRAIN>(vd, m0, m1, S, Q);
          consume<8, 0, TRAIN>(ve, m0, m1, S, Q);
        } else {
          uint32_t va[32], vb[32], vc[32];
          PP_TMEM_LD32(taddr + 104, va);
          PP_TMEM_LD32(taddr + 136, vb);
          tmem_ld_wait();
          PP_TMEM_LD32(taddr + 168, vc);
          consume<32, 0, TRAIN>(va, m0, m1, S, Q);
          consume<32, 0, TRAIN>(vb, m0, m1, S, Q);
          tmem_ld_wait();
          release();
          consume<32, 0, TRAIN>(vc, m0, m1, S, Q);
        }
      } else {
        if (jh == 0) {
          uint32_t v[16];
          PP_TMEM_LD16(taddr, v);
          tmem_ld_wait();
          head(v);
          consume_range<TRAIN>(taddr, 16, hs < 48 ? hs : 48, m0, m1, S, Q);
          m16 = fmaxf(m0, m1);
          m0 = m1 = -INFINITY;
          if (hs > 48) consume_range<TRAIN>(taddr, 48, hs, m0, m1, S, Q);
        } else {
          consume_range<TRAIN>(taddr, hs, N, m0, m1, S, Q);
        }
        release();
      }
      const float m = fmaxf(m0, m1);
      // the two column halves of a row meet in shared memory; two slots per group (visit parity): the lower-half
      // warp can only write a slot again after the accumulator hand-off that follows the upper-half warp's read
      float* part = s_part + (e * 2 + (int)(n & 1u)) * kPartSlot;
      uint64_t* pb = &pbar[(e * 2 + (int)(n & 1u)) * 4 + q];
      if (jh == 0) {
        part[rowid] = lo2; part[128 + rowid] = lo4; part[256 + rowid] = lo8; part[384 + rowid] = m16; part[512 + rowid] = m;
        __syncwarp();
        if (lane == 0) mbar_arrive(pb);
        if (pon) pacc[1] += clock64() - tq0;
      } else {
        if (pon) pacc[1] += clock64() - tq0;
        mbar_wait_spin_t(pb, (n >> 1) & 1u, pon, pacc[2]);
        const float t = fmaxf(m, part[512 + rowid]);                                        // slots >= 48
        const float r16 = fmaxf(part[384 + rowid], t);
        const float r8 = fmaxf(part[256 + rowid], r16);
        const float r4 = fmaxf(part[128 + rowid], r8);
        const float r2 = fmaxf(part[rowid], r4);
        const size_t pillar = 2 * ((size_t)blockIdx.x + (size_t)it * gridDim.x) + h;
        float* o = padtab + pillar * (kRows * 64) + c;
        o[0] = sgn * r2 * (1.f / 256.f);
        o[64] = sgn * r4 * (1.f / 256.f);
        o[128] = sgn * r8 * (1.f / 256.f);
        o[192] = sgn * r16 * (1.f / 256.f);
        o[256] = sgn * t * (1.f / 256.f);
      }
      if (TRAIN) {
        two_sum(sh, sl, (S[0] + S[1]) + (S[2] + S[3]));
        two_sum(qh, ql, (Q[0] + Q[1]) + (Q[2] + Q[3]));
      }
    }
    if (TRAIN) {
      double accS = (double)sh + (double)sl, accQ = (double)qh + (double)ql;
      accS += __shfl_xor_sync(0xffffffffu, accS, 16);
      accQ += __shfl_xor_sync(0xffffffffu, accQ, 16);
      const int jj = warp >> 2;
      if (lane < 16) {
        s_stat[(jj * 2 + 0) * 64 + c] = accS * (1.0 / 256.0);
        s_stat[(jj * 2 + 1) * 64 + c] = accQ * ((double)sgn / (256.0 * 256.0));
      }
    }
  }

  if (pon && lane == 0) {
    pacc[3] = clock64() - prole0;
    for (int kq = 0; kq < 4; ++kq) prof[warp * 4 + kq] = pacc[kq];
  }
  // ---- teardown ---------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (TRAIN && threadIdx.x < 128) {
    // fixed-order combination of the four column ranges -> deterministic per-CTA partials
    const int qq = threadIdx.x >> 6, cc = threadIdx.x & 63;
    double v = 0.0;
    for (int k = 0; k < 4; ++k) v += s_stat[(k * 2 + qq) * 64 + cc];
    partials[((size_t)blockIdx.x * 2 + qq) * 64 + cc] = v;
  }
}

// ---- pp_mean_prepare: data_mean [9,P,N] fp32 -> operand image + moments (once per data_mean) ------------
constexpr int kPrepBlocks = 592;
constexpr int kMom = 55;            // upper triangle of the 10 x 10 moment matrix of (x_0..x_8, 1)

__global__ void __launch_bounds__(256) k_mean_prepare(const float* __restrict__ mean, int P, int N,
                                                      unsigned char* __restrict__ hl, double* __restrict__ scratch,
                                                      int* __restrict__ flag) {
  __shared__ double s_red[8][kMom];
  const long long PN = (long long)P * N;
  double acc[kMom];
#pragma unroll
  for (int i = 0; i < kMom; ++i) acc[i] = 0.0;
  bool bad = false;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < PN; i += (long long)gridDim.x * 256) {
    const int p = (int)(i / N), n = (int)(i - (long long)p * N);
    unsigned short xh[9], xl[9];
    double xv[10];
#pragma unroll
    for (int d = 0; d < 9; ++d) {
      const float x = __fsub_rn(0.f, mean[(long long)d * PN + i]);       // what a padding slot holds
      bad |= !(fabsf(x) < 32768.f);
      xh[d] = h_bits(x);
      const float fh = h_value(xh[d]);
      xl[d] = h_bits(x - fh);                                             // x - fh is exact in fp32
      xv[d] = (double)fh + (double)h_value(xl[d]);                        // the value the tensor core sees
    }
    xv[9] = 1.0;
    int k = 0;
#pragma unroll
    for (int d = 0; d < 10; ++d)
#pragma unroll
      for (int e = d; e < 10; ++e) acc[k++] += xv[d] * xv[e];
    uint4 w0, w1, w2;
    w0.x = xh[0] | ((unsigned)xh[1] << 16); w0.y = xh[2] | ((unsigned)xh[3] << 16);
    w0.z = xh[4] | ((unsigned)xh[5] << 16); w0.w = xh[6] | ((unsigned)xh[7] << 16);
    w1.x = xh[8] | ((unsigned)xl[0] << 16); w1.y = xl[1] | ((unsigned)xl[2] << 16);
    w1.z = xl[3] | ((unsigned)xl[4] << 16); w1.w = xl[5] | ((unsigned)xl[6] << 16);
    w2.x = xl[7] | ((unsigned)xl[8] << 16); w2.y = 0x3c003c00u; w2.z = xh[8]; w2.w = 0u;
    uint4* o = reinterpret_cast<uint4*>(hl + ((size_t)p * 3 * N + n) * 16);
    o[0] = w2;              // memory order P2, P0, P1
    o[N] = w0;
    o[2 * N] = w1;
  }
  if (bad) atomicOr(flag, 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < kMom; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < kMom) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += s_red[w][threadIdx.x];
    scratch[(size_t)blockIdx.x * kMom + threadIdx.x] = v;
  }
}

__global__ void k_mean_moments(const double* __restrict__ scratch, int nblk, double* __restrict__ mom) {
  __shared__ double s[kMom];
  if (threadIdx.x < kMom) {
    double v = 0.0;
    for (int b = 0; b < nblk; ++b) v += scratch[(size_t)b * kMom + threadIdx.x];     // fixed order: deterministic
    s[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x < 100) {
    int d = threadIdx.x / 10, e = threadIdx.x % 10;
    if (d > e) { const int t = d; d = e; e = t; }
    const int k = d * 10 - d * (d - 1) / 2 + (e - d);
    mom[threadIdx.x] = s[k];
  }
}

struct PrepLayout { size_t hl, mom, flag, scratch, total; };
static PrepLayout prep_layout(int P, int N) {
  PrepLayout l;
  l.hl = 0;
  l.mom = align_up((size_t)P * N * 48);
  l.flag = l.mom + align_up(100 * sizeof(double));
  l.scratch = l.flag + kAlign;
  l.total = l.scratch + align_up((size_t)kPrepBlocks * kMom * sizeof(double));
  return l;
}

}  // namespace padk

bool pfn_pad_supported(int N, int C, int P) {
  return C == 64 && N >= 16 && N <= 256 && (N % 8) == 0 && (P % 2) == 0 && padk::smem_plan(N).stages >= 2;
}

size_t mean_prepared_bytes(int P, int N) { return padk::prep_layout(P, N).total; }
const double* mean_prepared_moments(const void* prep, int P, int N) {
  return reinterpret_cast<const double*>((const char*)prep + padk::prep_layout(P, N).mom);
}
const int* mean_prepared_flag(const void* prep, int P, int N) {
  return reinterpret_cast<const int*>((const char*)prep + padk::prep_layout(P, N).flag);
}

int mean_prepare(const float* d_mean, int P, int N, void* d_prep, size_t bytes, cudaStream_t st) {
  using namespace padk;
  if (d_mean == nullptr || d_prep == nullptr || P < 1 || N < 1 || ((uintptr_t)d_prep % kAlign) != 0) return PP_ERR_INVALID_ARG;
  const PrepLayout l = prep_layout(P, N);
  if (bytes < l.total) return PP_ERR_WORKSPACE;
  char* base = (char*)d_prep;
  PP_CUDA(cudaMemsetAsync(base + l.flag, 0, sizeof(int), st));
  PP_KERNEL("k_mean_prepare", st,
            k_mean_prepare<<<kPrepBlocks, 256, 0, st>>>(d_mean, P, N, (unsigned char*)base + l.hl, (double*)(base + l.scratch),
                                                        (int*)(base + l.flag)));
  PP_KERNEL("k_mean_moments", st, k_mean_moments<<<1, 128, 0, st>>>((const double*)(base + l.scratch), kPrepBlocks,
                                                                   (double*)(base + l.mom)));
  return PP_OK;
}

// padding pass over the prepared operand: padtab [P][3][64], partials [nblocks][2][64] (training only)
extern int g_opt_pfn_tc_timing;
extern int g_opt_pfn_tc_debug;
long long* tc_prof_ptr();   // pfn_tc.cu

int launch_pad_tc(const void* d_prep, int P, int N, const float* w, const float* bias, const float* bn_w, int training,
                  float* padtab, double* partials, int nblocks, int* range_flag, cudaStream_t st) {
  using namespace padk;
  long long* prof = g_opt_pfn_tc_timing ? tc_prof_ptr() : nullptr;
  const Smem sp = smem_plan(N);
  if (sp.stages < 2) return PP_ERR_UNSUPPORTED;
  const unsigned char* hl = (const unsigned char*)d_prep + prep_layout(P, N).hl;
  PP_CUDA(cudaMemsetAsync(range_flag, 0, sizeof(int), st));
#define PP_PAD(TR, NT)                                                                                              \
  do {                                                                                                              \
    PP_CUDA(cudaFuncSetAttribute(k_pfn_pad_tc<TR, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.total));     \
    PP_KERNEL("k_pfn_pad_tc", st,                                                                                   \
              (k_pfn_pad_tc<TR, NT><<<nblocks, kThreads, sp.total, st>>>(hl, P, N, w, bias, bn_w, padtab, partials, \
                                                                         range_flag, prof, g_opt_pfn_tc_debug)));   \
  } while (0)
  if (N == 200) {
    if (training) PP_PAD(true, 200); else PP_PAD(false, 200);
  } else {
    if (training) PP_PAD(true, 0); else PP_PAD(false, 0);
  }
#undef PP_PAD
  return PP_OK;
}

}  // namespace pp
