// pfn_tc16.cu -- K2 statistics pass on tcgen05 tensor cores with an fp16 error-compensated split
// (kind::f16, fp32 accumulation in TMEM), sm_100a.  Successor of the TF32 kernel in pfn_tc.cu, which
// stays as the guarded fallback for inputs outside the fp16 range.
//
// Why fp16 instead of TF32: every tcgen05.mma of shape M=64, N=200 costs ~100 SM cycles and reads its
// B operand (6.4 KB) from shared memory whatever the element type.  TF32 covers K=8 per instruction,
// fp16 K=16, and both carry 11 significand bits, so the same 3-term compensated product needs four
// TF32 instructions per pillar but only two fp16 ones: half the tensor-pipe time and 35 % less
// shared-memory traffic (ncu r1i: tensor pipe 46 % busy, smem ~530 of 885 cycles per pillar).
//
// Per pillar row (b,p):  Y'[c, n] = sum_k A'[c,k] * B'[k,n],  M = 64 channels, N = the pillar's N
// slots, K = 32 (two k-steps of 16).  With x = xh + xl (xh = x truncated to 11 significand bits,
// xl = x - xh exact in fp32, both stored as fp16) and W' = 256 W = Wh + Wl (fp16 pieces):
//   K = 2d, 2d+1 (d = 0..8)   B' = xh_d, xl_d       A' = Wh_d, Wh_d
//   K = 18 + d                B' = xh_d (copy)      A' = Wl_d
//   K = 27, 28                B' = 1, 1             A' = bh, bl      (b' = 256 b = bh + bl)
//   K = 29..31                0
// => Y' = 256 (W x + b) up to the dropped Wl*xl term (~2^-22 |w||x|).  The factor 256 keeps the
// residual weights Wl out of the fp16 subnormal range for |w| >= 2^-10 and is undone for free when
// the per-pillar maximum and the BatchNorm sums are written.  A' rows are pre-multiplied by
// s_c = sign(gamma_c) (see pfn_tc.cu): only max_n(s_c y) is tracked.
// fp16 range guard: |x| >= 2^15 or |256 w| >= 2^15 raises *range_flag; the caller then runs the TF32
// kernel, which recomputes every output (it exits at once when the flag is clear).  Below 2^-14 the
// fp16 pieces are subnormal: absolute representation error <= 2^-24 per piece, i.e. <= ~6e-8 |w|
// per term, far inside the 1e-5 relative + 2e-6 absolute parity tolerance.
//
// Orientation as in pfn_tc.cu: channels on TMEM lanes, slots on TMEM columns, two pillars interleaved
// at lane offset 16, so every epilogue thread owns one (pillar, channel) row and reduces in registers.
//
// Warp roles (832 threads, one persistent CTA per SM, work unit = pair of adjacent pillars).  The warp
// scheduler prefers the highest warp id among eligible warps, and the epilogue warps are FMA-pipe-bound
// (eligible almost every cycle), so the single-thread roles that feed the pipeline get the HIGHEST ids;
// with the producer as warp 0 the loads only advanced while the epilogue stalled (load time and
// epilogue time added up instead of overlapping):
//   warps 0-15    epilogue: group e = accumulator buffer (pair parity), half j = column range,
//                 quarter q = TMEM lane quarter; tcgen05.ld -> max / sum relu / sum relu^2
//   warps 16-23   converters: raw fp32 rows -> fp16 (xh, xl) k-vectors, K-major no-swizzle layout; two
//                 teams of four warps take alternate pairs (one team alone is latency-bound: ~1000
//                 cycles per pair for ~200 dependent instructions, slower than the HBM stream)
//   warp 24       TMA producer: ONE 3-D tensor-map copy per pair ([9 features][2 pillars][N] box)
//   warp 25       TMEM allocation + MMA issuer (4 tcgen05.mma per pair)
#include <cuda.h>

#include "tc_common.cuh"
#include "internal.cuh"

namespace pp {

namespace tch {

using namespace tcx;

constexpr int kThreads = 832;
constexpr int kConvWarp0 = 16, kProducerWarp = 24, kMmaWarp = 25;
constexpr int kMaxRawStages = 12;
constexpr int kBStages = 3;            // B' operand tiles (one pillar pair each) between converters and MMA
constexpr int kABytes = 2 * 2048;      // 2 k-steps x (64 rows x 16 k) fp16
constexpr int kSboB = 528;             // bytes between 8-slot groups of B' (4 core matrices of 128 B + 16 pad: conflict-free STS.128)
constexpr int kLboB = 128;             // bytes between the two 8-wide k chunks of one k-step
constexpr int kAccCols = 256;          // TMEM column stride between the two accumulator buffers
constexpr int kSmemBudget = 227 * 1024;

struct Smem {
  int a_off, raw_off, b_off, stat_off, bar_off, total;
  int raw_stage_bytes, raw_stages, b_pillar_bytes;
};

__host__ __device__ inline Smem smem_plan(int N) {
  Smem s;
  s.a_off = 0;
  s.raw_off = kABytes;
  s.raw_stage_bytes = (2 * 9 * N * 4 + 127) & ~127;            // TMA tensor destinations are 128-byte aligned
  s.b_pillar_bytes = (N / 8) * kSboB;
  const int fixed = kABytes + kBStages * 2 * s.b_pillar_bytes + 4 * 2 * 64 * 8 + (2 * kMaxRawStages + 2 * kBStages + 4) * 8 + 16 + 256;
  int r = (kSmemBudget - fixed) / s.raw_stage_bytes;
  s.raw_stages = r > kMaxRawStages ? kMaxRawStages : r;
  s.b_off = s.raw_off + s.raw_stages * s.raw_stage_bytes;
  s.stat_off = (s.b_off + kBStages * 2 * s.b_pillar_bytes + 15) & ~15;
  s.bar_off = s.stat_off + 4 * 2 * 64 * 8;
  s.total = s.bar_off + (2 * kMaxRawStages + 2 * kBStages + 4) * 8 + 16 + 128;   // +128: manual base alignment
  return s;
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// {hi16: fp16(a), lo16: fp16(b)}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint32_t lo_halves(uint32_t a, uint32_t b) {   // {hi16: lo16(b), lo16: lo16(a)}
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(r) : "r"(a), "r"(b));
  return r;
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

template <bool TRAIN>
__global__ void __launch_bounds__(kThreads, 1)
k_pfn_stats_tc(const __grid_constant__ CUtensorMap tmap, int B, int P, int N,
               const float* __restrict__ conv_w, const float* __restrict__ conv_b,
               const float* __restrict__ bn_w, float* __restrict__ ext, double* __restrict__ partials,
               int* __restrict__ range_flag, int dbg, long long* __restrict__ prof) {
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((128u - (smem_u32(smem_unaligned) & 127u)) & 127u);
  const Smem sp = smem_plan(N);
  const int warp = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31u;
  const int R = sp.raw_stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bar_off);
  uint64_t* raw_full = bars;
  uint64_t* raw_empty = bars + kMaxRawStages;
  uint64_t* b_full = bars + 2 * kMaxRawStages;     // [kBStages]
  uint64_t* b_empty = b_full + kBStages;           // [kBStages]
  uint64_t* acc_full = b_empty + kBStages;         // [2]
  uint64_t* acc_empty = acc_full + 2;              // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);
  double* s_stat = reinterpret_cast<double*>(smem + sp.stat_off);   // [4 epilogue sets][sum, sum sq][64]

  // host guarantees P even (pairs never straddle two sweeps) and B*P < 2^31
  const int pairs = (B * P) >> 1;
  const int my_pairs = (int)blockIdx.x < pairs ? (pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int G4 = N / 4;

  // ---- one-time setup -------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    for (int i = 0; i < R; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 4); }
    for (int i = 0; i < kBStages; ++i) { mbar_init(&b_full[i], 4); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    fence_barrier_init();
  }
  // A' tile: 64 x 32 fp16, K-major, no swizzle: core matrix = 8 rows x 16 B (8 k), k-chunk stride 128 B,
  // row-group stride 256 B, k-step stride 2048 B
  {
    bool bad = false;
    for (int idx = threadIdx.x; idx < 64 * 32; idx += kThreads) {
      const int m = idx >> 5, K = idx & 31;
      const float sgn = bn_w[m] < 0.f ? -1.f : 1.f;
      float v = 0.f;
      if (K < 18 || (K >= 18 && K <= 26)) {
        const int d = K < 18 ? (K >> 1) : (K - 18);
        const float w = 256.f * conv_w[m * 9 + d];
        bad |= !(fabsf(w) < 32768.f);
        const float wh = h_value(h_bits(w));
        v = K < 18 ? wh : (w - wh);
      } else if (K == 27 || K == 28) {
        const float bb = 256.f * conv_b[m];
        bad |= !(fabsf(bb) < 32768.f);
        const float bh = h_value(h_bits(bb));
        v = K == 27 ? bh : (bb - bh);
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
  const bool pon = prof != nullptr && blockIdx.x == 0;
  long long pacc[4] = {0, 0, 0, 0};
  const long long prole0 = pon ? clock64() : 0;

  // ---- roles ----------------------------------------------------------------------------------
  if (warp == kProducerWarp) {
    // ===== TMA producer: one 3-D box {N, 2 pillars, 9 features} per pair; smem layout [d][h][N] =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;                                  // parity to wait on raw_empty (first pass falls through)
      int r0 = 2 * (int)blockIdx.x;                     // first row of the current pair
      int b0 = r0 / P, p0 = r0 - b0 * P;                // (sweep, pillar) of that row, advanced incrementally
      const int step = 2 * (int)gridDim.x;
      const uint32_t bytes = (uint32_t)(2 * 9 * N * 4);
      for (int it = 0; it < my_pairs; ++it) {
        mbar_wait_t(&raw_empty[s], ph, pon, pacc[0]);
        mbar_expect_tx(&raw_full[s], bytes);
        tma_load_3d(smem + sp.raw_off + s * sp.raw_stage_bytes, &tmap, 0, p0, b0 * 9, &raw_full[s]);
        p0 += step;
        while (p0 >= P) { p0 -= P; ++b0; }
        if (++s == R) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): fp32 accumulate,
      // A = B = f16 (format 0), both K-major, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
      const uint32_t a_addr = smem_u32(smem + sp.a_off);
      int bs = 0;
      uint32_t bph = 0;
      for (int it = 0; it < my_pairs; ++it) {
        const int t = it & 1;                          // accumulator buffer
        const uint32_t n = (uint32_t)(it >> 1);
        mbar_wait_t(&b_full[bs], bph, pon, pacc[0]);
        mbar_wait_t(&acc_empty[t], (n & 1u) ^ 1u, pon, pacc[1]);
        tc_fence_after();
        const long long tq0 = pon ? clock64() : 0;
        for (int h = 0; h < 2 && !(dbg & 2); ++h) {
          const uint32_t b_addr = smem_u32(smem + sp.b_off + (bs * 2 + h) * sp.b_pillar_bytes);
          const uint32_t d_tmem = tmem_base + ((uint32_t)(h * 16) << 16) + (uint32_t)(t * kAccCols);
#pragma unroll
          for (int j = 0; j < ((dbg & 8) ? 1 : 2); ++j) {
            const uint64_t ad = smem_desc(a_addr + j * 2048, 128, 256);
            const uint64_t bd = smem_desc(b_addr + j * 2 * kLboB, kLboB, kSboB);
            umma_f16(d_tmem, ad, bd, idesc, j > 0 ? 1u : 0u);
          }
        }
        umma_commit(&b_empty[bs]);    // B' tile free once these MMAs have read it
        umma_commit(&acc_full[t]);    // accumulators ready for the epilogue
        if (pon) pacc[2] += clock64() - tq0;
        if (++bs == kBStages) { bs = 0; bph ^= 1u; }
      }
    }
  } else if (warp < kConvWarp0) {
    // ===== epilogue: one (pillar-of-pair, channel) row per thread over this warp's column range =====
    const int k = warp >> 2;               // epilogue set 0..3
    const int e = k >> 1;                  // accumulator buffer / pair parity
    const int j = k & 1;                   // column half
    const int q = warp & 3;                // TMEM lane quarter (must equal warp % 4)
    const int h = lane >> 4;
    const int c = 16 * q + (int)(lane & 15u);
    const float sgn = bn_w[c] < 0.f ? -1.f : 1.f;
    // running sums of the visits as unevaluated fp32 pairs (TwoSum): keeps the fp64 pipe (F2F + DADD chain, 13 %
    // of the stall samples of the padding pass in ncu r2d) out of the loop
    float sh = 0.f, sl = 0.f, qh = 0.f, ql = 0.f;
    auto two_sum = [](float& hi, float& lo, float b) {
      const float s2 = __fadd_rn(hi, b);
      const float bb = __fsub_rn(s2, hi);
      lo = __fadd_rn(lo, __fadd_rn(__fsub_rn(hi, __fsub_rn(s2, bb)), __fsub_rn(b, bb)));
      hi = s2;
    };
    const int split = ((N / 8 + 1) / 2) * 8;
    const int n0 = j ? split : 0, n1 = j ? N : split;
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(e * kAccCols + n0);
    const int ncols = n1 - n0;             // multiple of 8
    for (int it = e; it < my_pairs; it += 2) {
      const uint32_t n = (uint32_t)(it >> 1);
      const int r = 2 * ((int)blockIdx.x + it * (int)gridDim.x) + h;      // row: pillar (b,p)
      mbar_wait_t(&acc_full[e], n & 1u, pon, pacc[0]);
      tc_fence_after();
      const long long tq0 = pon ? clock64() : 0;
      // four independent accumulator sets break the dependent chains; sums are packed fp32 pairs.
      // 2*relu(s*y) = s*y + |y| is one FFMA (the ALU pipe that FMNMX runs on is half rate).
      float mx[2] = {-INFINITY, -INFINITY};
      constexpr int NCH = 4;                 // accumulator chains
      unsigned long long S[4] = {0ull, 0ull, 0ull, 0ull}, Q[4] = {0ull, 0ull, 0ull, 0ull};
      // three passes over the chunk keep every instruction independent of its neighbours (the t values
      // replace the y values in place).  One 32-column tcgen05.ld per chunk, no register double buffer:
      // the four epilogue warps of a scheduler hide each other's TMEM latency, and ptxas sinks a
      // prefetching tcgen05.ld below the arithmetic anyway (measured: x32 single 4.2 cycles per
      // warp-value per scheduler, x16 double-buffered 5.1; scripts/ubench/epi3.cu)
      auto consume = [&](uint32_t* v, const int cnt, const int col0) {
        {
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            if (i < cnt) mx[(i >> 1) & 1] = fmaxf(mx[(i >> 1) & 1], fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
        }
        if (TRAIN) {
          {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < cnt) { const float y = __uint_as_float(v[i]); v[i] = __float_as_uint(fmaf(sgn, y, fabsf(y))); }
#pragma unroll
            for (int i = 0; i < 32; i += 2)
              if (i < cnt) acc_pair(S[(i >> 1) & (NCH - 1)], Q[(i >> 1) & (NCH - 1)], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
          }
        }
      };
      if (!(dbg & 4)) {
        uint32_t v[32];
        int col = 0;
        // the narrow chunks come first
        if (ncols & 8) {
          PP_TMEM_LD8(taddr + col, v);
          tmem_ld_wait();
          consume(v, 8, n0 + col);
          col += 8;
        }
        if (ncols & 16) {
          PP_TMEM_LD16(taddr + col, v);
          tmem_ld_wait();
          consume(v, 16, n0 + col);
          col += 16;
        }
        for (; col + 32 <= ncols; col += 32) {
          PP_TMEM_LD32(taddr + col, v);
          tmem_ld_wait();
          consume(v, 32, n0 + col);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[e]);
      if (pon) pacc[1] += clock64() - tq0;
      // TMEM holds 256*s*y: the partial extreme of this column half goes to ext[r][j][c]; consumers
      // combine the fields (max when gamma >= 0, min otherwise)
      ext[(size_t)r * 128 + j * 64 + c] = sgn * fmaxf(mx[0], mx[1]) * (1.f / 256.f);
      if (TRAIN) {
        two_sum(sh, sl, (pair_sum(S[0]) + pair_sum(S[1])) + (pair_sum(S[2]) + pair_sum(S[3])));
        two_sum(qh, ql, (pair_sum(Q[0]) + pair_sum(Q[1])) + (pair_sum(Q[2]) + pair_sum(Q[3])));
      }
    }
    if (TRAIN) {
      double accS = (double)sh + (double)sl, accQ = (double)qh + (double)ql;
      accS += __shfl_xor_sync(0xffffffffu, accS, 16);
      accQ += __shfl_xor_sync(0xffffffffu, accQ, 16);
      if (lane < 16) {
        s_stat[(k * 2 + 0) * 64 + c] = accS * (1.0 / 512.0);              // t = 512 relu(y)
        s_stat[(k * 2 + 1) * 64 + c] = accQ * (1.0 / (512.0 * 512.0));
      }
    }
  } else if (warp < kProducerWarp) {
    // ===== converters: raw fp32 rows -> packed fp16 (xh | xl) words, UMMA K-major layout =====
    const int team = (warp - kConvWarp0) >> 2;                      // takes pairs it = team, team + 2, ...
    const int ct = (threadIdx.x - kConvWarp0 * 32) & 127;           // 0..127 within the team
    const int h = ct >> 6;                  // pillar of the pair
    const int g = ct & 63;                  // 4-slot group
    // ring positions of pair `team` (raw stage it % R, B' stage it % kBStages), advanced by two pairs per turn
    int s = team % R, t = team % kBStages;
    uint32_t ph_raw = (uint32_t)((team / R) & 1), ph_b = (uint32_t)((team / kBStages) & 1) ^ 1u;
    float amax = 0.f;
    for (int it = team; it < my_pairs; it += 2) {
      mbar_wait_t(&raw_full[s], ph_raw, pon, pacc[0]);
      mbar_wait_t(&b_empty[t], ph_b, pon, pacc[1]);
      if (g < G4 && !(dbg & 1)) {
        const unsigned char* raw = smem + sp.raw_off + s * sp.raw_stage_bytes + h * N * 4 + g * 16;
        uint32_t w[4][9];
#pragma unroll
        for (int d = 0; d < 9; ++d) {
          const float4 v = *reinterpret_cast<const float4*>(raw + d * 2 * N * 4);
          const float vv[4] = {v.x, v.y, v.z, v.w};
          amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            // xh = x truncated to 11 significand bits (exactly representable in fp16 inside its normal
            // range), xl = x - xh exact in fp32 and rounded to fp16
            const float xh = __uint_as_float(__float_as_uint(vv[i]) & 0xffffe000u);
            w[i][d] = pack_h2(vv[i] - xh, xh);
          }
        }
        unsigned char* bt = smem + sp.b_off + (t * 2 + h) * sp.b_pillar_bytes;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = 4 * g + i;
          unsigned char* row = bt + (n >> 3) * kSboB + (n & 7) * 16;      // + c*128 per 8-wide k chunk
          *reinterpret_cast<uint4*>(row + 0 * kLboB) = make_uint4(w[i][0], w[i][1], w[i][2], w[i][3]);
          *reinterpret_cast<uint4*>(row + 1 * kLboB) = make_uint4(w[i][4], w[i][5], w[i][6], w[i][7]);
          *reinterpret_cast<uint4*>(row + 2 * kLboB) =
              make_uint4(w[i][8], lo_halves(w[i][0], w[i][1]), lo_halves(w[i][2], w[i][3]), lo_halves(w[i][4], w[i][5]));
          *reinterpret_cast<uint4*>(row + 3 * kLboB) =
              make_uint4(lo_halves(w[i][6], w[i][7]), (w[i][8] & 0xffffu) | 0x3c000000u, 0x00003c00u, 0u);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&b_full[t]);
        mbar_arrive(&raw_empty[s]);
      }
      s += 2; if (s >= R) { s -= R; ph_raw ^= 1u; }
      t += 2; if (t >= kBStages) { t -= kBStages; ph_b ^= 1u; }
    }
    if (!(amax < 32768.f)) atomicOr(range_flag, 1);     // also catches NaN / Inf
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
    // fixed-order combination of the four epilogue sets -> deterministic per-CTA partials
    const int qq = threadIdx.x >> 6, cc = threadIdx.x & 63;
    double v = 0.0;
    for (int kq = 0; kq < 4; ++kq) v += s_stat[(kq * 2 + qq) * 64 + cc];
    partials[((size_t)blockIdx.x * 2 + qq) * 64 + cc] = v;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

}  // namespace tch

void* tensor_map_encode_fn() { return (void*)tch::encode_fn(); }    // shared with loss.cu

extern int g_opt_pfn_tc_debug;
extern int g_opt_pfn_tc_timing;
long long* tc_prof_ptr();   // pfn_tc.cu

bool pfn_tc16_supported(int D, int N, int C, int P, const void* x) {
  return D == 9 && C == 64 && N >= 16 && N <= 256 && (N % 8) == 0 && (P % 2) == 0 && ((uintptr_t)x % 16) == 0 &&
         tch::encode_fn() != nullptr;
}

// statistics + per-pillar extremes of x [B,9,P,N] -> ext [B*P][2][64]
int launch_stats_tc16(const float* d_x, int B, int P, int N, const float* w, const float* bias,
                      const float* bn_w, int training, float* ext, double* partials, int nblocks,
                      int* range_flag, cudaStream_t st) {
  const tch::Smem sp = tch::smem_plan(N);
  if (sp.raw_stages < 2) return PP_ERR_UNSUPPORTED;
  CUtensorMap tmap;
  const cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)P, (cuuint64_t)B * 9};
  const cuuint64_t gstride[2] = {(cuuint64_t)N * 4, (cuuint64_t)P * N * 4};
  const cuuint32_t box[3] = {(cuuint32_t)N, 2u, 9u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  if (tch::encode_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(d_x), gdim, gstride, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return PP_ERR_UNSUPPORTED;
  long long* prof = g_opt_pfn_tc_timing ? tc_prof_ptr() : nullptr;
  PP_CUDA(cudaMemsetAsync(range_flag, 0, sizeof(int), st));
#define PP_TC16(TR)                                                                                               \
  do {                                                                                                            \
    PP_CUDA(cudaFuncSetAttribute(tch::k_pfn_stats_tc<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.total)); \
    PP_KERNEL("k_pfn_stats_tc", st,                                                                               \
              (tch::k_pfn_stats_tc<TR><<<nblocks, tch::kThreads, sp.total, st>>>(                                 \
                  tmap, B, P, N, w, bias, bn_w, ext, partials, range_flag, g_opt_pfn_tc_debug, prof)));          \
  } while (0)
  if (training) PP_TC16(true); else PP_TC16(false);
#undef PP_TC16
  return PP_OK;
}

}  // namespace pp
