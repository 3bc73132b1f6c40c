// pfn_tc.cu -- K2 statistics pass on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Why tensor cores: on CUDA cores the 9->64 contraction of PPFeatureNet is FP32-issue-bound, not
// HBM-bound (ncu r1a: 856 M warp instructions per batch-4 launch, DRAM 7 % of peak), which is the
// case BASELINE.json:north_star reserves tensor cores for.
//
// Per pillar row (b,p) the kernel computes  Y^T[c, n] = sum_k A'[c,k] * B'[k,n]  with
//   M = 64 channels, N = the N point slots of the pillar (one UMMA tile, N % 8 == 0, N <= 256),
//   K = 4 k-steps of 8 (kind::tf32), fp32 accumulation in TMEM.
// 1e-5 parity rules out plain TF32, so x and W are split  v = v_hi + v_lo  (both exactly
// representable in tf32) and the three significant products are laid out along K:
//   B' k     0.. 8  x_hi[0..8]      k     9..17  x_lo[0..8]    k   18  x_hi[8] (copy)
//            19, 20  1.0 (bias rows) k    21..23  0
//   k-step 0: rows 0-7   x A cols  W_hi[0..7]
//   k-step 1: rows 8-15  x A cols [W_hi[8], W_hi[0..6]]             (x_hi[8], x_lo[0..6])
//   k-step 2: rows 16-23 x A cols [W_hi[7], W_hi[8], W_lo[8], b_hi, b_lo, 0, 0, 0]
//   k-step 3: rows 0-7   x A cols  W_lo[0..7]                        (re-reads the x_hi rows)
// => y = W_hi x_hi + W_hi x_lo + W_lo x_hi + b  (dropped term W_lo x_lo ~ 2^-22 |w||x|).
// A' rows are pre-multiplied by s_c = sign(gamma_c): BN(relu(.)) is monotone increasing in y when
// gamma >= 0 and decreasing otherwise, so only max_n(s_c * y) has to be tracked per channel.
//
// Orientation: channels on the M axis (TMEM lanes), point slots on the N axis (TMEM columns), so
// the max / sum over the N slots is an in-thread reduction over registers filled by tcgen05.ld
// -- no shuffles.  M = 64 in cta_group::1 occupies lanes 0-15 of each 32-lane sub-partition; a
// second pillar's accumulator is interleaved at lane offset 16, so a 32x32b TMEM load hands every
// thread of the four epilogue warps one (pillar, channel) row.
//
// Warp roles (512 threads, one CTA per SM, persistent over pillar pairs):
//   warp 0      TMA producer: 9 bulk copies (one per feature, both pillars of the pair) into a 3-stage ring
//   warp 1      TMEM allocation + MMA issuer (one elected lane issues 8 tcgen05.mma per pair)
//   warps 4-7   epilogue group 0 (even pairs, accumulator buffer 0): tcgen05.ld -> max / sum relu /
//               sum relu^2 -> ext + fp64 partial sums;  warps 12-15: group 1 (odd pairs, buffer 1)
//   warps 8-11  converters: raw fp32 rows -> per-slot (x_hi, x_lo) k-vectors in the UMMA K-major
//               no-swizzle layout (a register-level transpose; kind::tf32 was measured to return
//               zeros for an MN-major B operand on this part, so B' is stored K-major)
// Pipelines (mbarriers): raw ring full/empty, B' tile full/empty (x2), TMEM accumulator full/empty (x2).
#include "tc_common.cuh"

namespace pp {

namespace tc {

constexpr int kThreads = 512;
constexpr int kRawStages = 4;
constexpr int kBStages = 3;            // B' operand tiles (one pillar pair each) in flight between converters and MMA
constexpr int kABytes = 4 * 2048;      // 4 k-steps x (64 rows x 8 k) tf32
constexpr int kSboB = 784;             // bytes between 8-slot groups of B' (6 core matrices of 128 B + 16 pad: conflict-free STS.128)
constexpr int kLboB = 128;             // bytes between the two 4-wide k chunks of one k-step

using namespace tcx;

struct Smem {
  // byte offsets inside the dynamic shared memory block
  int a_off, raw_off, b_off, bar_off, total;
  int raw_stage_bytes, b_pillar_bytes, lbo_b;
};

__host__ __device__ inline Smem smem_plan(int N) {
  Smem s;
  s.a_off = 0;
  s.raw_off = kABytes;
  s.raw_stage_bytes = 2 * 9 * N * 4;
  s.b_off = s.raw_off + kRawStages * s.raw_stage_bytes;
  s.lbo_b = kLboB;
  s.b_pillar_bytes = (N / 8) * kSboB;
  s.bar_off = s.b_off + kBStages * 2 * s.b_pillar_bytes;
  s.bar_off = (s.bar_off + 15) & ~15;
  s.total = s.bar_off + (2 * kRawStages + 2 * kBStages + 4) * 8 + 16;
  return s;
}

template <bool TRAIN>
__global__ void __launch_bounds__(kThreads, 1)
k_pfn_stats_tc(const float* __restrict__ x, int B, int P, int N, const float* __restrict__ conv_w,
               const float* __restrict__ conv_b, const float* __restrict__ bn_w,
               float* __restrict__ ext, double* __restrict__ partials, int dbg, long long* __restrict__ prof,
               const int* __restrict__ run_flag) {
  // fallback use: the fp16 kernel (pfn_tc16.cu) raises *run_flag when an input is outside the fp16 range
  if (run_flag != nullptr && *run_flag == 0) return;
  extern __shared__ __align__(128) unsigned char smem[];
  const Smem sp = smem_plan(N);
  const int warp = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bar_off);
  uint64_t* raw_full = bars;
  uint64_t* raw_empty = bars + kRawStages;
  uint64_t* b_full = bars + 2 * kRawStages;       // [kBStages]
  uint64_t* b_empty = b_full + kBStages;           // [kBStages]
  uint64_t* acc_full = b_empty + kBStages;         // [2]
  uint64_t* acc_empty = acc_full + 2;              // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);
  __shared__ double s_stat[2][2][64];      // [epilogue group][sum, sum of squares][channel]

  // all loop counters are 32-bit and advance incrementally: the producer and issuer are single
  // threads, 64-bit divisions there would sit on the critical path (host guarantees B*P < 2^31)
  const int rows = B * P;
  const int pairs = (rows + 1) / 2;
  const int my_pairs = (int)blockIdx.x < pairs ? (pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const size_t PN = (size_t)P * N;
  const int G4 = N / 4;

  // ---- one-time setup -------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRawStages; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 4); }
    for (int i = 0; i < kBStages; ++i) { mbar_init(&b_full[i], 4); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_barrier_init();
  }
  // A' tile: 64 x 32 tf32, K-major, no swizzle: core matrix = 8 rows x 16 B, k-chunk stride 128 B,
  // row-group stride 256 B, k-step stride 2048 B
  for (int idx = threadIdx.x; idx < 64 * 32; idx += kThreads) {
    const int m = idx >> 5, kk = idx & 31, j = kk >> 3, k = kk & 7;
    const float sgn = bn_w[m] < 0.f ? -1.f : 1.f;
    auto whi = [&](int d) { return to_tf32(conv_w[m * 9 + d]); };
    auto wlo = [&](int d) { const float w = conv_w[m * 9 + d]; return to_tf32(w - to_tf32(w)); };
    float v = 0.f;
    if (j == 0) v = whi(k);
    else if (j == 1) v = (k == 0) ? whi(8) : whi(k - 1);
    else if (j == 2) {
      const float bb = conv_b[m];
      if (k == 0) v = whi(7);
      else if (k == 1) v = whi(8);
      else if (k == 2) v = wlo(8);
      else if (k == 3) v = to_tf32(bb);
      else if (k == 4) v = to_tf32(bb - to_tf32(bb));
    } else v = wlo(k);
    *reinterpret_cast<float*>(smem + sp.a_off + j * 2048 + (m >> 3) * 256 + (k >> 2) * 128 + (m & 7) * 16 + (k & 3) * 4) = sgn * v;
  }
  fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
  if (warp == 1) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const bool pon = prof != nullptr && blockIdx.x == 0;
  long long pacc[4] = {0, 0, 0, 0};
  const long long prole0 = pon ? clock64() : 0;

  // ---- roles ----------------------------------------------------------------------------------
  if (warp == 0) {
    // ===== TMA producer =====
    // The two pillars of a pair are adjacent in memory (rows p and p+1 of the same feature plane), so
    // one bulk copy per feature moves both: 9 copies of 2*N*4 bytes, issued by 9 lanes at once.
    // Raw stage layout: [d][h][N].
    {
      int s = 0;
      uint32_t ph = 1;                                  // parity to wait on raw_empty (first pass falls through)
      int r0 = 2 * (int)blockIdx.x;                     // first row of the current pair
      int b0 = r0 / P, p0 = r0 - b0 * P;                // (sweep, pillar) of that row, advanced incrementally
      const int step = 2 * (int)gridDim.x;
      for (int it = 0; it < my_pairs; ++it) {
        mbar_wait_t(&raw_empty[s], ph, pon, pacc[0]);
        if (lane == 0) mbar_expect_tx(&raw_full[s], (uint32_t)sp.raw_stage_bytes);
        __syncwarp();
        unsigned char* dst = smem + sp.raw_off + s * sp.raw_stage_bytes;
        const bool contiguous = (p0 + 1 < P) && (r0 + 1 < rows);
        if (contiguous) {
          if (lane < 9) {
            const float* src = x + ((size_t)b0 * 9 + lane) * PN + (size_t)p0 * N;
            bulk_g2s(dst + lane * 2 * N * 4, src, (uint32_t)(2 * N * 4), &raw_full[s]);
          }
        } else if (lane < 18) {
          const int d = (int)lane >> 1, h = (int)lane & 1;
          int bb = b0, pp = p0 + h;
          if (pp >= P) { pp -= P; ++bb; }
          if (r0 + h >= rows) { bb = b0; pp = p0; }     // odd row count: duplicate the last row
          const float* src = x + ((size_t)bb * 9 + d) * PN + (size_t)pp * N;
          bulk_g2s(dst + (d * 2 + h) * N * 4, src, (uint32_t)(N * 4), &raw_full[s]);
        }
        r0 += step; p0 += step;
        while (p0 >= P) { p0 -= P; ++b0; }
        if (++s == kRawStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): fp32 accumulate,
      // A = B = tf32, both K-major, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) |
                             ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
      const uint32_t a_addr = smem_u32(smem + sp.a_off);
      int bs = 0;
      uint32_t bph = 0;
      for (int it = 0; it < my_pairs; ++it) {
        const int t = it & 1;                          // accumulator buffer
        const uint32_t n = (uint32_t)(it >> 1);
        mbar_wait_t(&b_full[bs], bph, pon, pacc[0]);
        mbar_wait_t(&acc_empty[t], (n & 1u) ^ 1u, pon, pacc[1]);
        tc_fence_after();
        for (int h = 0; h < 2 && !(dbg & 2); ++h) {
          const uint32_t b_addr = smem_u32(smem + sp.b_off + (bs * 2 + h) * sp.b_pillar_bytes);
          const uint32_t d_tmem = tmem_base + ((uint32_t)(h * 16) << 16) + (uint32_t)(t * 256);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kg = (j == 3) ? 0 : j;
            const uint64_t ad = smem_desc(a_addr + j * 2048, 128, 256);
            const uint64_t bd = smem_desc(b_addr + kg * 2 * kLboB, kLboB, kSboB);
            umma_tf32(d_tmem, ad, bd, idesc, j > 0 ? 1u : 0u);
          }
        }
        umma_commit(&b_empty[bs]);    // B' tile free once these MMAs have read it
        umma_commit(&acc_full[t]);    // accumulators ready for the epilogue
        if (++bs == kBStages) { bs = 0; bph ^= 1u; }
      }
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 12) {
    // ===== epilogue: one (pillar-of-pair, channel) row per thread; group e owns accumulator buffer e =====
    const int e = warp >= 12 ? 1 : 0;
    const int q = warp & 3;
    const int h = lane >> 4;
    const int c = 16 * q + (int)(lane & 15u);
    const float sgn = bn_w[c] < 0.f ? -1.f : 1.f;
    double accS = 0.0, accQ = 0.0;
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(e * 256);
    const int nfull = N >> 5;              // 32-column chunks
    for (int it = e; it < my_pairs; it += 2) {
      const uint32_t n = (uint32_t)(it >> 1);
      mbar_wait_t(&acc_full[e], n & 1u, pon, pacc[0]);
      tc_fence_after();
      // independent accumulator sets break the dependent chains; sums are kept as packed fp32 pairs
      float mx[2] = {-INFINITY, -INFINITY};
      unsigned long long S[2] = {0ull, 0ull}, Q[2] = {0ull, 0ull};
      // 2*relu(s*y) = s*y + |y| is one FFMA (FMA pipe; the ALU pipe that FMNMX runs on is half rate:
      // doing the relu with FMNMX measured 434 us vs 307 us), sums are packed FADD2 / FFMA2
      auto consume = [&](const uint32_t* v, int cnt) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          if (i < cnt) {
            const float y0 = __uint_as_float(v[i]), y1 = __uint_as_float(v[i + 1]);
            mx[(i >> 1) & 1] = fmaxf(mx[(i >> 1) & 1], fmaxf(y0, y1));
            if (TRAIN) {
              const float t0 = fmaf(sgn, y0, fabsf(y0));
              const float t1 = fmaf(sgn, y1, fabsf(y1));
              acc_pair(S[(i >> 1) & 1], Q[(i >> 1) & 1], t0, t1);
            }
          }
        }
      };
      // software-pipelined TMEM reads: chunk k+1 is in flight while chunk k is reduced
      uint32_t va[32], vb[32];
      if (nfull > 0) { PP_TMEM_LD32(taddr, va); }
      for (int k = 0; k < ((dbg & 4) ? 0 : nfull); k += 2) {
        tmem_ld_wait();
        if (k + 1 < nfull) { PP_TMEM_LD32(taddr + 32 * (k + 1), vb); }
        consume(va, 32);
        if (k + 1 < nfull) {
          tmem_ld_wait();
          if (k + 2 < nfull) { PP_TMEM_LD32(taddr + 32 * (k + 2), va); }
          consume(vb, 32);
        }
      }
      for (int col = nfull * 32; col + 8 <= N; col += 8) {
        uint32_t v8[8];
        PP_TMEM_LD8(taddr + col, v8);
        tmem_ld_wait();
        consume(v8, 8);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[e]);
      const int r = 2 * ((int)blockIdx.x + it * (int)gridDim.x) + h;
      if (r < rows) {
        const float m = fmaxf(mx[0], mx[1]);
        const float val = sgn * m;               // max_n y when gamma >= 0, min_n y otherwise
        float* eo = ext + (size_t)r * 128 + c;
        eo[0] = val;
        eo[64] = val;
        if (TRAIN) {
          accS += 0.5 * ((double)pair_sum(S[0]) + (double)pair_sum(S[1]));
          accQ += 0.25 * ((double)pair_sum(Q[0]) + (double)pair_sum(Q[1]));
        }
      }
    }
    if (TRAIN) {
      accS += __shfl_xor_sync(0xffffffffu, accS, 16);
      accQ += __shfl_xor_sync(0xffffffffu, accQ, 16);
      if (lane < 16) {
        s_stat[e][0][c] = accS;
        s_stat[e][1][c] = accQ;
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ===== converters: raw fp32 rows -> x_hi / x_lo rows in the UMMA MN-major layout =====
    const int ct = threadIdx.x - 8 * 32;    // 0..127
    const int h = ct >> 6;                  // pillar of the pair
    const int g = ct & 63;                  // 4-slot group
    int s = 0, t = 0;
    uint32_t ph_raw = 0, ph_b = 1;
    for (int it = 0; it < my_pairs; ++it) {
      mbar_wait_t(&raw_full[s], ph_raw, pon, pacc[0]);
      mbar_wait_t(&b_empty[t], ph_b, pon, pacc[1]);
      if (g < G4 && !(dbg & 1)) {
        const unsigned char* raw = smem + sp.raw_off + s * sp.raw_stage_bytes + h * N * 4 + g * 16;
        float hi[9][4], lo[9][4];
#pragma unroll
        for (int d = 0; d < 9; ++d) {
          const float4 v = *reinterpret_cast<const float4*>(raw + d * 2 * N * 4);
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            // hi = x truncated to tf32 (what the tensor core would do with the raw word); lo = x - hi is
            // exact in fp32 (<= 13 significant bits) and is truncated to tf32 by the hardware on read
            hi[d][i] = __uint_as_float(__float_as_uint(vv[i]) & 0xffffe000u);
            lo[d][i] = vv[i] - hi[d][i];
          }
        }
        unsigned char* bt = smem + sp.b_off + (t * 2 + h) * sp.b_pillar_bytes;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = 4 * g + i;
          unsigned char* row = bt + (n >> 3) * kSboB + (n & 7) * 16;      // + kc*128 per 4-wide k chunk
          *reinterpret_cast<float4*>(row + 0 * kLboB) = make_float4(hi[0][i], hi[1][i], hi[2][i], hi[3][i]);
          *reinterpret_cast<float4*>(row + 1 * kLboB) = make_float4(hi[4][i], hi[5][i], hi[6][i], hi[7][i]);
          *reinterpret_cast<float4*>(row + 2 * kLboB) = make_float4(hi[8][i], lo[0][i], lo[1][i], lo[2][i]);
          *reinterpret_cast<float4*>(row + 3 * kLboB) = make_float4(lo[3][i], lo[4][i], lo[5][i], lo[6][i]);
          *reinterpret_cast<float4*>(row + 4 * kLboB) = make_float4(lo[7][i], lo[8][i], hi[8][i], 1.f);
          *reinterpret_cast<float4*>(row + 5 * kLboB) = make_float4(1.f, 0.f, 0.f, 0.f);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&b_full[t]);
        mbar_arrive(&raw_empty[s]);
      }
      if (++s == kRawStages) { s = 0; ph_raw ^= 1u; }
      if (++t == kBStages) { t = 0; ph_b ^= 1u; }
    }
  }

  if (pon && lane == 0) {
    pacc[3] = clock64() - prole0;
    for (int k = 0; k < 4; ++k) prof[warp * 4 + k] = pacc[k];
  }
  // ---- teardown ---------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (TRAIN && threadIdx.x < 128) {
    // fixed-order combination of the two epilogue groups -> deterministic per-CTA partials
    const int qq = threadIdx.x >> 6, cc = threadIdx.x & 63;
    partials[((size_t)blockIdx.x * 2 + qq) * 64 + cc] = s_stat[0][qq][cc] + s_stat[1][qq][cc];
  }
}

}  // namespace tc

extern int g_opt_pfn_tc_debug;
extern int g_opt_pfn_tc_timing;
__device__ long long g_tc_prof[128];

bool pfn_tc_supported(int D, int N, int C, const void* x) {
  return D == 9 && C == 64 && N >= 8 && N <= 256 && (N % 8) == 0 && ((uintptr_t)x % 16) == 0;
}

int launch_stats_tc(const float* d_x, int B, int P, int N, const float* w, const float* bias,
                    const float* bn_w, int training, float* ext, double* partials, int nblocks,
                    const int* run_flag, cudaStream_t st) {
  const tc::Smem sp = tc::smem_plan(N);
  long long* prof = nullptr;
  if (g_opt_pfn_tc_timing) PP_CUDA(cudaGetSymbolAddress((void**)&prof, g_tc_prof));
  if (training) {
    PP_CUDA(cudaFuncSetAttribute(tc::k_pfn_stats_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.total));
    PP_KERNEL("k_pfn_stats_tf32", st,
              tc::k_pfn_stats_tc<true><<<nblocks, tc::kThreads, sp.total, st>>>(d_x, B, P, N, w, bias, bn_w, ext, partials, g_opt_pfn_tc_debug, prof, run_flag));
  } else {
    PP_CUDA(cudaFuncSetAttribute(tc::k_pfn_stats_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.total));
    PP_KERNEL("k_pfn_stats_tf32", st,
              tc::k_pfn_stats_tc<false><<<nblocks, tc::kThreads, sp.total, st>>>(d_x, B, P, N, w, bias, bn_w, ext, partials, g_opt_pfn_tc_debug, prof, run_flag));
  }
  return PP_OK;
}

long long* tc_prof_ptr() {
  long long* p = nullptr;
  return cudaGetSymbolAddress((void**)&p, g_tc_prof) == cudaSuccess ? p : nullptr;
}

int read_tc_prof(long long* out64) {
  PP_CUDA(cudaDeviceSynchronize());
  PP_CUDA(cudaMemcpyFromSymbol(out64, g_tc_prof, sizeof(long long) * 128));
  return PP_OK;
}

}  // namespace pp
