// pfn.cu -- K2: PPFeatureNet (1x1 conv 9->C, ReLU, BatchNorm, max over N) and PPScatter for
// sm_100a.  Replaces model/model.py:31-40 and :53-62 of the reference.
//
// Algebra.  ReLU and BatchNorm are per-channel monotone maps, so
//     max_n BN(relu(y_n)) = BN(relu(max_n y_n))   when gamma*invstd >= 0
//                         = BN(relu(min_n y_n))   otherwise,
// with y = W x + b.  One pass over x therefore only has to keep, per (b,p,c), the running max and
// min of y and, in training mode, the per-channel sums of relu(y) and relu(y)^2 over (B,P,N);
// the [B,C,P,N] intermediate the reference materialises four times never exists.
//
// Kernels
//   k_pfn_stats   persistent, one CTA per SM, one pillar row per warp at a time.  The nine
//                 feature rows of a pillar ([9][N] floats) are fetched into a per-warp shared
//                 memory ring by 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), the
//                 weights live in registers (2 channels per lane), x is read by warp-wide
//                 broadcast LDS.128.  Slots whose nine features are all zero give y == b exactly
//                 and are only counted (the tensor is ~98.7 % padding when no data_mean is used).
//                 Output: ext[b,p,{max,min},c] (channel-contiguous, coalesced) + per-CTA partial
//                 sums in fp64 (fixed reduction order => run-to-run deterministic).
//   k_bn_finalize per-channel batch statistics -> affine (mean, scale, beta, use-min), running
//                 statistics update exactly as nn.BatchNorm2d.
//   k_pfn_out     ext -> out[B,C,P] (transpose through shared memory).
//   k_build_map   inds -> cell->pillar map;  k_canvas  dense, fully coalesced canvas write
//                 (zeros + gathered pillars), from ext (+affine) or from a [B,C,P] feature tensor.
#include <type_traits>

#include "internal.cuh"

namespace pp {

constexpr int kD = 9;
constexpr int kWarps = 8;          // warps per CTA in k_pfn_stats
constexpr int kMaxChunk = 256;     // slots per shared-memory tile row
constexpr int kStages = 3;

// ---- mbarrier / bulk-copy PTX (sm_90+ / sm_100a) -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

struct Affine {   // per channel, produced by k_bn_finalize
  float mean, scale, beta, use_min;
};

// ------------------------------------------------------------------------------------------
// work item = (row r = b*P+p, chunk k of the N axis); rows are dealt round-robin to warps.
template <int CPL, bool TRAIN>
__global__ void __launch_bounds__(kWarps * 32, 1)
k_pfn_stats(const float* __restrict__ x, int B, int P, int N, int chunk, int nchunks, bool use_bulk,
            const float* __restrict__ conv_w, const float* __restrict__ conv_b,
            float* __restrict__ ext, double* __restrict__ partials) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int C = CPL * 32;
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  const int tile_floats = kD * chunk;
  float* tiles = reinterpret_cast<float*>(smem_raw) + (size_t)warp * kStages * tile_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kWarps * kStages * tile_floats * 4) +
                   warp * kStages;
  __shared__ double s_red[kWarps][2][64];

  // weights of this lane's channels in registers
  float w[CPL][kD], bias[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = CPL * lane + j;
    bias[j] = conv_b[c];
#pragma unroll
    for (int d = 0; d < kD; ++d) w[j][d] = conv_w[c * kD + d];
  }

  if (use_bulk && lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncwarp();

  const long long rows = (long long)B * P;
  const long long gw = (long long)blockIdx.x * kWarps + warp;   // global warp id
  const long long nw = (long long)gridDim.x * kWarps;
  const long long my_rows = gw < rows ? (rows - gw + nw - 1) / nw : 0;
  const long long my_items = my_rows * nchunks;
  const size_t PN = (size_t)P * N;

  auto issue = [&](long long item) {
    const long long r = gw + (item / nchunks) * nw;
    const int k = (int)(item % nchunks);
    const int b = (int)(r / P), p = (int)(r % P);
    const int n0 = k * chunk;
    const int len = min(chunk, N - n0);
    const int stage = (int)(item % kStages);
    float* dst = tiles + (size_t)stage * tile_floats;
    const float* src = x + (size_t)b * kD * PN + (size_t)p * N + n0;
    if (use_bulk) {
      if (lane == 0) {
        mbar_expect_tx(&bars[stage], (uint32_t)(kD * len * 4));
#pragma unroll
        for (int d = 0; d < kD; ++d) bulk_g2s(dst + d * chunk, src + d * PN, (uint32_t)(len * 4), &bars[stage]);
      }
    } else {
      for (int d = 0; d < kD; ++d)
        for (int n = lane; n < len; n += 32) dst[d * chunk + n] = __ldg(src + d * PN + n);
    }
  };

  double accS[CPL], accQ[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { accS[j] = 0.0; accQ[j] = 0.0; }

  if (use_bulk) {
    for (long long it = 0; it < my_items && it < kStages; ++it) issue(it);
  }

  float mx[CPL], mn[CPL], rs[CPL], rq[CPL];
  int nz = 0;
  for (long long it = 0; it < my_items; ++it) {
    const int k = (int)(it % nchunks);
    const int stage = (int)(it % kStages);
    const int n0 = k * chunk;
    const int len = min(chunk, N - n0);
    if (k == 0) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) { mx[j] = -INFINITY; mn[j] = INFINITY; rs[j] = 0.f; rq[j] = 0.f; }
      nz = 0;
    }
    if (use_bulk) {
      mbar_wait(&bars[stage], (uint32_t)((it / kStages) & 1));
    } else {
      __syncwarp();
      issue(it);
      __syncwarp();
    }
    const float* t = tiles + (size_t)stage * tile_floats;

    auto slot = [&](const float (&xv)[kD]) {
      unsigned bits = 0;
#pragma unroll
      for (int d = 0; d < kD; ++d) bits |= __float_as_uint(xv[d]);
      if ((bits & 0x7fffffffu) == 0u) { ++nz; return; }   // y == b exactly; accounted per row
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        float y = bias[j];
#pragma unroll
        for (int d = 0; d < kD; ++d) y = fmaf(w[j][d], xv[d], y);
        mx[j] = fmaxf(mx[j], y);
        mn[j] = fminf(mn[j], y);
        if (TRAIN) {
          const float r = fmaxf(y, 0.f);
          rs[j] += r;
          rq[j] = fmaf(r, r, rq[j]);
        }
      }
    };

    const int len4 = len & ~3;
    for (int n = 0; n < len4; n += 4) {
      float4 v[kD];
#pragma unroll
      for (int d = 0; d < kD; ++d) v[d] = *reinterpret_cast<const float4*>(t + d * chunk + n);
      {
        float xv[kD];
#pragma unroll
        for (int d = 0; d < kD; ++d) xv[d] = v[d].x;
        slot(xv);
#pragma unroll
        for (int d = 0; d < kD; ++d) xv[d] = v[d].y;
        slot(xv);
#pragma unroll
        for (int d = 0; d < kD; ++d) xv[d] = v[d].z;
        slot(xv);
#pragma unroll
        for (int d = 0; d < kD; ++d) xv[d] = v[d].w;
        slot(xv);
      }
    }
    for (int n = len4; n < len; ++n) {
      float xv[kD];
#pragma unroll
      for (int d = 0; d < kD; ++d) xv[d] = t[d * chunk + n];
      slot(xv);
    }

    __syncwarp();
    if (use_bulk && it + kStages < my_items) issue(it + kStages);

    if (k == nchunks - 1) {
      const long long r = gw + (it / nchunks) * nw;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        if (nz > 0) {
          mx[j] = fmaxf(mx[j], bias[j]);
          mn[j] = fminf(mn[j], bias[j]);
        }
        if (TRAIN) {
          const double rb = (double)fmaxf(bias[j], 0.f);
          accS[j] += (double)rs[j] + (double)nz * rb;
          accQ[j] += (double)rq[j] + (double)nz * rb * rb;
        }
      }
      float* e = ext + (size_t)r * 2 * C + CPL * lane;
      if (CPL == 2) {
        *reinterpret_cast<float2*>(e) = make_float2(mx[0], mx[CPL - 1]);
        *reinterpret_cast<float2*>(e + C) = make_float2(mn[0], mn[CPL - 1]);
      } else {
        e[0] = mx[0];
        e[C] = mn[0];
      }
    }
  }

  if (TRAIN) {
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      s_red[warp][0][CPL * lane + j] = accS[j];
      s_red[warp][1][CPL * lane + j] = accQ[j];
    }
    __syncthreads();
    if (threadIdx.x < 2 * C) {
      const int q = threadIdx.x / C, c = threadIdx.x % C;
      double v = 0.0;
      for (int wv = 0; wv < kWarps; ++wv) v += s_red[wv][q][c];
      partials[((size_t)blockIdx.x * 2 + q) * C + c] = v;
    }
  }
}

// 64 channels x 8 segments: each thread sums a contiguous run of per-CTA partials, the eight
// segment sums are combined in fixed order => deterministic, ~8x shorter dependency chain.
// Sparse path: sum = mult * (sum of partials) + (sum of partials2) + base, where partials come from
// the padding pass (every padding value occurs once per sweep: mult = B), partials2 from k_pfn_real
// (real slot minus the padding value it replaces) and base = pad_count * relu(bias) when there is no
// data_mean (all padding slots are exactly zero, y == bias).
struct SparseFinalize {
  double mult;
  const double* partials2;
  int nparts2;
  const float* conv_b;     // non-null: analytic padding base
  double pad_count;
  const int* range_flag;   // non-null: raise PP_STATUS_RANGE in *status when set
  int* status;
  // padding pass of pfn_pad.cu: its partials are sum |y| and sum y|y|; with the input moments mom[10][10]
  // (x_0..x_8, 1) they give sum relu(y) = (sum y + sum |y|) / 2 and sum relu(y)^2 = (sum y^2 + sum y|y|) / 2
  const double* mom;       // non-null: add mom_mult * (linear / quadratic form of the moments)
  const float* mom_w;      // conv weight [C,9]
  const float* mom_b;      // conv bias [C]
  double mom_mult;
  const int* range_flag2;  // second range flag (the prepared operand's)
};

constexpr int kFinSegs = 16;
__global__ void __launch_bounds__(64 * kFinSegs) k_bn_finalize(int C, int nparts, double count, int training,
                                                     float momentum, float eps, SparseFinalize sf,
                                                     const double* __restrict__ partials,
                                                     const float* __restrict__ bn_w,
                                                     const float* __restrict__ bn_b,
                                                     float* __restrict__ running_mean,
                                                     float* __restrict__ running_var,
                                                     long long* __restrict__ num_batches_tracked,
                                                     Affine* __restrict__ affine) {
  __shared__ double s_part[2][kFinSegs][64];
  const int c = threadIdx.x & 63, seg = threadIdx.x >> 6;
  // a data_mean value or weight outside the fp16 range of the padding pass: the statistics of this pass are not
  // valid -- report it and leave the running statistics untouched (the canvas of this call is invalid)
  const bool out_of_range = (sf.range_flag != nullptr && *sf.range_flag != 0) || (sf.range_flag2 != nullptr && *sf.range_flag2 != 0);
  if (threadIdx.x == 0 && out_of_range) atomicOr(sf.status, PP_STATUS_RANGE);
  if (training && c < C) {
    const int per = (nparts + kFinSegs - 1) / kFinSegs;
    const int k0 = seg * per, k1 = min(nparts, k0 + per);
    // four independent accumulators (fixed combination order => still deterministic): the loads of a
    // segment are in flight together instead of one L2 round trip per partial
    double S4[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, Q4[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int k = k0;
    for (; k + 8 <= k1; k += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        S4[j] += partials[((size_t)(k + j) * 2 + 0) * C + c];
        Q4[j] += partials[((size_t)(k + j) * 2 + 1) * C + c];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (k + j < k1) {
        S4[j] += partials[((size_t)(k + j) * 2 + 0) * C + c];
        Q4[j] += partials[((size_t)(k + j) * 2 + 1) * C + c];
      }
    }
    double S = (((S4[0] + S4[1]) + (S4[2] + S4[3])) + ((S4[4] + S4[5]) + (S4[6] + S4[7]))) * sf.mult;
    double Q = (((Q4[0] + Q4[1]) + (Q4[2] + Q4[3])) + ((Q4[4] + Q4[5]) + (Q4[6] + Q4[7]))) * sf.mult;
    if (sf.partials2 != nullptr) {
      const int per2 = (sf.nparts2 + kFinSegs - 1) / kFinSegs;
      const int j0 = seg * per2, j1 = min(sf.nparts2, j0 + per2);
      double S2[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, Q2[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      int j = j0;
      for (; j + 8 <= j1; j += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          S2[u] += sf.partials2[((size_t)(j + u) * 2 + 0) * C + c];
          Q2[u] += sf.partials2[((size_t)(j + u) * 2 + 1) * C + c];
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (j + u < j1) {
          S2[u] += sf.partials2[((size_t)(j + u) * 2 + 0) * C + c];
          Q2[u] += sf.partials2[((size_t)(j + u) * 2 + 1) * C + c];
        }
      }
      S += ((S2[0] + S2[1]) + (S2[2] + S2[3])) + ((S2[4] + S2[5]) + (S2[6] + S2[7]));
      Q += ((Q2[0] + Q2[1]) + (Q2[2] + Q2[3])) + ((Q2[4] + Q2[5]) + (Q2[6] + Q2[7]));
    }
    if (sf.conv_b != nullptr && seg == 0) {
      const double rb = (double)fmaxf(sf.conv_b[c], 0.f);
      S += sf.pad_count * rb;
      Q += sf.pad_count * rb * rb;
    }
    if (sf.mom != nullptr && seg == 0) {
      double wt[10];
      for (int d = 0; d < 9; ++d) wt[d] = (double)sf.mom_w[c * 9 + d];
      wt[9] = (double)sf.mom_b[c];
      double lin = 0.0, quad = 0.0;
      for (int d = 0; d < 10; ++d) {
        lin += wt[d] * sf.mom[d * 10 + 9];
        double row = 0.0;
        for (int e2 = 0; e2 < 10; ++e2) row += wt[e2] * sf.mom[d * 10 + e2];
        quad += wt[d] * row;
      }
      S += sf.mom_mult * lin;
      Q += sf.mom_mult * quad;
    }
    s_part[0][seg][c] = S;
    s_part[1][seg][c] = Q;
  }
  __syncthreads();
  if (seg != 0 || c >= C) return;
  double mean, var;
  if (training) {
    double S = 0.0, Q = 0.0;
    for (int k = 0; k < kFinSegs; ++k) { S += s_part[0][k][c]; Q += s_part[1][k][c]; }
    mean = S / count;
    var = Q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double unbiased = count > 1.0 ? var * (count / (count - 1.0)) : var;
    if (!out_of_range) {
      running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mean);
      running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unbiased);
      if (c == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
    }
  } else {
    mean = (double)running_mean[c];
    var = (double)running_var[c];
  }
  const double invstd = 1.0 / sqrt(var + (double)eps);
  const double scale = (double)bn_w[c] * invstd;
  Affine a;
  a.mean = (float)mean;
  a.scale = (float)scale;
  a.beta = bn_b[c];
  a.use_min = scale < 0.0 ? 1.f : 0.f;
  affine[c] = a;
}

// ext holds two fields per (row, channel): {max_n y, min_n y} from the CUDA-core kernel, or the partial
// extremes of the two column halves from the tensor-core kernels (already sign-selected); in both
// cases the pre-activation that survives BN(relu(.)) + max_n is max(e0,e1) for gamma*invstd >= 0 and
// min(e0,e1) otherwise.
__device__ __forceinline__ float apply_affine(const Affine& a, float e0, float e1) {
  const float v = fmaxf(a.use_min != 0.f ? fminf(e0, e1) : fmaxf(e0, e1), 0.f);   // relu of the extreme pre-activation
  return fmaf(v - a.mean, a.scale, a.beta);
}

// ext [B*P, 2, C] -> out [B, C, P]
__global__ void __launch_bounds__(256) k_pfn_out(const float* __restrict__ ext,
                                                 const Affine* __restrict__ affine, int P, int C,
                                                 float* __restrict__ out) {
  __shared__ float tile[64][33];
  __shared__ Affine s_aff[64];
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  if (threadIdx.x < C) s_aff[threadIdx.x] = affine[threadIdx.x];
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * C; idx += 256) {
    const int pp = idx / C, c = idx % C;
    if (p0 + pp < P) {
      const float* e = ext + ((size_t)b * P + p0 + pp) * 2 * C;
      tile[c][pp] = apply_affine(s_aff[c], e[c], e[C + c]);
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * C; idx += 256) {
    const int c = idx / 32, pp = idx % 32;
    if (p0 + pp < P) out[((size_t)b * C + c) * P + p0 + pp] = tile[c][pp];
  }
}

__global__ void __launch_bounds__(256) k_build_map(const long long* __restrict__ inds, int B, int P,
                                                   int H, int W, int* __restrict__ map,
                                                   int* __restrict__ status) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * P) return;
  const long long* row = inds + i * 3;
  if (row[0] == 0) return;                       // model/model.py:56 nonzero(inds[:,:,0])
  const long long cx = row[1], cy = row[2];      // model/model.py:59-61: out[b,:,y_inds,x_inds]
  if (cx < 0 || cx >= W || cy < 0 || cy >= H) {
    atomicOr(status, PP_STATUS_BAD_INDEX);
    return;
  }
  const int b = (int)(i / P);
  atomicMax(&map[(size_t)b * H * W + cy * W + cx], (int)(i % P));
}

// Dense canvas write, warp-centric.  Work unit = (128 consecutive cells, kChanPerUnit channels): 4 cells
// per lane, so every store is a coalesced 512-byte row segment and no block-level synchronisation is
// needed.  ~95 % of the cells are empty but ~99.8 % of the 128-cell groups hold a pillar (~6 on
// average).  Per unit the occupied cells are listed, their channel values (ext row slice + BN affine)
// are staged into a small per-warp shared-memory tile by a flattened (cell, channel) loop, and the
// store stream splices them in with one LDS per occupied component; lanes that own no pillar only
// store zeros.  Gathering per (lane, channel) from global memory inside the store loop cost ~45
// instructions per channel per warp and was issue-bound (ncu r1h).  Units with more than kStage
// occupied cells take that gather path.
// Units are dealt round-robin to a grid that is exactly one resident wave (occupancy x SM count CTAs):
// with one unit of 64 channels per warp the 11264 units of the reference shape ran as 1.19 waves of
// the 9472 resident warps, i.e. two passes with the second almost empty (96 us for 369 MB); 16-channel
// units (45056) leave a 5 % imbalance.  A pure store stream with this access pattern reaches 6.2 TB/s
// (scripts/ubench/wr.cu).
//   FROM_EXT: source is ext[b*P+p][2][C] + affine (fused path); else feat[b][c][p] (PPScatter).
constexpr int kChanPerUnit = 16;
constexpr int kStage = 20;      // occupied cells staged per unit
// FROM_EXT: 0 = feature tensor, 2 = ext rows with two fields per (pillar, channel), 3 = sparse path (one field)
template <int FROM_EXT>
__global__ void __launch_bounds__(256) k_canvas(const float* __restrict__ src,
                                                const Affine* __restrict__ affine,
                                                const int* __restrict__ map, int map_bias, int B, int P, int C, int HW,
                                                bool vec_ok, float* __restrict__ canvas) {
  constexpr int kCells = 128;
  __shared__ Affine s_aff[64];
  __shared__ float s_tile[8][kStage][kChanPerUnit + 1];   // [warp][occupied cell][channel of the slice], padded
  __shared__ int s_slot[8][kStage];
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  if (FROM_EXT) {
    if (threadIdx.x < C) s_aff[threadIdx.x] = affine[threadIdx.x];
    __syncthreads();
  }
  const int ngroups = (HW + kCells - 1) / kCells;
  const int nq = (C + kChanPerUnit - 1) / kChanPerUnit;       // channel slices per group
  const long long units = (long long)B * ngroups * nq;
  const long long nwarps = (long long)gridDim.x * 8;
  for (long long u = (long long)blockIdx.x * 8 + warp; u < units; u += nwarps) {
    const int cq = (int)(u % nq);
    const long long t = u / nq;
    const int grp = (int)(t % ngroups);
    const int b = (int)(t / ngroups);
    const int c0 = cq * kChanPerUnit, c1 = min(C, c0 + kChanPerUnit);
    const int* mb = map + (size_t)b * HW;
    float* cb = canvas + (size_t)b * C * HW;
    auto value = [&](int slot, int c) -> float {
      if (FROM_EXT == 2) {
        const float* e = src + ((size_t)b * P + slot) * 2 * C;
        return apply_affine(s_aff[c], e[c], e[C + c]);
      }
      if (FROM_EXT == 3) {       // sparse path: one field per (pillar, channel), already the pillar's extreme
        const float* e = src + ((size_t)b * P + slot) * C;
        return apply_affine(s_aff[c], e[c], e[c]);
      }
      return src[((size_t)b * C + c) * P + slot];
    };
    const int cell0 = grp * kCells;
    if (vec_ok && cell0 + kCells <= HW) {
      int4 sl = __ldg(reinterpret_cast<const int4*>(mb + cell0) + lane);
      sl.x -= map_bias; sl.y -= map_bias; sl.z -= map_bias; sl.w -= map_bias;   // K1's map holds slot + 1 (0 = none)
      float4* o = reinterpret_cast<float4*>(cb + cell0) + lane;
      const size_t cs = (size_t)HW / 4;                 // float4 stride between channels
      const int mine = (sl.x >= 0) + (sl.y >= 0) + (sl.z >= 0) + (sl.w >= 0);
      int incl = mine;                                  // inclusive warp scan of the occupied counts
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += v;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (total == 0) {
#pragma unroll 8
        for (int c = c0; c < c1; ++c) __stcs(o + c * cs, z);
      } else if (total <= kStage) {
        int k0 = incl - mine;                           // first tile row of this lane's cells
        const int kx = k0; if (sl.x >= 0) s_slot[warp][k0++] = sl.x;
        const int ky = k0; if (sl.y >= 0) s_slot[warp][k0++] = sl.y;
        const int kz = k0; if (sl.z >= 0) s_slot[warp][k0++] = sl.z;
        const int kw = k0; if (sl.w >= 0) s_slot[warp][k0++] = sl.w;
        __syncwarp();
        const int nc = c1 - c0;
#pragma unroll 2
        for (int idx = (int)lane; idx < total * nc; idx += 32) {
          const int k = idx / nc, cc = idx - k * nc;
          s_tile[warp][k][cc] = value(s_slot[warp][k], c0 + cc);
        }
        __syncwarp();
        if (mine == 0) {
#pragma unroll 8
          for (int c = c0; c < c1; ++c) __stcs(o + c * cs, z);
        } else {
#pragma unroll 4
          for (int c = c0; c < c1; ++c) {
            float4 v = z;
            if (sl.x >= 0) v.x = s_tile[warp][kx][c - c0];
            if (sl.y >= 0) v.y = s_tile[warp][ky][c - c0];
            if (sl.z >= 0) v.z = s_tile[warp][kz][c - c0];
            if (sl.w >= 0) v.w = s_tile[warp][kw][c - c0];
            __stcs(o + c * cs, v);
          }
        }
        __syncwarp();                                   // the tile is reused by the next unit
      } else {
#pragma unroll 4
        for (int c = c0; c < c1; ++c) {
          float4 v = z;
          if (mine) {
            if (sl.x >= 0) v.x = value(sl.x, c);
            if (sl.y >= 0) v.y = value(sl.y, c);
            if (sl.z >= 0) v.z = value(sl.z, c);
            if (sl.w >= 0) v.w = value(sl.w, c);
          }
          __stcs(o + c * cs, v);
        }
      }
    } else {
      for (int j = lane; j < kCells && cell0 + j < HW; j += 32) {
        const int slot = mb[cell0 + j] - map_bias;
        for (int c = c0; c < c1; ++c) cb[(size_t)c * HW + cell0 + j] = slot >= 0 ? value(slot, c) : 0.f;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
struct PfnWs {
  float* ext;        // [B*P, 2, C]
  double* partials;  // [nblocks, 2, C]
  Affine* affine;    // [C]
  int* map;          // [B, H*W]
  int* flags;        // [0]: fp16 range guard of the tensor-core statistics kernel
};

template <class A>
static void pfn_layout(A& a, PfnWs* ws, int B, int P, int C, int H, int W, int nblocks) {
  auto p0 = a.template take<float>((size_t)B * P * 2 * C);
  auto p1 = a.template take<double>((size_t)nblocks * 2 * C);
  auto p2 = a.template take<Affine>(64);
  auto p3 = a.template take<int>((size_t)B * H * W + 1);
  auto p4 = a.template take<int>(64);
  if (ws) { ws->ext = p0; ws->partials = p1; ws->affine = p2; ws->map = p3; ws->flags = p4; }
}

struct SizeArena2 {
  size_t used = 0;
  template <class T>
  T* take(size_t count) { used += align_up(count * sizeof(T)); return nullptr; }
};

static int stats_blocks(long long rows) {
  long long nb = (rows + kWarps - 1) / kWarps;
  const int sms = sm_count();
  if (nb > sms) nb = sms;
  if (nb < 1) nb = 1;
  return (int)nb;
}

// tensor-core (tcgen05) variant, pfn_tc.cu
bool pfn_tc_supported(int D, int N, int C, const void* x);
int launch_stats_tc(const float* d_x, int B, int P, int N, const float* w, const float* bias,
                    const float* bn_w, int training, float* ext, double* partials, int nblocks,
                    const int* run_flag, cudaStream_t st);
// fp16 compensated-split variant, pfn_tc16.cu (raises *range_flag when an input leaves the fp16 range)
bool pfn_tc16_supported(int D, int N, int C, int P, const void* x);
int launch_stats_tc16(const float* d_x, int B, int P, int N, const float* w, const float* bias,
                      const float* bn_w, int training, float* ext, double* partials, int nblocks,
                      int* range_flag, cudaStream_t st);
extern int g_opt_pfn_tensor_cores;
extern int g_opt_pad_reserve_sms;
extern int g_opt_pfn_tc_debug;

static int launch_stats(const float* d_x, int B, int P, int N, int C, const float* w, const float* bias,
                        const float* bn_w, int training, PfnWs& ws, int& nblocks, cudaStream_t st) {
  if (g_opt_pfn_tensor_cores && pfn_tc_supported(kD, N, C, d_x)) {
    const long long pairs = ((long long)B * P + 1) / 2;
    nblocks = (int)(pairs < sm_count() ? pairs : sm_count());
    if (g_opt_pfn_tensor_cores == 1 && pfn_tc16_supported(kD, N, C, P, d_x)) {
      // fp16 fast path, then the TF32 kernel as a guarded fallback: it returns at once unless the
      // fast path found a value outside the fp16 range, in which case it recomputes every output
      const int rc = launch_stats_tc16(d_x, B, P, N, w, bias, bn_w, training, ws.ext, ws.partials, nblocks, ws.flags, st);
      if (rc != PP_OK) return rc;
      return launch_stats_tc(d_x, B, P, N, w, bias, bn_w, training, ws.ext, ws.partials, nblocks, ws.flags, st);
    }
    return launch_stats_tc(d_x, B, P, N, w, bias, bn_w, training, ws.ext, ws.partials, nblocks, nullptr, st);
  }
  int chunk = N < kMaxChunk ? N : kMaxChunk;
  chunk = (chunk + 3) & ~3;
  const int nchunks = (N + chunk - 1) / chunk;
  const bool use_bulk = (N % 4 == 0) && ((uintptr_t)d_x % 16 == 0);
  const size_t smem = (size_t)kWarps * kStages * kD * chunk * 4 + (size_t)kWarps * kStages * 8;
#define PP_STATS(CPL, TR)                                                                         \
  do {                                                                                            \
    PP_CUDA(cudaFuncSetAttribute(k_pfn_stats<CPL, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)smem));                                                     \
    PP_KERNEL("k_pfn_stats", st,                                                                  \
              (k_pfn_stats<CPL, TR><<<nblocks, kWarps * 32, smem, st>>>(                          \
                  d_x, B, P, N, chunk, nchunks, use_bulk, w, bias, ws.ext, ws.partials)));        \
  } while (0)
  if (C == 64) {
    if (training) PP_STATS(2, true); else PP_STATS(2, false);
  } else {
    if (training) PP_STATS(1, true); else PP_STATS(1, false);
  }
#undef PP_STATS
  return PP_OK;
}

static int pfn_common(const float* d_x, int B, int D, int P, int N, int C, const float* w,
                      const float* bias, const float* bn_w, const float* bn_b, float* rm, float* rv,
                      int64_t* nbt, int training, float momentum, float eps, PfnWs& ws, int nblocks,
                      cudaStream_t st) {
  int rc = launch_stats(d_x, B, P, N, C, w, bias, bn_w, training, ws, nblocks, st);
  if (rc != PP_OK) return rc;
  PP_KERNEL("k_bn_finalize", st,
            k_bn_finalize<<<1, 64 * kFinSegs, 0, st>>>(C, nblocks, (double)B * P * N, training, momentum, eps,
                                            SparseFinalize{1.0, nullptr, 0, nullptr, 0.0, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, nullptr}, ws.partials, bn_w, bn_b,
                                            rm, rv, (long long*)nbt, ws.affine));
  return PP_OK;
}

static bool pfn_args_ok(const void* x, int B, int D, int P, int N, int C, const void* w,
                        const void* b, const void* g, const void* be, const void* rm,
                        const void* rv) {
  return x && w && b && g && be && rm && rv && B >= 1 && D == kD && P >= 1 && N >= 1 &&
         (C == 32 || C == 64) && (long long)B * P < 0x7fffffffll;
}

static int canvas_launch(int from_ext, const float* src, const Affine* aff, const int* map, int B,
                         int P, int C, int H, int W, float* d_canvas, cudaStream_t st, int map_bias = 0) {
  const int HW = H * W;
  const bool vec_ok = (HW % 4 == 0) && ((uintptr_t)d_canvas % 16 == 0);
  // exactly one resident wave (8 CTAs of 8 warps per SM), fewer when there is less work than that
  const long long units = (long long)B * ((HW + 127) / 128) * ((C + kChanPerUnit - 1) / kChanPerUnit);
  long long gx = (units + 7) / 8;
  int per_sm = 0;
  if (from_ext == 2) PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_canvas<2>, 256, 0));
  else if (from_ext == 3) PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_canvas<3>, 256, 0));
  else PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_canvas<0>, 256, 0));
  const long long cap = (long long)sm_count() * (per_sm > 0 ? per_sm : 1);
  if (gx > cap) gx = cap;
  if (from_ext == 2) {
    PP_KERNEL("k_canvas", st, k_canvas<2><<<(int)gx, 256, 0, st>>>(src, aff, map, map_bias, B, P, C, HW, vec_ok, d_canvas));
  } else if (from_ext == 3) {
    PP_KERNEL("k_canvas", st, k_canvas<3><<<(int)gx, 256, 0, st>>>(src, aff, map, map_bias, B, P, C, HW, vec_ok, d_canvas));
  } else {
    PP_KERNEL("k_canvas", st, k_canvas<0><<<(int)gx, 256, 0, st>>>(src, aff, map, map_bias, B, P, C, HW, vec_ok, d_canvas));
  }
  return PP_OK;
}

__global__ void k_flag_status(const int* __restrict__ flag, int* __restrict__ status, int bit) {
  if (threadIdx.x == 0 && *flag != 0) atomicOr(status, bit);
}

static int build_map(const int64_t* d_inds, int B, int P, int H, int W, int* map, int32_t* d_status,
                     cudaStream_t st) {
  PP_CUDA(cudaMemsetAsync(map, 0xff, (size_t)B * H * W * sizeof(int), st));
  const long long n = (long long)B * P;
  PP_KERNEL("k_build_map", st, k_build_map<<<(int)((n + 255) / 256), 256, 0, st>>>((const long long*)d_inds, B, P, H, W, map,
                                                      d_status));
  return PP_OK;
}

// ---- sparse path (pp_input_path): K1's compact state in, canvas out, x never materialised --------
// For slot (b,p,n) the network input is x = f - mean[:,p,n] with f = the point's decorated features,
// or f = 0 for a padding slot.  A padding slot therefore holds the same value in every sweep, and
//   sum_{b,p,n} g(y)            = B * sum_{p,n} g(y_pad[p,n]) + sum_{real} (g(y_real) - g(y_pad[p,n]))
//   max_n y[b,p,n]              = max( max_{n < cnt} y_real , max_{n >= cnt} y_pad[p,n] )
// for g = relu, relu^2.  The padding pass (k_pfn_pad_tc, tensor cores) evaluates y_pad once per (p,n)
// instead of once per sweep and keeps one suffix maximum per sweep; k_pfn_real below handles the
// ~1.3 % of slots that hold a point.  Results equal the dense path's up to summation order.

// k_pfn_real: the live pillars (~1.3 % of the slots hold a point).  CPL channels per lane, conv weights in
// registers (pre-multiplied by sign(gamma): f = s*y, so only a maximum is tracked).  For a pillar with cnt points
//   slots n <  cnt : y_real and y_pad  -> extreme of y_real, statistics corrections g(y_real) - g(y_pad)
//   slots cnt <= n < E : y_pad only    -> extreme over the padding slots below the next ladder boundary E
// (E = 2, 4, 8, 16, 48 or N, see pfn_pad.cu) and the extreme over n >= E comes from the padding table.
// Work unit = group of 8 consecutive live pillars, handed out by an atomic counter (the median pillar holds 2
// points, p99 22, max 200: static assignment left a 15 % tail).  One batch of loads covers the first four slots
// of all eight pillars (lane = pillar * 4 + slot: features and per-slot means, 18 loads per lane, plus the eight
// table rows), staged in a per-warp shared-memory tile as {x_real[9], x_pad[9]} records and read back as
// warp-uniform LDS.128 broadcasts; pillars with more than four slots to evaluate (30 %) continue 32 slots at a
// time.  The metadata of the next group is in flight while the current one is evaluated.  The round-1 kernel
// took one pillar per warp with three dependent memory round trips each and ~670 instructions per pillar
// (86 us for 66 k pillars).
// ext_s[b*P+p][c] = extreme pre-activation of the whole pillar (sign-selected).
constexpr int kRealWarps = 8;
constexpr int kRealRec = 20;       // floats per staged slot: 9 real + 9 padding + 2 pad (16-byte multiples)
constexpr int kRealGroup = 8;      // pillars per work unit
constexpr int kRealLongFrom = 48;  // pillars with more slots to evaluate than this go to k_pfn_real_long
constexpr int kLongWarps = 4;      // warps per block of k_pfn_real_long
template <int CPL>
__global__ void __launch_bounds__(kRealWarps * 32, 2) k_pfn_real(CompactPillars cp, int C,
                                                  const float* __restrict__ conv_w,
                                                  const float* __restrict__ conv_b,
                                                  const float* __restrict__ bn_w,
                                                  const float* __restrict__ padtab,
                                                  float* __restrict__ ext_s,
                                                  double* __restrict__ partials2,
                                                  int* __restrict__ work_counter,
                                                  int* __restrict__ long_count, int* __restrict__ long_list) {
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  extern __shared__ __align__(16) unsigned char real_smem[];
  typedef float Tile[32][kRealRec];
  typedef float Rows[kRealGroup][64];
  Tile* s_pts = reinterpret_cast<Tile*>(real_smem);                                   // [warps] batch tile
  Tile* s_pts2 = s_pts + kRealWarps;                                                  // [warps] long-pillar tile; the fp64 sums at the end
  Rows* s_tab = reinterpret_cast<Rows*>(s_pts2 + kRealWarps);                         // [warps] the group's eight table rows
  int* s_pref = reinterpret_cast<int*>(s_tab + kRealWarps);                           // [PP_MAX_SWEEPS + 1]
  const int P = cp.P, N = cp.N, B = cp.sw.n_sweeps;
  if (threadIdx.x == 0) {
    int a = 0;
    for (int b = 0; b < B; ++b) { s_pref[b] = a; a += min(cp.num_pillars[b], P); }
    s_pref[B] = a;
  }
  float w[CPL][kD], bias[CPL], sgn[CPL];       // sign-folded: f = s*y = sum_d w*x + bias
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = CPL * lane + j;
    sgn[j] = bn_w[c] < 0.f ? -1.f : 1.f;
    bias[j] = sgn[j] * conv_b[c];
#pragma unroll
    for (int d = 0; d < kD; ++d) w[j][d] = sgn[j] * conv_w[c * kD + d];
  }
  __syncthreads();
  double accS[CPL], accQ[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { accS[j] = 0.0; accQ[j] = 0.0; }
  const unsigned PN = (unsigned)P * (unsigned)N;          // host guarantees P*N < 2^31
  const bool has_mean = cp.data_mean != nullptr;
  const int total = s_pref[B];
  const int ngroups = (total + kRealGroup - 1) / kRealGroup;
  float* tile = &s_pts[warp][0][0];

  // per-lane metadata of pillar g = lane (lanes 0..7): row, pillar, first point, clamped count, table limit / row
  struct Meta { int r, p, cnt, E, row; long long first; };
  auto next_group = [&]() -> int {
    int g = 0;
    if (lane == 0) g = atomicAdd(work_counter, 1);
    return __shfl_sync(0xffffffffu, g, 0);
  };
  auto load_meta = [&](int gid, Meta& m, int& cnt_raw, int& off_raw) {
    m.r = -1;
    cnt_raw = 0; off_raw = 0;
    // group gid = live pillars gid, gid + ngroups, gid + 2 ngroups, ...: dense pillars sit next to each other in
    // slot order (first-touch order follows the lidar rings), and a group of eight CONSECUTIVE pillars could hold
    // eight 200-point pillars (ncu r2g: 28 % of the samples were warps waiting at the end for such a group)
    const int i = gid + (int)lane * ngroups;
    if (gid < ngroups && lane < kRealGroup && i < total) {
      int b = 0;
      while (i >= s_pref[b + 1]) ++b;
      m.p = i - s_pref[b];
      m.r = b * P + m.p;
      m.first = cp.sw.off[b];
      cnt_raw = __ldg(cp.pil_cnt + m.r);
      off_raw = __ldg(cp.pil_off + m.r);
    }
  };
  auto finish_meta = [&](Meta& m, int cnt_raw, int off_raw) {
    if (m.r < 0) return;
    m.cnt = min(cnt_raw, N);
    m.first += off_raw;
    if (!has_mean) { m.row = -1; m.E = m.cnt; }
    else if (m.cnt <= 2) { m.row = 0; m.E = min(2, N); }
    else if (m.cnt <= 4) { m.row = 1; m.E = min(4, N); }
    else if (m.cnt <= 8) { m.row = 2; m.E = min(8, N); }
    else if (m.cnt <= 16) { m.row = 3; m.E = min(16, N); }
    else if (m.cnt <= 48) { m.row = 4; m.E = min(48, N); }
    else { m.row = -1; m.E = N; }
    if (m.E > kRealLongFrom) {
      // 4 % of the pillars, a third of all slots, and up to 40 us of serial evaluation in one warp (the slowest
      // warps took twice the median, scripts per-warp timeline): left to k_pfn_real_long, whose blocks deal the
      // 32-slot chunks of a pillar to their warps
      long_list[atomicAdd(long_count, 1)] = m.r;
      m.r = -1;
    }
  };
  float mx[CPL], ds[CPL], dq[CPL];
  auto eval = [&](const float* rec_base, int k_real, int k_all) {
    int k = 0;
    for (; k < k_real; ++k) {
      const float4* rec = reinterpret_cast<const float4*>(rec_base + k * kRealRec);
      const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3], r4 = rec[4];
      const float a[kD] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x};
      const float q[kD] = {r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w, r4.x, r4.y};
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        float f = bias[j], fp = bias[j];
#pragma unroll
        for (int d = 0; d < kD; ++d) { f = fmaf(w[j][d], a[d], f); fp = fmaf(w[j][d], q[d], fp); }
        mx[j] = fmaxf(mx[j], f);
        const float t = fmaf(sgn[j], f, fabsf(f)), tp = fmaf(sgn[j], fp, fabsf(fp));   // 2 relu(y), y = s f
        ds[j] += t - tp;                                   // the padding pass counted this slot as padding
        dq[j] += fmaf(t, t, -tp * tp);
      }
    }
    for (; k < k_all; ++k) {                               // padding slots below the table's first slot
      const float4* rec = reinterpret_cast<const float4*>(rec_base + k * kRealRec);
      const float4 r2 = rec[2], r3 = rec[3], r4 = rec[4];
      const float q[kD] = {r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w, r4.x, r4.y};
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        float fp = bias[j];
#pragma unroll
        for (int d = 0; d < kD; ++d) fp = fmaf(w[j][d], q[d], fp);
        mx[j] = fmaxf(mx[j], fp);
      }
    }
  };
  // slot n of the pillar (first point `first`, pillar index p) -> record at tile[slot_in_tile]
  auto fetch_stage = [&](bool on, long long first, int p, int n, int cnt, float* rec) {
    if (!on) return;
    const float* f = cp.feat_c + (size_t)first * kFeatStride;
    float fv[kD], mv[kD];
#pragma unroll
    for (int d = 0; d < kD; ++d) {
      fv[d] = n < cnt ? __ldg(f + (unsigned)n * kFeatStride + d) : 0.f;
      mv[d] = has_mean ? __ldg(cp.data_mean + (size_t)p * N + (unsigned)d * PN + (unsigned)n) : 0.f;
    }
#pragma unroll
    for (int d = 0; d < kD; ++d) {
      rec[d] = __fsub_rn(fv[d], mv[d]);                    // data/dataset.py:105
      rec[kD + d] = __fsub_rn(0.f, mv[d]);                 // what the slot holds when it is padding
    }
  };

  Meta mA{}, mB{};
  int cntA = 0, offA = 0, cntB = 0, offB = 0;
  int gidA = next_group();
  load_meta(gidA, mA, cntA, offA);
  int gidB = next_group();
  load_meta(gidB, mB, cntB, offB);
  while (gidA < ngroups) {
    finish_meta(mA, cntA, offA);
    // one batch: lane = 4 * pillar + slot
    {
      const int g = (int)lane >> 2, n = (int)lane & 3;
      const int r_g = __shfl_sync(0xffffffffu, mA.r, g);
      const int p_g = __shfl_sync(0xffffffffu, mA.p, g);
      const int cnt_g = __shfl_sync(0xffffffffu, mA.cnt, g);
      const int E_g = __shfl_sync(0xffffffffu, mA.E, g);
      const long long first_g = __shfl_sync(0xffffffffu, mA.first, g);
      // the eight table rows (one float2 per lane each), all in flight together with the slot loads
      float2 tb[kRealGroup];
#pragma unroll
      for (int k = 0; k < kRealGroup; ++k) {
        const int row_k = __shfl_sync(0xffffffffu, mA.row, k);
        const int p_k = __shfl_sync(0xffffffffu, mA.p, k);
        tb[k] = make_float2(0.f, 0.f);
        if (row_k >= 0) tb[k] = __ldg(reinterpret_cast<const float2*>(padtab + (size_t)p_k * 320 + row_k * 64) + lane);
      }
      fetch_stage(r_g >= 0 && n < E_g, first_g, p_g, n, cnt_g, tile + lane * kRealRec);
#pragma unroll
      for (int k = 0; k < kRealGroup; ++k) reinterpret_cast<float2*>(&s_tab[warp][k][0])[lane] = tb[k];
    }
    __syncwarp();
    // the metadata of the group after next is in flight during this group's evaluation (each warp holds at most
    // two groups: with ~3.5 groups per warp a deeper pipeline would hand out everything in the prologue)
    Meta mC{};
    int cntC = 0, offC = 0;
    const int gidC = next_group();
    load_meta(gidC, mC, cntC, offC);
#pragma unroll 1
    for (int g = 0; g < kRealGroup; ++g) {
      const int r_g = __shfl_sync(0xffffffffu, mA.r, g);
      if (r_g < 0) continue;
      const int p_g = __shfl_sync(0xffffffffu, mA.p, g);
      const int cnt_g = __shfl_sync(0xffffffffu, mA.cnt, g);
      const int E_g = __shfl_sync(0xffffffffu, mA.E, g);
      const int row_g = __shfl_sync(0xffffffffu, mA.row, g);
#pragma unroll
      for (int j = 0; j < CPL; ++j) { mx[j] = -INFINITY; ds[j] = 0.f; dq[j] = 0.f; }
      eval(tile + 4 * g * kRealRec, min(cnt_g, 4), min(E_g, 4));
      if (E_g > 4) {
        // the rest of a longer pillar, 32 slots at a time; the batch tile is still needed by the following pillars
        // of the group, so these records go to a second tile
        const long long first_g = __shfl_sync(0xffffffffu, mA.first, g);
        for (int n0 = 4; n0 < E_g; n0 += 32) {
          __syncwarp();
          fetch_stage(n0 + (int)lane < E_g, first_g, p_g, n0 + (int)lane, cnt_g, &s_pts2[warp][lane][0]);
          __syncwarp();
          eval(&s_pts2[warp][0][0], max(0, min(32, cnt_g - n0)), min(32, E_g - n0));
        }
      }
      // extreme over the whole pillar: its points and padding slots below E (above), the table row for n >= E
      float* e = ext_s + (size_t)r_g * C + CPL * lane;
      const float2 tab = reinterpret_cast<const float2*>(&s_tab[warp][g][0])[lane];
      const float tb2[2] = {tab.x, tab.y};
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        float m = mx[j];
        if (has_mean) { if (row_g >= 0) m = fmaxf(m, sgn[j] * tb2[j]); }
        else if (cnt_g < N) m = fmaxf(m, bias[j]);
        e[j] = sgn[j] * m;
        accS[j] += (double)ds[j];
        accQ[j] += (double)dq[j];
      }
    }
    __syncwarp();                                          // the tile is refilled by the next group
    mA = mB; cntA = cntB; offA = offB; gidA = gidB;
    mB = mC; cntB = cntC; offB = offC; gidB = gidC;
  }
  static_assert(32 * kRealRec * 4 >= 2 * 64 * 8, "per-warp tile holds the warp's fp64 sums");
  double* red = reinterpret_cast<double*>(&s_pts2[warp][0][0]);   // [2][64], this warp's own tile
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    red[CPL * lane + j] = accS[j] * 0.5;                      // t = 2 relu(y)
    red[64 + CPL * lane + j] = accQ[j] * 0.25;
  }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    const int q = threadIdx.x / C, c = threadIdx.x % C;
    double v = 0.0;
    for (int wv = 0; wv < kRealWarps; ++wv) v += reinterpret_cast<const double*>(&s_pts2[wv][0][0])[q * 64 + c];
    partials2[((size_t)blockIdx.x * 2 + q) * C + c] = v;
  }
}

// The pillars k_pfn_real left on its list (more than kRealLongFrom slots to evaluate): one pillar per block at a
// time, its 32-slot chunks dealt to the block's warps, partial extremes combined in shared memory.  Same per-slot
// arithmetic as k_pfn_real.
template <int CPL>
__global__ void __launch_bounds__(kLongWarps * 32, 4) k_pfn_real_long(CompactPillars cp, int C,
                                                  const float* __restrict__ conv_w,
                                                  const float* __restrict__ conv_b,
                                                  const float* __restrict__ bn_w,
                                                  const float* __restrict__ padtab,
                                                  float* __restrict__ ext_s,
                                                  double* __restrict__ partials3,
                                                  const int* __restrict__ long_count,
                                                  const int* __restrict__ long_list) {
  const int warp = threadIdx.x >> 5;
  const int lane = (int)lane_id();
  __shared__ __align__(16) float s_tile[kLongWarps][32][kRealRec];
  __shared__ float s_mx[kLongWarps][64];
  __shared__ double s_red[kLongWarps][2][64];
  const int P = cp.P, N = cp.N;
  const unsigned PN = (unsigned)P * (unsigned)N;
  const bool has_mean = cp.data_mean != nullptr;
  float w[CPL][kD], bias[CPL], sgn[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = CPL * lane + j;
    sgn[j] = bn_w[c] < 0.f ? -1.f : 1.f;
    bias[j] = sgn[j] * conv_b[c];
#pragma unroll
    for (int d = 0; d < kD; ++d) w[j][d] = sgn[j] * conv_w[c * kD + d];
  }
  double accS[CPL], accQ[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { accS[j] = 0.0; accQ[j] = 0.0; }
  const int n_long = *long_count;
  float* tile = &s_tile[warp][0][0];
  for (int item = (int)blockIdx.x; item < n_long; item += (int)gridDim.x) {
    const int r = long_list[item];
    const int b = r / P, p = r - b * P;
    const int cnt = min(__ldg(cp.pil_cnt + r), N);
    const long long first = cp.sw.off[b] + __ldg(cp.pil_off + r);
    int row = -1, E = cnt;
    if (has_mean) {
      if (cnt <= 16) { row = 3; E = min(16, N); }
      else if (cnt <= 48) { row = 4; E = min(48, N); }
      else { row = -1; E = N; }
    }
    float mx[CPL], ds[CPL], dq[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) { mx[j] = -INFINITY; ds[j] = 0.f; dq[j] = 0.f; }
    for (int n0 = 32 * warp; n0 < E; n0 += 32 * kLongWarps) {
      const int n = n0 + lane;
      if (n < E) {
        const float* f = cp.feat_c + (size_t)first * kFeatStride;
        float fv[kD], mv[kD];
#pragma unroll
        for (int d = 0; d < kD; ++d) {
          fv[d] = n < cnt ? __ldg(f + (unsigned)n * kFeatStride + d) : 0.f;
          mv[d] = has_mean ? __ldg(cp.data_mean + (size_t)p * N + (unsigned)d * PN + (unsigned)n) : 0.f;
        }
        float* rec = tile + lane * kRealRec;
#pragma unroll
        for (int d = 0; d < kD; ++d) {
          rec[d] = __fsub_rn(fv[d], mv[d]);                  // data/dataset.py:105
          rec[kD + d] = __fsub_rn(0.f, mv[d]);               // what the slot holds when it is padding
        }
      }
      __syncwarp();
      const int k_all = min(32, E - n0), k_real = max(0, min(32, cnt - n0));
      int k = 0;
      for (; k < k_real; ++k) {
        const float4* rec = reinterpret_cast<const float4*>(tile + k * kRealRec);
        const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3], r4 = rec[4];
        const float a[kD] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x};
        const float q[kD] = {r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w, r4.x, r4.y};
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          float fx = bias[j], fp = bias[j];
#pragma unroll
          for (int d = 0; d < kD; ++d) { fx = fmaf(w[j][d], a[d], fx); fp = fmaf(w[j][d], q[d], fp); }
          mx[j] = fmaxf(mx[j], fx);
          const float t = fmaf(sgn[j], fx, fabsf(fx)), tp = fmaf(sgn[j], fp, fabsf(fp));   // 2 relu(y), y = s f
          ds[j] += t - tp;
          dq[j] += fmaf(t, t, -tp * tp);
        }
      }
      for (; k < k_all; ++k) {
        const float4* rec = reinterpret_cast<const float4*>(tile + k * kRealRec);
        const float4 r2 = rec[2], r3 = rec[3], r4 = rec[4];
        const float q[kD] = {r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w, r4.x, r4.y};
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          float fp = bias[j];
#pragma unroll
          for (int d = 0; d < kD; ++d) fp = fmaf(w[j][d], q[d], fp);
          mx[j] = fmaxf(mx[j], fp);
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      s_mx[warp][CPL * lane + j] = mx[j];
      accS[j] += (double)ds[j];
      accQ[j] += (double)dq[j];
    }
    __syncthreads();
    if (warp == 0) {
      float2 tab = make_float2(0.f, 0.f);
      if (row >= 0) tab = __ldg(reinterpret_cast<const float2*>(padtab + (size_t)p * 320 + row * 64) + lane);
      const float tb2[2] = {tab.x, tab.y};
      float* e = ext_s + (size_t)r * C + CPL * lane;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        float m = s_mx[0][CPL * lane + j];
        for (int wv = 1; wv < kLongWarps; ++wv) m = fmaxf(m, s_mx[wv][CPL * lane + j]);
        if (has_mean) { if (row >= 0) m = fmaxf(m, sgn[j] * tb2[j]); }
        else if (cnt < N) m = fmaxf(m, bias[j]);
        e[j] = sgn[j] * m;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    s_red[warp][0][CPL * lane + j] = accS[j] * 0.5;           // t = 2 relu(y)
    s_red[warp][1][CPL * lane + j] = accQ[j] * 0.25;
  }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    const int q = threadIdx.x / C, c = threadIdx.x % C;
    double v = 0.0;
    for (int wv = 0; wv < kLongWarps; ++wv) v += s_red[wv][q][c];
    partials3[((size_t)blockIdx.x * 2 + q) * C + c] = v;
  }
}

struct SparseWs {
  float* ext_s;        // [B*P, C]
  float* padtab;       // [P, 5, C]   padding table (pfn_pad.cu)
  double* partials;    // [nblocks, 2, C]   padding pass
  double* partials2;   // [nblocks2 + nblocks3, 2, C]  k_pfn_real, then k_pfn_real_long
  int* long_list;      // [B*P] rows of the pillars left to k_pfn_real_long
  Affine* affine;
  int* map;
  int* flags;
  void* prep;          // operand prepared on the fly when the caller did not pass one
};

static int real_blocks() { return sm_count() * 2; }
static int long_blocks() { return sm_count() * 2; }
constexpr size_t kRealSmem = (size_t)kRealWarps * (2 * 32 * kRealRec * 4 + kRealGroup * 64 * 4) + (PP_MAX_SWEEPS + 1) * 4 + 12;

bool pfn_pad_supported(int N, int C, int P);
size_t mean_prepared_bytes(int P, int N);
const double* mean_prepared_moments(const void* prep, int P, int N);
const int* mean_prepared_flag(const void* prep, int P, int N);
int mean_prepare(const float* d_mean, int P, int N, void* d_prep, size_t bytes, cudaStream_t st);
int launch_pad_tc(const void* d_prep, int P, int N, const float* w, const float* bias, const float* bn_w, int training,
                  float* padtab, double* partials, int nblocks, int* range_flag, cudaStream_t st);

template <class A>
static void sparse_layout(A& a, SparseWs* ws, int B, int P, int N, int C, int H, int W, bool own_prep) {
  auto p0 = a.template take<float>((size_t)B * P * C);
  auto p1 = a.template take<double>((size_t)sm_count() * 2 * C);
  auto p2 = a.template take<double>((size_t)(real_blocks() + long_blocks()) * 2 * C);
  auto p3 = a.template take<Affine>(64);
  auto p4 = a.template take<int>((size_t)B * H * W + 1);
  auto p5 = a.template take<int>(64);
  auto p6 = a.template take<float>((size_t)P * 5 * C);
  auto p7 = a.template take<char>(own_prep ? mean_prepared_bytes(P, N) : 0);
  auto p8 = a.template take<int>((size_t)B * P);
  if (ws) { ws->ext_s = p0; ws->partials = p1; ws->partials2 = p2; ws->affine = p3; ws->map = p4; ws->flags = p5; ws->padtab = p6; ws->prep = p7; ws->long_list = p8; }
}

size_t pfn_sparse_workspace_bytes(int B, int P, int N, int C, int H, int W, bool own_prep) {
  SizeArena2 a;
  sparse_layout(a, (SparseWs*)nullptr, B, P, N, C, H, W, own_prep);
  return a.used + kAlign;
}

bool pfn_sparse_supported(int B, int P, int N, int C, const void* data_mean) {
  if (!g_opt_pfn_tensor_cores) return false;
  if (B < 1 || B > PP_MAX_SWEEPS || C != 64 || N > 255) return false;
  return data_mean == nullptr || pfn_pad_supported(N, C, P);
}

int pfn_sparse_scatter(const CompactPillars& cp, const int64_t* d_inds, int C, const PfnParams& prm, int H, int W,
                       float* d_canvas, int32_t* d_status, void* d_ws, size_t ws_bytes, cudaStream_t st) {
  const int B = cp.sw.n_sweeps, P = cp.P, N = cp.N;
  if (!pfn_sparse_supported(B, P, N, C, cp.data_mean)) return PP_ERR_UNSUPPORTED;
  const bool own_prep = cp.data_mean != nullptr && cp.mean_prepared == nullptr;
  Arena arena(d_ws, ws_bytes);
  SparseWs ws{};
  sparse_layout(arena, &ws, B, P, N, C, H, W, own_prep);
  if (!arena.ok) return PP_ERR_WORKSPACE;
  int rc = PP_OK;
  if (cp.cell_map == nullptr) {                 // a caller without K1's map (indices from elsewhere)
    rc = build_map(d_inds, B, P, H, W, ws.map, d_status, st);
    if (rc != PP_OK) return rc;
  }
  int nblocks = 0;
  const void* prep = cp.mean_prepared;
  PP_CUDA(cudaMemsetAsync(ws.flags, 0, 3 * sizeof(int), st));     // [0] range flag of the padding pass, [1] k_pfn_real's work counter, [2] long-pillar count
  if (cp.data_mean != nullptr) {
    if (own_prep) {
      // a caller without a prepared operand pays a streaming pass over data_mean per call (pp_mean_prepare once
      // per data_mean avoids it)
      rc = mean_prepare(cp.data_mean, P, N, ws.prep, mean_prepared_bytes(P, N), st);
      if (rc != PP_OK) return rc;
      prep = ws.prep;
    }
    const long long pairs = P / 2;
    const int sms = sm_count() - g_opt_pad_reserve_sms > 8 ? sm_count() - g_opt_pad_reserve_sms : 8;
    nblocks = (int)(pairs < sms ? pairs : sms);
    // the padding pass does not depend on the sweeps: once per call, whatever the batch size
    rc = launch_pad_tc(prep, P, N, prm.conv_w, prm.conv_b, prm.bn_w, prm.training ? 1 : 0, ws.padtab, ws.partials, nblocks,
                       ws.flags, st);
    if (rc != PP_OK) return rc;
  }
  const int nb2 = real_blocks();
  PP_CUDA(cudaFuncSetAttribute(k_pfn_real<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRealSmem));
  PP_KERNEL("k_pfn_real", st,
            k_pfn_real<2><<<nb2, kRealWarps * 32, kRealSmem, st>>>(cp, C, prm.conv_w, prm.conv_b, prm.bn_w, ws.padtab, ws.ext_s,
                                                           ws.partials2, ws.flags + 1, ws.flags + 2, ws.long_list));
  const int nb3 = long_blocks();
  PP_KERNEL("k_pfn_real_long", st,
            k_pfn_real_long<2><<<nb3, kLongWarps * 32, 0, st>>>(cp, C, prm.conv_w, prm.conv_b, prm.bn_w, ws.padtab, ws.ext_s,
                                                                ws.partials2 + (size_t)nb2 * 2 * C, ws.flags + 2, ws.long_list));
  // a mean or weight outside the fp16 range is reported through the status word (by the finalize kernel, which
  // runs anyway)
  SparseFinalize sf{};
  sf.mult = 0.5 * (double)B;
  sf.partials2 = prm.training ? ws.partials2 : nullptr;
  sf.nparts2 = nb2 + nb3;
  sf.conv_b = cp.data_mean == nullptr ? prm.conv_b : nullptr;
  sf.pad_count = (double)B * P * N;
  sf.range_flag = cp.data_mean != nullptr ? ws.flags : nullptr;
  sf.status = d_status;
  if (cp.data_mean != nullptr) {
    sf.mom = mean_prepared_moments(prep, P, N);
    sf.mom_w = prm.conv_w;
    sf.mom_b = prm.conv_b;
    sf.mom_mult = 0.5 * (double)B;
    sf.range_flag2 = mean_prepared_flag(prep, P, N);
  }
  PP_KERNEL("k_bn_finalize", st,
            k_bn_finalize<<<1, 64 * kFinSegs, 0, st>>>(C, nblocks, (double)B * P * N, prm.training, prm.momentum, prm.eps, sf,
                                            ws.partials, prm.bn_w, prm.bn_b, prm.running_mean, prm.running_var,
                                            (long long*)prm.num_batches_tracked, ws.affine));
  return canvas_launch(3, ws.ext_s, ws.affine, cp.cell_map != nullptr ? cp.cell_map : ws.map, B, P, C, H, W, d_canvas, st,
                       cp.cell_map != nullptr ? 1 : 0);
}

}  // namespace pp

extern "C" {

size_t pp_pfn_workspace_bytes(int32_t B, int32_t P, int32_t C, int32_t canvas_h, int32_t canvas_w) {
  if (B < 1 || P < 1 || C < 1 || canvas_h < 0 || canvas_w < 0) return 0;
  pp::SizeArena2 a;
  pp::pfn_layout(a, (pp::PfnWs*)nullptr, B, P, C, canvas_h, canvas_w, pp::sm_count());
  return a.used + pp::kAlign;
}

int pp_pfn_forward(const float* d_x, int32_t B, int32_t D, int32_t P, int32_t N, int32_t C,
                   const float* d_conv_w, const float* d_conv_b, const float* d_bn_w,
                   const float* d_bn_b, float* d_running_mean, float* d_running_var,
                   int64_t* d_num_batches_tracked, int32_t training, float momentum, float eps,
                   float* d_out, void* d_workspace, size_t workspace_bytes, pp_stream_t stream) {
  using namespace pp;
  cudaStream_t st = (cudaStream_t)stream;
  if (!pfn_args_ok(d_x, B, D, P, N, C, d_conv_w, d_conv_b, d_bn_w, d_bn_b, d_running_mean,
                   d_running_var) || d_out == nullptr)
    return D != kD || !(C == 32 || C == 64) ? PP_ERR_UNSUPPORTED : PP_ERR_INVALID_ARG;
  Arena arena(d_workspace, workspace_bytes);
  PfnWs ws{};
  const int nblocks = stats_blocks((long long)B * P);
  pfn_layout(arena, &ws, B, P, C, 0, 0, sm_count());
  if (!arena.ok) return PP_ERR_WORKSPACE;
  int rc = pfn_common(d_x, B, D, P, N, C, d_conv_w, d_conv_b, d_bn_w, d_bn_b, d_running_mean,
                      d_running_var, d_num_batches_tracked, training, momentum, eps, ws, nblocks, st);
  if (rc != PP_OK) return rc;
  dim3 grid((P + 31) / 32, B);
  PP_KERNEL("k_pfn_out", st, k_pfn_out<<<grid, 256, 0, st>>>(ws.ext, ws.affine, P, C, d_out));
  return PP_OK;
}

int pp_scatter(const float* d_feat, const int64_t* d_inds, int32_t B, int32_t C, int32_t P,
               int32_t canvas_h, int32_t canvas_w, float* d_canvas, int32_t* d_status,
               void* d_workspace, size_t workspace_bytes, pp_stream_t stream) {
  using namespace pp;
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_feat || !d_inds || !d_canvas || !d_status || B < 1 || P < 1 || C < 1 || canvas_h < 1 ||
      canvas_w < 1 || (long long)canvas_h * canvas_w > 0x3fffffffll)
    return PP_ERR_INVALID_ARG;
  if (C > 64) return PP_ERR_UNSUPPORTED;
  Arena arena(d_workspace, workspace_bytes);
  int* map = arena.take<int>((size_t)B * canvas_h * canvas_w + 1);
  if (!arena.ok) return PP_ERR_WORKSPACE;
  int rc = build_map(d_inds, B, P, canvas_h, canvas_w, map, d_status, st);
  if (rc != PP_OK) return rc;
  return canvas_launch(0, d_feat, nullptr, map, B, P, C, canvas_h, canvas_w, d_canvas, st);
}

int pp_pfn_scatter(const float* d_x, const int64_t* d_inds, int32_t B, int32_t D, int32_t P,
                   int32_t N, int32_t C, const float* d_conv_w, const float* d_conv_b,
                   const float* d_bn_w, const float* d_bn_b, float* d_running_mean,
                   float* d_running_var, int64_t* d_num_batches_tracked, int32_t training,
                   float momentum, float eps, int32_t canvas_h, int32_t canvas_w, float* d_canvas,
                   float* d_out, int32_t* d_status, void* d_workspace, size_t workspace_bytes,
                   pp_stream_t stream) {
  using namespace pp;
  cudaStream_t st = (cudaStream_t)stream;
  if (!pfn_args_ok(d_x, B, D, P, N, C, d_conv_w, d_conv_b, d_bn_w, d_bn_b, d_running_mean,
                   d_running_var))
    return D != kD || !(C == 32 || C == 64) ? PP_ERR_UNSUPPORTED : PP_ERR_INVALID_ARG;
  if (!d_inds || !d_canvas || !d_status || canvas_h < 1 || canvas_w < 1 ||
      (long long)canvas_h * canvas_w > 0x3fffffffll)
    return PP_ERR_INVALID_ARG;
  Arena arena(d_workspace, workspace_bytes);
  PfnWs ws{};
  const int nblocks = stats_blocks((long long)B * P);
  pfn_layout(arena, &ws, B, P, C, canvas_h, canvas_w, sm_count());
  if (!arena.ok) return PP_ERR_WORKSPACE;
  int rc = build_map(d_inds, B, P, canvas_h, canvas_w, ws.map, d_status, st);
  if (rc != PP_OK) return rc;
  rc = pfn_common(d_x, B, D, P, N, C, d_conv_w, d_conv_b, d_bn_w, d_bn_b, d_running_mean,
                  d_running_var, d_num_batches_tracked, training, momentum, eps, ws, nblocks, st);
  if (rc != PP_OK) return rc;
  rc = canvas_launch(2, ws.ext, ws.affine, ws.map, B, P, C, canvas_h, canvas_w, d_canvas, st);
  if (rc != PP_OK) return rc;
  if (d_out != nullptr) {
    dim3 grid((P + 31) / 32, B);
    PP_KERNEL("k_pfn_out", st, k_pfn_out<<<grid, 256, 0, st>>>(ws.ext, ws.affine, P, C, d_out));
  }
  return PP_OK;
}

}  // extern "C"
