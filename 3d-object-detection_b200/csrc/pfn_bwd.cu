// pfn_bwd.cu -- backward of PPFeatureNet (and of PPScatter in front of it): the gradients torch autograd
// produces through model/model.py:36-39 (conv1 1x1 -> relu -> bn1 -> max over N) for conv1.weight,
// conv1.bias, bn1.weight, bn1.bias -- SURVEY 8(f) N1, the step train.py:147 needs.
//
// With z = W x + b, r = relu(z), batch statistics mu, var over M = B*P*N elements, s = sqrt(var + eps),
// G[b,c,p] the incoming gradient routed to the arg-max element e* of every (b,c,p):
//   d beta  = sum G                 d gamma = (sum G r* - mu sum G) / s
//   d W[c,d] = gamma/s * ( T[c,d] - (d beta / M) S1[c,d] - (d gamma / (M s)) (S2[c,d] - mu S1[c,d]) )
//   T = sum G 1[z*>0] x_d(e*),   S1 = sum_e 1[z>0] x_d,   S2 = sum_e relu(z) x_d          (x_9 == 1: the bias)
// The two dense moment matrices come from training-mode BatchNorm (mean and variance depend on W); in eval
// mode only T is needed.  mu = S2[c,9]/M and var = sum r^2/M - mu^2 fall out of the same pass, so the
// backward needs nothing saved by the forward except its inputs.
//
// k_pfn_bwd: one warp per (pillar, channel half), lane = channel.  The pillar's x rows are staged in shared
// memory with cp.async (double buffered over groups of 4 pillars), every lane walks the N slots reading x
// as warp-uniform float4 broadcasts: 9 FMA for z, 21 for the moments, 4 for the running arg-max per slot and
// channel.  Per-pillar fp32 sums go through a shared staging tile into fp64 accumulators that are spread
// over the block's threads ((row, channel) pairs), so the inner loop keeps 125 registers and two blocks fit
// an SM; one fp64 partial tile per block is combined in a fixed order by k_pfn_bwd_finalize (deterministic).
#include "internal.cuh"

namespace pp {

constexpr int kBwdPil = 4;                 // pillars per block iteration (8 warps = 4 pillars x 2 channel halves)
constexpr int kBwdAcc = 33;                // S1[10], S2[10], sum r^2, T[10], sum G, sum G r*
constexpr int kBwdD = 9, kBwdC = 64;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

struct BwdGrad {             // where G[b,c,p] lives: the [B,C,P] gradient, or the canvas gradient through inds
  const float* g;
  const long long* inds;     // null: dense [B,C,P]
  int H, W;
};

constexpr int kBwdRows = 34;               // staged per pillar and channel: the 33 sums (+1 pad row)
constexpr int kBwdOwn = (kBwdAcc * kBwdC + 255) / 256;     // (k, c) pairs owned by each thread of the block

// DESC: the slots are visited in descending order, so ">=" keeps the first index among equal keys
template <bool TRAIN, bool DESC = false>
__device__ __forceinline__ void bwd_slot(const float (&wr)[kBwdD], float bc, float sgn, const float (&xs)[kBwdD], int n,
                                         float (&s1)[10], float (&s2)[10], float& q2, float& best, int& nbest) {
  float z = bc;
#pragma unroll
  for (int d = 0; d < kBwdD; ++d) z = fmaf(wr[d], xs[d], z);
  const float r = fmaxf(z, 0.f);
  if (TRAIN) {
    const float on = z > 0.f ? 1.f : 0.f;
#pragma unroll
    for (int d = 0; d < kBwdD; ++d) {
      s1[d] = fmaf(on, xs[d], s1[d]);
      s2[d] = fmaf(r, xs[d], s2[d]);
    }
    s1[9] += on;
    s2[9] += r;
    q2 = fmaf(r, r, q2);
  }
  const float key = sgn * r;
  if (DESC ? key >= best : key > best) { best = key; nbest = n; }  // the first index wins a tie
}

template <bool TRAIN>
__global__ void __launch_bounds__(256, 2) k_pfn_bwd(const float* __restrict__ x, int B, int P, int N, int Np,
                                                   const float* __restrict__ w, const float* __restrict__ bias,
                                                   const float* __restrict__ bn_w, BwdGrad gr, int vec16,
                                                   double* __restrict__ partials) {
  extern __shared__ __align__(16) float s_tile[];            // [2][kBwdPil][9][Np] | staging [kBwdPil][kBwdRows][64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pl = warp >> 1, c = (warp & 1) * 32 + lane;
  const long long BP = (long long)B * P;
  const long long ngroups = (BP + kBwdPil - 1) / kBwdPil;
  const size_t stage_floats = (size_t)kBwdPil * kBwdD * Np;
  float* s_sum = s_tile + 2 * stage_floats;

  float wr[kBwdD];
#pragma unroll
  for (int d = 0; d < kBwdD; ++d) wr[d] = __ldg(w + c * kBwdD + d);
  const float bc = __ldg(bias + c);
  const float gam = __ldg(bn_w + c);
  const float sgn = gam > 0.f ? 1.f : (gam < 0.f ? -1.f : 0.f);

  // fp64 accumulators are spread over the block: thread t owns the (k, c) pairs t, t+256, ...
  double acc[kBwdOwn];
#pragma unroll
  for (int k = 0; k < kBwdOwn; ++k) acc[k] = 0.0;

  auto issue = [&](long long grp, int stage) {
    float* dst0 = s_tile + (size_t)stage * stage_floats;
    if (vec16) {
      const int chunks = N >> 2;
      for (int i = tid; i < kBwdPil * kBwdD * chunks; i += 256) {
        const int row = i / chunks, ck = i - row * chunks;
        const int q = row / kBwdD, d = row - q * kBwdD;
        const long long task = grp * kBwdPil + q;
        if (task < BP) {
          const int b = (int)((unsigned)task / (unsigned)P), p = (int)task - b * P;
          cp_async16(dst0 + ((size_t)q * kBwdD + d) * Np + ck * 4, x + ((size_t)(b * kBwdD + d) * P + p) * N + ck * 4);
        }
      }
    } else {
      for (int i = tid; i < kBwdPil * kBwdD * N; i += 256) {
        const int row = i / N, n = i - row * N;
        const int q = row / kBwdD, d = row - q * kBwdD;
        const long long task = grp * kBwdPil + q;
        if (task < BP) {
          const int b = (int)((unsigned)task / (unsigned)P), p = (int)task - b * P;
          cp_async4(dst0 + ((size_t)q * kBwdD + d) * Np + n, x + ((size_t)(b * kBwdD + d) * P + p) * N + n);
        }
      }
    }
    cp_async_commit();
  };

  int stage = 0;
  long long grp = blockIdx.x;
  if (grp < ngroups) issue(grp, 0);
  for (; grp < ngroups; grp += gridDim.x) {
    const long long nxt = grp + gridDim.x;
    if (nxt < ngroups) {
      issue(nxt, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();                                             // tile landed; previous staging rows consumed
    const long long task = grp * kBwdPil + pl;
    float* srow = s_sum + (size_t)pl * kBwdRows * kBwdC + c;
    if (task < BP) {
      const float* t = s_tile + (size_t)stage * stage_floats + (size_t)pl * kBwdD * Np;
      float s1[10], s2[10], q2 = 0.f;
#pragma unroll
      for (int d = 0; d < 10; ++d) { s1[d] = 0.f; s2[d] = 0.f; }
      float best = -INFINITY;
      int nbest = 0;
      const int N4 = N & ~3;
      for (int n0 = 0; n0 < N4; n0 += 4) {
        float4 xv[kBwdD];
#pragma unroll
        for (int d = 0; d < kBwdD; ++d) xv[d] = *reinterpret_cast<const float4*>(t + d * Np + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float xs[kBwdD];
#pragma unroll
          for (int d = 0; d < kBwdD; ++d) xs[d] = j == 0 ? xv[d].x : (j == 1 ? xv[d].y : (j == 2 ? xv[d].z : xv[d].w));
          bwd_slot<TRAIN>(wr, bc, sgn, xs, n0 + j, s1, s2, q2, best, nbest);
        }
      }
      for (int n = N4; n < N; ++n) {
        float xs[kBwdD];
#pragma unroll
        for (int d = 0; d < kBwdD; ++d) xs[d] = t[d * Np + n];
        bwd_slot<TRAIN>(wr, bc, sgn, xs, n, s1, s2, q2, best, nbest);
      }
      // the arg-max element takes the incoming gradient
      const long long b = task / P, p = task - b * P;
      float G;
      if (gr.inds != nullptr) {
        const long long* row = gr.inds + task * 3;
        const long long fl = row[0], xi = row[1], yi = row[2];
        const bool ok = fl != 0 && xi >= 0 && xi < gr.W && yi >= 0 && yi < gr.H;
        G = ok ? __ldg(gr.g + ((size_t)(b * kBwdC + c) * gr.H + (size_t)yi) * gr.W + (size_t)xi) : 0.f;
      } else {
        G = __ldg(gr.g + (size_t)(b * kBwdC + c) * P + p);
      }
      float zs = bc;
#pragma unroll
      for (int d = 0; d < kBwdD; ++d) zs = fmaf(wr[d], t[d * Np + nbest], zs);
      const float Gon = zs > 0.f ? G : 0.f;
#pragma unroll
      for (int d = 0; d < 10; ++d) {
        srow[d * kBwdC] = s1[d];
        srow[(10 + d) * kBwdC] = s2[d];
      }
      srow[20 * kBwdC] = q2;
#pragma unroll
      for (int d = 0; d < kBwdD; ++d) srow[(21 + d) * kBwdC] = Gon * t[d * Np + nbest];
      srow[30 * kBwdC] = Gon;
      srow[31 * kBwdC] = G;
      srow[32 * kBwdC] = G * fmaxf(zs, 0.f);
    } else {
#pragma unroll
      for (int k = 0; k < kBwdAcc; ++k) srow[k * kBwdC] = 0.f;
    }
    __syncthreads();                                             // staging rows complete; tile stage free
#pragma unroll
    for (int k = 0; k < kBwdOwn; ++k) {
      const int idx = tid + k * 256;                             // = row * 64 + channel
      if (idx < kBwdAcc * kBwdC) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < kBwdPil; ++q) v += (double)s_sum[(size_t)q * kBwdRows * kBwdC + idx];
        acc[k] += v;
      }
    }
    stage ^= 1;
  }
  double* dst = partials + (size_t)blockIdx.x * kBwdAcc * kBwdC;
#pragma unroll
  for (int k = 0; k < kBwdOwn; ++k) {
    const int idx = tid + k * 256;
    if (idx < kBwdAcc * kBwdC) dst[idx] = acc[k];
  }
}

// ---- sparse formulation (pp_input_path_backward): the dense x is never built -------------------------------
// Slot n of live pillar (b,p) holds feat[n] - mean[p,n] for n < cnt[b,p] and the padding value 0 - mean[p,n]
// beyond; every slot of a pillar that is not live in sweep b holds the padding value.
//   pass A (k_pfn_bwd_pad, once over [P,N]): moment sums of the padding values (x B in the finalize; weights
//     negated: fmaf(w, -m, z) == fmaf(-w, m, z) exactly, rows d < 9 change sign in the finalize), and, scanning
//     the slots in descending order, the arg-max over the padding suffix n >= cnt[b,p] of every sweep in which
//     the pillar is live: ext[b,p,c] = (key, index), snapshot when the scan passes n == cnt[b,p].
//   pass B (k_pfn_bwd_live, live pillars only): the real slots (real - padding corrections of the moments,
//     best real slot), the winner against the padding suffix, and the routing of the canvas gradient.
constexpr int kLiveRec = 12;               // floats per staged point record (9 features, 16-byte multiple)
constexpr int kLiveStage = 32;            // points per pillar staged in shared memory; longer pillars read the rest from L2

template <bool TRAIN>
__global__ void __launch_bounds__(256, 2) k_pfn_bwd_pad(CompactPillars cp, int Np, const float* __restrict__ w,
                                                       const float* __restrict__ bias, const float* __restrict__ bn_w,
                                                       float2* __restrict__ ext, double* __restrict__ partials) {
  extern __shared__ __align__(16) float s_tile[];            // [2][kBwdPil][9][Np] | staging [kBwdPil][kBwdRows][64]
  __shared__ int s_cnt[kBwdPil][PP_MAX_SWEEPS];              // min(count, N) of the pillar in sweep b, -1 = not live there
  __shared__ unsigned s_snap[kBwdPil][8];                    // bit n: some sweep has count == n
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pl = warp >> 1, c = (warp & 1) * 32 + lane;
  const int B = cp.sw.n_sweeps, P = cp.P, N = cp.N;
  const int ngroups = (P + kBwdPil - 1) / kBwdPil;
  const size_t stage_floats = (size_t)kBwdPil * kBwdD * Np;
  float* s_sum = s_tile + 2 * stage_floats;
  float wr[kBwdD];
#pragma unroll
  for (int d = 0; d < kBwdD; ++d) wr[d] = -__ldg(w + c * kBwdD + d);
  const float bc = __ldg(bias + c);
  const float gam = __ldg(bn_w + c);
  const float sgn = gam > 0.f ? 1.f : (gam < 0.f ? -1.f : 0.f);
  double acc[kBwdOwn];
#pragma unroll
  for (int k = 0; k < kBwdOwn; ++k) acc[k] = 0.0;

  auto issue = [&](int grp, int stage) {
    float* dst0 = s_tile + (size_t)stage * stage_floats;
    const int chunks = N >> 2;
    for (int i = tid; i < kBwdPil * kBwdD * chunks; i += 256) {
      const int row = i / chunks, ck = i - row * chunks;
      const int q = row / kBwdD, d = row - q * kBwdD;
      const int p = grp * kBwdPil + q;
      if (p < P) cp_async16(dst0 + ((size_t)q * kBwdD + d) * Np + ck * 4, cp.data_mean + ((size_t)d * P + p) * N + ck * 4);
    }
    cp_async_commit();
  };

  // thread i < 4*B owns (pillar q = i / B, sweep b = i % B) of a group: its count is fetched one group ahead
  auto fetch_cnt = [&](int g) -> int {
    if (tid >= kBwdPil * B) return -1;
    const int q = tid / B, b = tid - q * B, p = g * kBwdPil + q;
    if (g >= ngroups || p >= P || p >= cp.num_pillars[b]) return -1;
    return min(cp.pil_cnt[(size_t)b * P + p], N);
  };
  int stage = 0;
  int grp = blockIdx.x;
  if (grp < ngroups) issue(grp, 0);
  int cnt_now = fetch_cnt(grp);
  for (; grp < ngroups; grp += gridDim.x) {
    const int nxt = grp + gridDim.x;
    if (nxt < ngroups) {
      issue(nxt, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    const int cnt_next = fetch_cnt(nxt);                        // in flight during this group's arithmetic
    if (tid < kBwdPil * 8) s_snap[tid >> 3][tid & 7] = 0u;
    __syncthreads();                                             // tile landed; previous staging rows consumed
    if (tid < kBwdPil * B) {
      const int q = tid / B, b = tid - q * B;
      if (cnt_now >= 0 && cnt_now < N) atomicOr(&s_snap[q][cnt_now >> 5], 1u << (cnt_now & 31));
      s_cnt[q][b] = cnt_now;
    }
    cnt_now = cnt_next;
    __syncthreads();
    const int p = grp * kBwdPil + pl;
    float* srow = s_sum + (size_t)pl * kBwdRows * kBwdC + c;
    if (p < P) {
      const float* t = s_tile + (size_t)stage * stage_floats + (size_t)pl * kBwdD * Np;
      float s1[10], s2[10], q2 = 0.f;
#pragma unroll
      for (int d = 0; d < 10; ++d) { s1[d] = 0.f; s2[d] = 0.f; }
      float best = -INFINITY;
      int nbest = 0;
      for (int b = 0; b < B; ++b)                                // full pillars: no padding slot competes
        if (s_cnt[pl][b] >= N) ext[((size_t)b * P + p) * kBwdC + c] = make_float2(-INFINITY, 0.f);
      for (int n0 = N - 4; n0 >= 0; n0 -= 4) {
        float4 xv[kBwdD];
#pragma unroll
        for (int d = 0; d < kBwdD; ++d) xv[d] = *reinterpret_cast<const float4*>(t + d * Np + n0);
        const unsigned snap = (s_snap[pl][n0 >> 5] >> (n0 & 31)) & 0xfu;      // n0 % 4 == 0: the four bits share a word
        if (snap == 0u) {                                        // the common group: no sweep's count falls in it
#pragma unroll
          for (int j = 3; j >= 0; --j) {
            float xs[kBwdD];
#pragma unroll
            for (int d = 0; d < kBwdD; ++d) xs[d] = j == 0 ? xv[d].x : (j == 1 ? xv[d].y : (j == 2 ? xv[d].z : xv[d].w));
            bwd_slot<TRAIN, true>(wr, bc, sgn, xs, n0 + j, s1, s2, q2, best, nbest);
          }
        } else {
#pragma unroll
          for (int j = 3; j >= 0; --j) {
            float xs[kBwdD];
#pragma unroll
            for (int d = 0; d < kBwdD; ++d) xs[d] = j == 0 ? xv[d].x : (j == 1 ? xv[d].y : (j == 2 ? xv[d].z : xv[d].w));
            bwd_slot<TRAIN, true>(wr, bc, sgn, xs, n0 + j, s1, s2, q2, best, nbest);
            if ((snap >> j) & 1u) {                              // warp-uniform, a handful of times per pillar
              for (int b = 0; b < B; ++b)
                if (s_cnt[pl][b] == n0 + j)
                  ext[((size_t)b * P + p) * kBwdC + c] = make_float2(best, __int_as_float(nbest));
            }
          }
        }
      }
#pragma unroll
      for (int d = 0; d < 10; ++d) {
        srow[d * kBwdC] = s1[d];
        srow[(10 + d) * kBwdC] = s2[d];
      }
      srow[20 * kBwdC] = q2;
#pragma unroll
      for (int k = 21; k < kBwdAcc; ++k) srow[k * kBwdC] = 0.f;
    } else {
#pragma unroll
      for (int k = 0; k < kBwdAcc; ++k) srow[k * kBwdC] = 0.f;
    }
    __syncthreads();
    if (TRAIN) {
#pragma unroll
      for (int k = 0; k < kBwdOwn; ++k) {
        const int idx = tid + k * 256;
        if (idx < kBwdAcc * kBwdC) {
          double v = 0.0;
#pragma unroll
          for (int q = 0; q < kBwdPil; ++q) v += (double)s_sum[(size_t)q * kBwdRows * kBwdC + idx];
          acc[k] += v;
        }
      }
    }
    stage ^= 1;
  }
  double* dst = partials + (size_t)blockIdx.x * kBwdAcc * kBwdC;
#pragma unroll
  for (int k = 0; k < kBwdOwn; ++k) {
    const int idx = tid + k * 256;
    if (idx < kBwdAcc * kBwdC) dst[idx] = acc[k];
  }
}

template <bool TRAIN>
__global__ void __launch_bounds__(256, 2) k_pfn_bwd_live(CompactPillars cp, int Np, const float* __restrict__ w,
                                                        const float* __restrict__ bias, const float* __restrict__ bn_w,
                                                        BwdGrad gr, const float2* __restrict__ ext,
                                                        double* __restrict__ partials) {
  extern __shared__ __align__(16) float s_tile[];            // 2 x { mean [kBwdPil][9][Np] | feat [kBwdPil][32][12] } | staging
  __shared__ int s_first[PP_MAX_SWEEPS + 1];                 // live pillars before sweep b
  __shared__ int s_cnt2[3][kBwdPil], s_b2[3][kBwdPil], s_p2[3][kBwdPil];   // three slots: written one pillar ahead,
                                                                          // read without a barrier after the arithmetic
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pl = warp >> 1, c = (warp & 1) * 32 + lane;
  const int B = cp.sw.n_sweeps, P = cp.P, N = cp.N;
  const size_t stage_floats = (size_t)kBwdPil * kBwdD * Np + (size_t)kBwdPil * kLiveStage * kLiveRec;
  float* s_sum = s_tile + 2 * stage_floats;
  if (tid == 0) {
    int a = 0;
    for (int b = 0; b < B; ++b) { s_first[b] = a; a += min(max(cp.num_pillars[b], 0), P); }
    s_first[B] = a;
  }
  __syncthreads();
  const int n_live = s_first[B];
  // The four warp pairs of a block run independently (pair-level named barriers): pair `pl` of block k takes the
  // live pillars pg, pg + npairs, ...  A pair that draws a 200-point pillar no longer stalls the other three.
  const int pg = blockIdx.x * kBwdPil + pl, npairs = gridDim.x * kBwdPil;
  const int t64 = tid & 63;
  auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + pl) : "memory"); };

  float wr[kBwdD];
#pragma unroll
  for (int d = 0; d < kBwdD; ++d) wr[d] = __ldg(w + c * kBwdD + d);
  const float bc = __ldg(bias + c);
  const float gam = __ldg(bn_w + c);
  const float sgn = gam > 0.f ? 1.f : (gam < 0.f ? -1.f : 0.f);
  // per-warp fp32 sums over this warp's ~60 pillars (a few hundred terms each); fp64 from the block level up
  float s1[10], s2[10], q2 = 0.f, tt[10], sG = 0.f, sGr = 0.f;
#pragma unroll
  for (int d = 0; d < 10; ++d) { s1[d] = 0.f; s2[d] = 0.f; tt[d] = 0.f; }

  // (sweep, pillar, count) of live pillar L: the pair's first thread
  auto meta = [&](int L, int ms) {
    if (t64 == 0) {
      int b = -1, p = 0, cnt = 0;
      if (L < n_live) {
        b = 0;
        while (s_first[b + 1] <= L) ++b;
        p = L - s_first[b];
        cnt = min(cp.pil_cnt[(size_t)b * P + p], N);
      }
      s_b2[ms][pl] = b; s_p2[ms][pl] = p; s_cnt2[ms][pl] = cnt;
    }
  };
  auto issue = [&](int st, int ms) {                           // the pair's 64 threads stage the pair's pillar
    float* s_mean = s_tile + (size_t)st * stage_floats + (size_t)pl * kBwdD * Np;
    float* s_feat = s_tile + (size_t)st * stage_floats + (size_t)kBwdPil * kBwdD * Np + (size_t)pl * kLiveStage * kLiveRec;
    if (s_b2[ms][pl] >= 0) {
      const int p = s_p2[ms][pl];
      for (int d = t64 >> 3; d < kBwdD; d += 8) {              // 8 threads per row of N floats
        const float* src = cp.data_mean + ((size_t)d * P + p) * N;
        for (int ck = t64 & 7; ck * 4 < N; ck += 8) cp_async16(s_mean + (size_t)d * Np + ck * 4, src + ck * 4);
      }
      const float* src = cp.feat_c + ((size_t)cp.sw.off[s_b2[ms][pl]] + cp.pil_off[(size_t)s_b2[ms][pl] * P + p]) * kFeatStride;
      for (int i = t64; i < min(s_cnt2[ms][pl], kLiveStage) * kBwdD; i += 64) {
        const int n = i / kBwdD, d = i - n * kBwdD;
        cp_async4(s_feat + (size_t)n * kLiveRec + d, src + (size_t)n * kFeatStride + d);
      }
    }
    cp_async_commit();
  };

  int stage = 0, ms = 0;
  meta(pg, 0);
  pair_sync();
  if (pg < n_live) issue(0, 0);
  // the scatter row (flag, x, y) of this warp's pillar is fetched one group ahead, the canvas gradient at the
  // top of the group: both round trips overlap the arithmetic instead of trailing it
  long long cfl = 0, cxi = 0, cyi = 0;
  if (s_b2[0][pl] >= 0) {
    const long long* row = gr.inds + ((long long)s_b2[0][pl] * P + s_p2[0][pl]) * 3;
    cfl = row[0]; cxi = row[1]; cyi = row[2];
  }
  for (int L = pg; L < n_live; L += npairs) {
    const int ms_next = ms == 2 ? 0 : ms + 1;
    meta(L + npairs, ms_next);                                 // the pair's next pillar
    cp_async_wait<0>();
    pair_sync();                                               // this pillar's tile landed, next metadata visible, the
                                                               // other tile stage is no longer read
    if (L + npairs < n_live) issue(stage ^ 1, ms_next);        // in flight during this pillar's arithmetic
    const float* s_mean = s_tile + (size_t)stage * stage_floats;
    const float* s_feat = s_mean + (size_t)kBwdPil * kBwdD * Np;
    const int* s_b = s_b2[ms];
    const int* s_p = s_p2[ms];
    const int* s_cnt = s_cnt2[ms];
    const int b = s_b[pl];
    float G = 0.f;
    if (b >= 0 && cfl != 0 && cxi >= 0 && cxi < gr.W && cyi >= 0 && cyi < gr.H)
      G = __ldg(gr.g + ((size_t)(b * kBwdC + c) * gr.H + (size_t)cyi) * gr.W + (size_t)cxi);
    long long nfl = 0, nxi = 0, nyi = 0;
    if (s_b2[ms_next][pl] >= 0) {
      const long long* row = gr.inds + ((long long)s_b2[ms_next][pl] * P + s_p2[ms_next][pl]) * 3;
      nfl = __ldg(row); nxi = __ldg(row + 1); nyi = __ldg(row + 2);
    }
    if (b >= 0) {
      const int p = s_p[pl], cnt = s_cnt[pl];
      const long long task = (long long)b * P + p;
      const float2 cand = __ldg(ext + (size_t)task * kBwdC + c);   // best padding slot at n >= cnt (pass A)
      const float* tm = s_mean + (size_t)pl * kBwdD * Np;
      const float* tf = s_feat + (size_t)pl * kLiveStage * kLiveRec;
      const float* gf = cp.feat_c + ((size_t)cp.sw.off[b] + cp.pil_off[(size_t)b * P + p]) * kFeatStride;   // the pillar's rows in HBM
      float best = -INFINITY;
      int nbest = 0;
      for (int n = 0; n < cnt; ++n) {                             // the slots that hold a point
        float m[kBwdD], xs[kBwdD];
#pragma unroll
        for (int d = 0; d < kBwdD; ++d) m[d] = tm[d * Np + n];
        if (n < kLiveStage) {
          const float4 f0 = *reinterpret_cast<const float4*>(tf + n * kLiveRec);
          const float4 f1 = *reinterpret_cast<const float4*>(tf + n * kLiveRec + 4);
          const float f8 = tf[n * kLiveRec + 8];
          xs[0] = f0.x - m[0]; xs[1] = f0.y - m[1]; xs[2] = f0.z - m[2]; xs[3] = f0.w - m[3];
          xs[4] = f1.x - m[4]; xs[5] = f1.y - m[5]; xs[6] = f1.z - m[6]; xs[7] = f1.w - m[7];
          xs[8] = f8 - m[8];
        } else {
#pragma unroll
          for (int d = 0; d < kBwdD; ++d) xs[d] = __ldg(gf + (size_t)n * kFeatStride + d) - m[d];
        }
        float z = bc;
#pragma unroll
        for (int d = 0; d < kBwdD; ++d) z = fmaf(wr[d], xs[d], z);
        const float r = fmaxf(z, 0.f);
        if (TRAIN) {
          float zp = bc;                                          // the padding value this slot replaced: 0 - mean
#pragma unroll
          for (int d = 0; d < kBwdD; ++d) zp = fmaf(-wr[d], m[d], zp);
          const float rp = fmaxf(zp, 0.f);
          const float on = z > 0.f ? 1.f : 0.f, onp = zp > 0.f ? 1.f : 0.f;
#pragma unroll
          for (int d = 0; d < kBwdD; ++d) {
            s1[d] = fmaf(on, xs[d], fmaf(onp, m[d], s1[d]));     // + on x_real - onp (0 - mean)
            s2[d] = fmaf(r, xs[d], fmaf(rp, m[d], s2[d]));
          }
          s1[9] += on - onp;
          s2[9] += r - rp;
          q2 += r * r - rp * rp;
        }
        const float key = sgn * r;
        if (key > best) { best = key; nbest = n; }               // strict: the first index wins a tie
      }
      const bool pad_wins = cand.x > best;                       // padding slots come after the real ones: ties stay real
      if (pad_wins) nbest = __float_as_int(cand.y);
      float xb[kBwdD];
#pragma unroll
      for (int d = 0; d < kBwdD; ++d) {
        const float md = tm[d * Np + nbest];
        xb[d] = pad_wins ? 0.f - md : (nbest < kLiveStage ? tf[nbest * kLiveRec + d] : __ldg(gf + (size_t)nbest * kFeatStride + d)) - md;
      }
      float zs = bc;
#pragma unroll
      for (int d = 0; d < kBwdD; ++d) zs = fmaf(wr[d], xb[d], zs);
      const float Gon = zs > 0.f ? G : 0.f;
#pragma unroll
      for (int d = 0; d < kBwdD; ++d) tt[d] = fmaf(Gon, xb[d], tt[d]);
      tt[9] += Gon;
      sG += G;
      sGr = fmaf(G, fmaxf(zs, 0.f), sGr);
    }
    stage ^= 1;
    ms = ms_next;
    cfl = nfl; cxi = nxi; cyi = nyi;
  }
  // one reduction per block: the eight warps' rows through the staging tile, pairs (pillar slot, half) as above
  cp_async_wait<0>();
  __syncthreads();
  {
    float* srow = s_sum + (size_t)pl * kBwdRows * kBwdC + c;
#pragma unroll
    for (int d = 0; d < 10; ++d) {
      srow[d * kBwdC] = s1[d];
      srow[(10 + d) * kBwdC] = s2[d];
      srow[(21 + d) * kBwdC] = tt[d];
    }
    srow[20 * kBwdC] = q2;
    srow[31 * kBwdC] = sG;
    srow[32 * kBwdC] = sGr;
  }
  __syncthreads();
  double acc[kBwdOwn];
#pragma unroll
  for (int k = 0; k < kBwdOwn; ++k) {
    acc[k] = 0.0;
    const int idx = tid + k * 256;
    if (idx < kBwdAcc * kBwdC) {
#pragma unroll
      for (int q = 0; q < kBwdPil; ++q) acc[k] += (double)s_sum[(size_t)q * kBwdRows * kBwdC + idx];
    }
  }
  double* dst = partials + (size_t)blockIdx.x * kBwdAcc * kBwdC;
#pragma unroll
  for (int k = 0; k < kBwdOwn; ++k) {
    const int idx = tid + k * 256;
    if (idx < kBwdAcc * kBwdC) dst[idx] = acc[k];
  }
}

// Sparse formulation: sums = multA * flip(partialsA) + partialsB, where set A is the padding pass over
// data_mean (rows d < 9 of S1 / S2 change sign: the slot value is 0 - mean) and B the live-pillar pass.
__global__ void __launch_bounds__(1024) k_pfn_bwd_finalize(const double* __restrict__ partials, int nblocks,
                                                           const double* __restrict__ partialsA, int nblocksA, double multA,
                                                           double M,
                                                           const float* __restrict__ bn_w,
                                                           const float* __restrict__ running_mean,
                                                           const float* __restrict__ running_var, int training, float eps,
                                                           double* __restrict__ sums_g, unsigned* __restrict__ ticket,
                                                           float* __restrict__ g_w, float* __restrict__ g_b,
                                                           float* __restrict__ g_gamma, float* __restrict__ g_beta) {
  __shared__ double part[16][kBwdC];
  __shared__ double sums[kBwdAcc][kBwdC];
  __shared__ bool last;
  const int c = threadIdx.x & 63, seg = threadIdx.x >> 6, k = blockIdx.x;
  {
    double s0 = 0.0, s1 = 0.0;
    int blk = seg;
    for (; blk + 16 < nblocks; blk += 32) {
      s0 += partials[((size_t)blk * kBwdAcc + k) * kBwdC + c];
      s1 += partials[((size_t)(blk + 16) * kBwdAcc + k) * kBwdC + c];
    }
    if (blk < nblocks) s0 += partials[((size_t)blk * kBwdAcc + k) * kBwdC + c];
    double sa = 0.0;
    if (partialsA != nullptr && k <= 20)
      for (int blk2 = seg; blk2 < nblocksA; blk2 += 16) sa += partialsA[((size_t)blk2 * kBwdAcc + k) * kBwdC + c];
    const bool flip = k < 9 || (k >= 10 && k < 19);
    part[seg][c] = s0 + s1 + (flip ? -multA : multA) * sa;
  }
  __syncthreads();
  if (seg == 0) {
    double tot = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) tot += part[q][c];
    sums_g[k * kBwdC + c] = tot;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int i = threadIdx.x; i < kBwdAcc * kBwdC; i += 1024) sums[i / kBwdC][i % kBwdC] = __ldcg(sums_g + i);
  __syncthreads();
  if (threadIdx.x == 0) *ticket = 0;                                // ready for the next call
  if (threadIdx.x < kBwdC) {
    const double gam = (double)bn_w[c];
    double mu, var;
    if (training) {
      mu = sums[19][c] / M;
      var = fmax(sums[20][c] / M - mu * mu, 0.0);
    } else {
      mu = (double)running_mean[c];
      var = (double)running_var[c];
    }
    const double s = sqrt(var + (double)eps);
    const double d_beta = sums[31][c];
    const double d_gamma = (sums[32][c] - mu * d_beta) / s;
    for (int d = 0; d < 10; ++d) {
      double v = sums[21 + d][c];
      if (training) v -= (d_beta / M) * sums[d][c] + (d_gamma / (M * s)) * (sums[10 + d][c] - mu * sums[d][c]);
      v *= gam / s;
      if (d < kBwdD) { if (g_w) g_w[c * kBwdD + d] = (float)v; }
      else if (g_b) g_b[c] = (float)v;
    }
    if (g_gamma) g_gamma[c] = (float)d_gamma;
    if (g_beta) g_beta[c] = (float)d_beta;
  }
}

// PPScatter backward: g_feat[b,c,p] = g_canvas[b,c,y,x] for rows with inds[b,p,0] != 0, else 0.
__global__ void __launch_bounds__(256) k_scatter_bwd(const float* __restrict__ g_canvas, const long long* __restrict__ inds,
                                                     int B, int C, int P, int H, int W, float* __restrict__ g_feat) {
  const size_t n = (size_t)B * C * P;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const size_t bc = i / P, p = i - bc * P, b = bc / C;
    const long long* row = inds + (b * P + p) * 3;
    const long long fl = row[0], xi = row[1], yi = row[2];
    const bool ok = fl != 0 && xi >= 0 && xi < W && yi >= 0 && yi < H;
    g_feat[i] = ok ? __ldg(g_canvas + (bc * H + (size_t)yi) * W + (size_t)xi) : 0.f;
  }
}

static int bwd_blocks() { return sm_count() * 2; }
static int bwd_live_blocks() { return sm_count() * 2; }

size_t pfn_sparse_backward_workspace_bytes(int B, int P) {
  return align_up((size_t)bwd_blocks() * kBwdAcc * kBwdC * sizeof(double)) +
         align_up((size_t)bwd_live_blocks() * kBwdAcc * kBwdC * sizeof(double)) +
         align_up((size_t)kBwdAcc * kBwdC * sizeof(double)) + align_up((size_t)B * P * kBwdC * sizeof(float2)) + 4 * kAlign;
}

// Gradients of conv1 / bn1 from K1's compact state and the canvas gradient (pp_input_path_backward).
int pfn_sparse_backward(const CompactPillars& cp, const int64_t* d_inds, int C, const float* conv_w, const float* conv_b,
                        const float* bn_w, const float* running_mean, const float* running_var, int training, float eps,
                        int H, int W, const float* d_grad_canvas, float* g_w, float* g_b, float* g_gamma, float* g_beta,
                        void* d_ws, size_t ws_bytes, cudaStream_t st) {
  const int B = cp.sw.n_sweeps, P = cp.P, N = cp.N;
  if (C != kBwdC || cp.data_mean == nullptr || N % 4 != 0 || ((uintptr_t)cp.data_mean % 16) != 0) return PP_ERR_UNSUPPORTED;
  if (!conv_w || !conv_b || !bn_w || !d_grad_canvas || !d_inds) return PP_ERR_INVALID_ARG;
  if (!training && (!running_mean || !running_var)) return PP_ERR_INVALID_ARG;
  Arena arena(d_ws, ws_bytes);
  const int nb = bwd_blocks();
  double* partialsA = arena.take<double>((size_t)nb * kBwdAcc * kBwdC);
  const int nbl = bwd_live_blocks();
  double* partialsB = arena.take<double>((size_t)nbl * kBwdAcc * kBwdC);
  double* sums_g = arena.take<double>((size_t)kBwdAcc * kBwdC);
  float2* ext = arena.take<float2>((size_t)B * P * kBwdC);
  unsigned* ticket = arena.take<unsigned>(1);
  if (!arena.ok) return PP_ERR_WORKSPACE;
  PP_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
  const int Np = N;
  BwdGrad gr{d_grad_canvas, (const long long*)d_inds, H, W};
  const size_t smem_a = ((size_t)2 * kBwdPil * kBwdD * Np + (size_t)kBwdPil * kBwdRows * kBwdC) * sizeof(float);
  const size_t smem_b = (2 * ((size_t)kBwdPil * kBwdD * Np + (size_t)kBwdPil * kLiveStage * kLiveRec) +
                         (size_t)kBwdPil * kBwdRows * kBwdC) * sizeof(float);
  if (smem_a > 110 * 1024 || smem_b > 110 * 1024) return PP_ERR_UNSUPPORTED;
  if (training) {
    PP_CUDA(cudaFuncSetAttribute(k_pfn_bwd_pad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    PP_KERNEL("k_pfn_bwd_pad", st,
              (k_pfn_bwd_pad<true><<<nb, 256, smem_a, st>>>(cp, Np, conv_w, conv_b, bn_w, ext, partialsA)));
    PP_CUDA(cudaFuncSetAttribute(k_pfn_bwd_live<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    PP_KERNEL("k_pfn_bwd_live", st,
              (k_pfn_bwd_live<true><<<nbl, 256, smem_b, st>>>(cp, Np, conv_w, conv_b, bn_w, gr, ext, partialsB)));
  } else {
    PP_CUDA(cudaFuncSetAttribute(k_pfn_bwd_pad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    PP_KERNEL("k_pfn_bwd_pad", st,
              (k_pfn_bwd_pad<false><<<nb, 256, smem_a, st>>>(cp, Np, conv_w, conv_b, bn_w, ext, partialsA)));
    PP_CUDA(cudaFuncSetAttribute(k_pfn_bwd_live<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    PP_KERNEL("k_pfn_bwd_live", st,
              (k_pfn_bwd_live<false><<<nbl, 256, smem_b, st>>>(cp, Np, conv_w, conv_b, bn_w, gr, ext, partialsB)));
  }
  PP_KERNEL("k_pfn_bwd_finalize", st,
            (k_pfn_bwd_finalize<<<kBwdAcc, 1024, 0, st>>>(partialsB, nbl, training ? partialsA : nullptr, nb, (double)B,
                                                          (double)B * P * N, bn_w, running_mean, running_var,
                                                          training ? 1 : 0, eps, sums_g, ticket, g_w, g_b, g_gamma, g_beta)));
  return PP_OK;
}

}  // namespace pp

extern "C" {

size_t pp_pfn_backward_workspace_bytes(int32_t B, int32_t P, int32_t C) {
  if (B < 1 || P < 1 || C != pp::kBwdC) return 0;
  return pp::align_up((size_t)pp::bwd_blocks() * pp::kBwdAcc * pp::kBwdC * sizeof(double)) +
         pp::align_up((size_t)pp::kBwdAcc * pp::kBwdC * sizeof(double)) + 2 * pp::kAlign;
}

int pp_pfn_backward(const float* d_x, int32_t B, int32_t D, int32_t P, int32_t N, int32_t C, const float* d_conv_w,
                    const float* d_conv_b, const float* d_bn_w, const float* d_running_mean,
                    const float* d_running_var, int32_t training, float eps, const float* d_grad_out,
                    const int64_t* d_inds, int32_t canvas_h, int32_t canvas_w, float* d_grad_conv_w,
                    float* d_grad_conv_b, float* d_grad_bn_w, float* d_grad_bn_b, void* d_workspace,
                    size_t workspace_bytes, pp_stream_t stream) {
  using namespace pp;
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_x || !d_conv_w || !d_conv_b || !d_bn_w || !d_grad_out || B < 1 || P < 1 || N < 1) return PP_ERR_INVALID_ARG;
  if (!training && (!d_running_mean || !d_running_var)) return PP_ERR_INVALID_ARG;
  if (d_inds != nullptr && (canvas_h < 1 || canvas_w < 1)) return PP_ERR_INVALID_ARG;
  if (D != kBwdD || C != kBwdC || (long long)B * P >= (1ll << 31)) return PP_ERR_UNSUPPORTED;
  const int Np = (N + 3) & ~3;
  const size_t smem = ((size_t)2 * kBwdPil * kBwdD * Np + (size_t)kBwdPil * kBwdRows * kBwdC) * sizeof(float);
  if (smem > 110 * 1024) return PP_ERR_UNSUPPORTED;
  Arena arena(d_workspace, workspace_bytes);
  const int nb = bwd_blocks();
  double* partials = arena.take<double>((size_t)nb * kBwdAcc * kBwdC);
  double* sums_g = arena.take<double>((size_t)kBwdAcc * kBwdC);
  unsigned* ticket = arena.take<unsigned>(1);
  if (!arena.ok) return PP_ERR_WORKSPACE;
  PP_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
  const int vec16 = (N % 4 == 0 && ((uintptr_t)d_x % 16) == 0) ? 1 : 0;
  BwdGrad gr{d_grad_out, (const long long*)d_inds, canvas_h, canvas_w};
  if (training) {
    PP_CUDA(cudaFuncSetAttribute(k_pfn_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PP_KERNEL("k_pfn_bwd", st,
              (k_pfn_bwd<true><<<nb, 256, smem, st>>>(d_x, B, P, N, Np, d_conv_w, d_conv_b, d_bn_w, gr, vec16, partials)));
  } else {
    PP_CUDA(cudaFuncSetAttribute(k_pfn_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PP_KERNEL("k_pfn_bwd_eval", st,
              (k_pfn_bwd<false><<<nb, 256, smem, st>>>(d_x, B, P, N, Np, d_conv_w, d_conv_b, d_bn_w, gr, vec16, partials)));
  }
  PP_KERNEL("k_pfn_bwd_finalize", st,
            (k_pfn_bwd_finalize<<<kBwdAcc, 1024, 0, st>>>(partials, nb, nullptr, 0, 0.0, (double)B * P * N, d_bn_w, d_running_mean,
                                                          d_running_var, training ? 1 : 0, eps, sums_g, ticket, d_grad_conv_w, d_grad_conv_b, d_grad_bn_w,
                                                    d_grad_bn_b)));
  return PP_OK;
}

int pp_scatter_backward(const float* d_grad_canvas, const int64_t* d_inds, int32_t B, int32_t C, int32_t P,
                        int32_t canvas_h, int32_t canvas_w, float* d_grad_feat, pp_stream_t stream) {
  using namespace pp;
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_grad_canvas || !d_inds || !d_grad_feat || B < 1 || C < 1 || P < 1 || canvas_h < 1 || canvas_w < 1)
    return PP_ERR_INVALID_ARG;
  const size_t n = (size_t)B * C * P;
  const int blocks = (int)((n + 255) / 256 < (size_t)sm_count() * 16 ? (n + 255) / 256 : (size_t)sm_count() * 16);
  PP_KERNEL("k_scatter_bwd", st,
            (k_scatter_bwd<<<blocks, 256, 0, st>>>(d_grad_canvas, (const long long*)d_inds, B, C, P, canvas_h, canvas_w,
                                                   d_grad_feat)));
  return PP_OK;
}

}  // extern "C"
