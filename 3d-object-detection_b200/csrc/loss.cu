// loss.cu -- loss front-end: the reference's PPLoss forward (model/loss.py:24-63) and the
// gradient of its total loss with respect to the two network outputs, in one pass over each tensor.
//
//   cls_out [B, Ad*K, H, W], reg_out [B, Ad*R, H, W]   NCHW network outputs (Ad = 6 anchors per cell,
//                                                      K = 9 classes, R = cfg.DATA.REG_DIMS = 8)
//   cls_t [B, A, K], reg_t [B, A, 9]                   targets of pp_assign_targets, A = H*W*Ad
//
// The reference permutes the outputs to NHWC and flattens, so logit channel d*K+k of cell (h,w) pairs
// with target element ((h*W+w)*Ad+d)*K+k: a [channels][32 cells] tile of the NCHW tensor is the
// transpose of a contiguous run of the target tensor.  k_loss_cls stages that tile in shared memory:
// every global access (logits, targets, scores, gradient) is coalesced and each element is touched
// once (the PyTorch formulation makes ~20 elementwise passes over 78 MB tensors).
//   focal weight   w = (t == 1 ? alpha : 1) * (1 - pt)^gamma, detached          (:39-44)
//   cls_loss       mean over all elements of w * bce_with_logits(x, t)            (:46)
//   in-place tanh  on channel 6 of the permuted regression output -- the reference indexes the LAST
//                  axis of the [B,H,W,Ad*R] view with 6, i.e. network channel 6 only (:50)
//   reg_loss       smooth-L1 (beta 1) over elements 0..6 of the positive anchors, mean     (:54-57)
//   ort_loss       bce_with_logits(element 7, target element 8) over the positives, mean   (:59-61)
// Sums are fp64 per-block partials combined in fixed order (deterministic).  With no positive anchor the
// reference's means over an empty tensor are NaN; so are ours.
//
// k_loss_cls_tma is the production kernel for aligned shapes (H*W % 4 == 0, 16-byte aligned tensors): a
// persistent block per SM streams [channels][128 cells] tiles through a three-stage shared-memory ring
// with TMA (one producer thread: a 2-D tensor-map box for the logit tile and a bulk copy for the target run
// complete on an mbarrier, results leave through the matching TMA stores), 16 compute warps transform each tile in place, so two tiles (110 KB) are always
// in flight per SM.  It also applies the in-place tanh.  k_loss_cls + k_loss_tanh remain for other shapes.
// k_loss_reg reads the target rows fully coalesced, appends the positives to a list and writes unscaled
// gradient rows; k_loss_finalize knows the positive count and scales the listed rows (this replaced a
// separate counting pass over the 78 MB target tensor).
#include <cuda.h>

#include "tc_common.cuh"

namespace pp {

extern int g_opt_loss_tma;
void* tensor_map_encode_fn();        // pfn_tc16.cu: cuTensorMapEncodeTiled through the runtime's entry-point query

constexpr int kLossCells = 32;      // cells (w positions) per tile
constexpr int kLossMaxCh = 96;      // Ad*K supported by the shared-memory tile

__device__ __forceinline__ double block_sum(double v, double* s_red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane_id() == 0) s_red[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < nw; ++w) r += s_red[w];
  __syncthreads();
  return r;                            // valid in thread 0
}

__global__ void __launch_bounds__(256) k_loss_cls(const float* __restrict__ cls, const float* __restrict__ cls_t,
                                                  int H, int W, int CK, float gamma, float alpha, float grad_scale,
                                                  float* __restrict__ scores, float* __restrict__ grad,
                                                  double* __restrict__ partials) {
  __shared__ float tile[kLossMaxCh][kLossCells + 1];
  __shared__ double s_red[8];
  const int b = blockIdx.z, h = blockIdx.y, w0 = blockIdx.x * kLossCells;
  const int nw = min(kLossCells, W - w0);
  const size_t plane = (size_t)H * W;
  const float* src = cls + (size_t)b * CK * plane + (size_t)h * W + w0;
  for (int idx = threadIdx.x; idx < CK * kLossCells; idx += 256) {
    const int ch = idx >> 5, wl = idx & 31;
    if (wl < nw) tile[ch][wl] = __ldg(src + (size_t)ch * plane + wl);
  }
  __syncthreads();
  const size_t base = (((size_t)b * H + h) * W + w0) * CK;       // first target element of the tile
  const bool g2 = gamma == 2.f;
  double acc = 0.0;
  for (int i = threadIdx.x; i < nw * CK; i += 256) {
    const int wl = i / CK, ch = i - wl * CK;
    const float x = tile[ch][wl];
    const float t = __ldg(cls_t + base + i);
    const float p = 1.f / (1.f + expf(-x));                    // torch.sigmoid (:39)
    const bool pos = t == 1.f;
    const float pt = pos ? p : 1.f - p;                          // (:40)
    const float om = 1.f - pt;
    const float wgt = (pos ? alpha : 1.f) * (g2 ? om * om : powf(om, gamma));   // (:41-43)
    const float bce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));        // F.binary_cross_entropy_with_logits
    acc += (double)(wgt * bce);
    if (scores != nullptr) scores[base + i] = p;
    tile[ch][wl] = grad_scale * wgt * (p - t);                   // d(b_cls * cls_loss)/dx, weight detached
  }
  __syncthreads();
  if (grad != nullptr) {
    float* dst = grad + (size_t)b * CK * plane + (size_t)h * W + w0;
    for (int idx = threadIdx.x; idx < CK * kLossCells; idx += 256) {
      const int ch = idx >> 5, wl = idx & 31;
      if (wl < nw) dst[(size_t)ch * plane + wl] = tile[ch][wl];
    }
  }
  const double s = block_sum(acc, s_red);
  if (threadIdx.x == 0) partials[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
}

constexpr int kTmaCells = 128;      // cells per tile of the TMA kernel
constexpr int kTmaStages = 3;
constexpr int kTmaComputeWarps = 16;

__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(tcx::smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(tcx::smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(tcx::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(tmap), "r"(c0), "r"(c1), "r"(tcx::smem_u32(src)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// One logit: returns w * bce, writes sigmoid and d(b_cls * cls_loss)/dx.  All terms derive from e = exp(-|x|):
//   sigmoid(x) and 1 - sigmoid(x) are {1, e} / (1 + e) (no cancellation on either side),
//   log1p(e) = 2 atanh(e / (2 + e)) as an odd series in s <= 1/3 (eight terms: below 2e-8 relative).
__device__ __forceinline__ float focal_term(float x, float t, float gamma, bool g2, float alpha, float grad_scale, float& p_out,
                                            float& g_out) {
  const float e = __expf(-fabsf(x));                             // in (0, 1]
  float inv, inv2;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.f + e));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv2) : "f"(2.f + e));
  const float q = e * inv;
  const bool nonneg = x >= 0.f;
  const float p = nonneg ? inv : q;                              // torch.sigmoid (:39)
  const float np1 = nonneg ? q : inv;                            // 1 - p
  const bool pos = t == 1.f;
  const float om = pos ? np1 : p;                                // 1 - pt (:40)
  const float wgt = (pos ? alpha : 1.f) * (g2 ? om * om : powf(om, gamma));   // (:41-43)
  const float sv = e * inv2, z = sv * sv;
  float poly = fmaf(z, 1.f / 15.f, 1.f / 13.f);
  poly = fmaf(poly, z, 1.f / 11.f);
  poly = fmaf(poly, z, 1.f / 9.f);
  poly = fmaf(poly, z, 1.f / 7.f);
  poly = fmaf(poly, z, 1.f / 5.f);
  poly = fmaf(poly, z, 1.f / 3.f);
  poly = fmaf(poly, z, 1.f);
  const float l1p = 2.f * sv * poly;                             // log1p(exp(-|x|))
  const float bce = fmaxf(x, 0.f) - x * t + l1p;                 // F.binary_cross_entropy_with_logits (:46)
  p_out = p;
  g_out = grad_scale * wgt * (p - t);                            // d(b_cls * cls_loss)/dx, weight detached
  return wgt * bce;
}

constexpr int kTmaRangeSlots = 128;  // list ranges kept in shared memory for the first 64 tiles of a block
constexpr int kMaxListSmem = 8192;   // positives whose anchor ids are kept in shared memory (else searched in global)

struct PosListIn {                   // sorted by global anchor id b*A + a (pp_assign_targets_list)
  const int* anchor;
  const float* cls;                  // [n, K]
  const int* offsets;                // [B+1], offsets[B] = n
  int Ad;
};

__device__ __forceinline__ int lower_bound_i(const int* __restrict__ v, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (v[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// SPARSE: the targets are implicit zeros plus the positives list; no target tile is loaded.  Tiles that hold
// positives evaluate those elements first (true t), mark them in a bitmap, and the main loop skips them.
template <bool SPARSE>
__global__ void __launch_bounds__((kTmaComputeWarps + 1) * 32, 1)
k_loss_cls_tma(const __grid_constant__ CUtensorMap tm_cls, const __grid_constant__ CUtensorMap tm_grad,
               const float* __restrict__ cls_t, PosListIn pl, int B, int plane, int CK, float gamma, float alpha,
               float grad_scale, float* __restrict__ scores, int want_grad, float* __restrict__ reg, int CR,
               double* __restrict__ partials) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ uint64_t full[kTmaStages], done[kTmaStages];
  __shared__ double s_red[kTmaComputeWarps];
  __shared__ unsigned s_bits[SPARSE ? kLossMaxCh * kTmaCells / 32 : 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int stage_floats = 2 * CK * kTmaCells;                  // X [CK][128] then T [128*CK]
  float* s_f = reinterpret_cast<float*>(s_raw);
  const int tiles_per_plane = (plane + kTmaCells - 1) / kTmaCells;
  const int ntiles = B * tiles_per_plane;
  if (tid == 0) {
    for (int i = 0; i < kTmaStages; ++i) {
      tcx::mbar_init(&full[i], 1);
      tcx::mbar_init(&done[i], kTmaComputeWarps);
    }
    tcx::fence_barrier_init();
  }
  __syncthreads();

  if (warp == kTmaComputeWarps) {
    // ---------------- producer: one thread issues two TMA loads and two TMA stores per tile
    if (lane != 0) return;
    auto load = [&](int it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int b = tile / tiles_per_plane, c0 = (tile - b * tiles_per_plane) * kTmaCells;
      const int nc = min(kTmaCells, plane - c0);
      const int st = it % kTmaStages;
      float* X = s_f + (size_t)st * stage_floats;
      float* T = X + CK * kTmaCells;
      // the box always lands whole (zero fill past the plane)
      tcx::mbar_expect_tx(&full[st], (uint32_t)(CK * kTmaCells + (SPARSE ? 0 : nc * CK)) * 4u);
      tma_load_2d(X, &tm_cls, c0, b * CK, &full[st]);
      if (!SPARSE) tcx::bulk_g2s(T, cls_t + ((size_t)b * plane + c0) * CK, (uint32_t)nc * CK * 4u, &full[st]);
    };
    const int my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    for (int it = 0; it < min(my_tiles, kTmaStages); ++it) load(it);
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % kTmaStages;
      tcx::mbar_wait(&done[st], (it / kTmaStages) & 1);
      const int tile = blockIdx.x + it * gridDim.x;
      const int b = tile / tiles_per_plane, c0 = (tile - b * tiles_per_plane) * kTmaCells;
      const int nc = min(kTmaCells, plane - c0);
      float* X = s_f + (size_t)st * stage_floats;
      float* T = X + CK * kTmaCells;
      if (want_grad) tma_store_2d(&tm_grad, X, c0, b * CK);       // columns past the plane are clipped
      if (scores != nullptr) bulk_s2g(scores + ((size_t)b * plane + c0) * CK, T, (uint32_t)nc * CK * 4u);
      bulk_commit();
      if (it + kTmaStages < my_tiles) {
        bulk_wait_read<0>();                                     // the stores have left shared memory
        load(it + kTmaStages);
      }
    }
    bulk_wait<0>();
    return;
  }

  // ---------------- compute warps
  const bool g2 = gamma == 2.f;
  constexpr int kStep = kTmaComputeWarps * 32;
  const int K = SPARSE ? CK / pl.Ad : 0;
  int n_list = 0;
  const int* ids = nullptr;
  unsigned short* s_mask = nullptr;                              // per listed row: bit k = (t_k == 1), bit 15 = other values
  __shared__ int s_range[SPARSE ? kTmaRangeSlots + 1 : 1];      // list range start of this block's tiles
  if (SPARSE) {
    n_list = pl.offsets[B];
    int* s_ids = reinterpret_cast<int*>(s_raw + (size_t)kTmaStages * 2 * CK * kTmaCells * sizeof(float));
    if (n_list <= kMaxListSmem) {
      s_mask = reinterpret_cast<unsigned short*>(s_ids + kMaxListSmem);
      for (int i = tid; i < n_list; i += kStep) {
        s_ids[i] = pl.anchor[i];
        unsigned m = 0;
        for (int k = 0; k < K; ++k) {
          const float t = __ldg(pl.cls + (size_t)i * K + k);
          if (t == 1.f && k < 15) m |= 1u << k; else if (t != 0.f) m |= 0x8000u;
        }
        s_mask[i] = (unsigned short)m;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kStep));
      ids = s_ids;
    } else {
      ids = pl.anchor;
    }
    // list ranges of this block's tiles, once: tile j of the block covers ids in [s_range[2j], s_range[2j+1])
    for (int j = tid; j < kTmaRangeSlots; j += kStep) {
      const int tile = blockIdx.x + (j >> 1) * gridDim.x;
      if (tile < ntiles) {
        const int b = tile / tiles_per_plane, c0 = (tile - b * tiles_per_plane) * kTmaCells;
        const int a0 = (b * plane + c0) * pl.Ad;
        s_range[j] = lower_bound_i(ids, n_list, (j & 1) ? a0 + min(kTmaCells, plane - c0) * pl.Ad : a0);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kStep));
  }
  double acc = 0.0;
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int b = tile / tiles_per_plane, c0 = (tile - b * tiles_per_plane) * kTmaCells;
    const int nc = min(kTmaCells, plane - c0);
    const int st = it % kTmaStages;
    float* X = s_f + (size_t)st * stage_floats;
    float* T = X + CK * kTmaCells;
    int lo = 0, hi = 0;
    if (SPARSE) {                                                // positives of this tile: ids in [a0, a1)
      if (2 * it + 1 < kTmaRangeSlots) {
        lo = s_range[2 * it];
        hi = s_range[2 * it + 1];
      } else {
        const int a0 = (b * plane + c0) * pl.Ad, a1 = a0 + nc * pl.Ad;
        if (lane == 0) { lo = lower_bound_i(ids, n_list, a0); hi = lower_bound_i(ids, n_list, a1); }
        lo = __shfl_sync(0xffffffffu, lo, 0);
        hi = __shfl_sync(0xffffffffu, hi, 0);
      }
    }
    tcx::mbar_wait(&full[st], (it / kTmaStages) & 1);
    float part = 0.f, part2 = 0.f;
    const bool marked = SPARSE && hi > lo;                       // uniform over the block
    if (marked) {
      const int a0 = (b * plane + c0) * pl.Ad;
      for (int i = tid; i < CK * kTmaCells / 32; i += kStep) s_bits[i] = 0u;
      asm volatile("bar.sync 1, %0;" ::"n"(kStep));
      for (int j = tid; j < (hi - lo) * K; j += kStep) {
        const int e = lo + j / K, k = j - (j / K) * K;
        float t;
        if (s_mask != nullptr && !(s_mask[e] & 0x8000u)) t = (s_mask[e] >> k) & 1u ? 1.f : 0.f;
        else t = __ldg(pl.cls + (size_t)e * K + k);
        if (t != 0.f) {
          const int local = ids[e] - a0, pc = local / pl.Ad, ch = (local - pc * pl.Ad) * K + k;
          const int idx = ch * kTmaCells + pc;
          float pa, ga;
          part += focal_term(X[idx], t, gamma, g2, alpha, grad_scale, pa, ga);
          T[pc * CK + ch] = pa;
          X[idx] = ga;
          atomicOr(&s_bits[idx >> 5], 1u << (idx & 31));
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kStep));
    }
    const int cell = tid & (kTmaCells - 1);                      // fixed per thread: kStep is a multiple of the tile width
    if (cell < nc) {
      const int total = CK * kTmaCells;
      int idx = tid;
      for (; idx + kStep < total; idx += 2 * kStep) {             // two independent elements per trip
        const int ta = cell * CK + (idx >> 7), tb = ta + kStep / kTmaCells;
        const bool skip_a = marked && ((s_bits[idx >> 5] >> (idx & 31)) & 1u);
        const bool skip_b = marked && ((s_bits[(idx + kStep) >> 5] >> ((idx + kStep) & 31)) & 1u);
        float pa, ga, pb, gb;
        const float va = focal_term(X[idx], SPARSE ? 0.f : T[ta], gamma, g2, alpha, grad_scale, pa, ga);
        const float vb = focal_term(X[idx + kStep], SPARSE ? 0.f : T[tb], gamma, g2, alpha, grad_scale, pb, gb);
        if (!skip_a) { part += va; T[ta] = pa; X[idx] = ga; }
        if (!skip_b) { part2 += vb; T[tb] = pb; X[idx + kStep] = gb; }
      }
      if (idx < total) {
        const int ta = cell * CK + (idx >> 7);
        const bool skip_a = marked && ((s_bits[idx >> 5] >> (idx & 31)) & 1u);
        float pa, ga;
        const float va = focal_term(X[idx], SPARSE ? 0.f : T[ta], gamma, g2, alpha, grad_scale, pa, ga);
        if (!skip_a) { part += va; T[ta] = pa; X[idx] = ga; }
      }
    }
    acc += (double)(part + part2);
    if (marked) asm volatile("bar.sync 1, %0;" ::"n"(kStep));     // the bitmap is free for the next marked tile
    tcx::fence_proxy_async();                                    // generic-proxy writes -> visible to the bulk stores
    __syncwarp();
    if (lane == 0) tcx::mbar_arrive(&done[st]);
  }
  // in-place tanh on network channel 6 of reg_out (model/loss.py:50)
  if (reg != nullptr) {
    const size_t n = (size_t)B * plane;
    for (size_t i = (size_t)blockIdx.x * (kTmaComputeWarps * 32) + tid; i < n; i += (size_t)gridDim.x * kTmaComputeWarps * 32) {
      const size_t b = i / plane, c = i - b * plane;
      float* q = reg + (b * CR + 6) * plane + c;
      *q = tanhf(*q);
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) s_red[warp] = acc;
  asm volatile("bar.sync 1, %0;" ::"n"(kTmaComputeWarps * 32));  // compute warps only (the producer has left)
  if (tid == 0) {
    double r = 0.0;
    for (int w = 0; w < kTmaComputeWarps; ++w) r += s_red[w];
    partials[blockIdx.x] = r;
  }
}

// in-place tanh on network channel 6 of reg_out (model/loss.py:50)
__global__ void __launch_bounds__(256) k_loss_tanh(float* __restrict__ reg, int B, size_t plane, int CR) {
  const size_t n = (size_t)B * plane;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const size_t b = i / plane, c = i - b * plane;
    float* p = reg + (b * CR + 6) * plane + c;
    *p = tanhf(*p);
  }
}

// one positive anchor row: smooth-L1 / orientation terms and its unscaled gradient row.  Kept out of line so
// the streaming loop of k_loss_reg stays small (positives are a few hundred per sweep).
__device__ __noinline__ void loss_reg_row(const float* __restrict__ reg, const float* __restrict__ tr, size_t i, size_t A,
                                          size_t plane, int Ad, int R, float* __restrict__ grad, double& sr, double& so) {
  const size_t b = i / A, a = i - b * A;
  const size_t cell = a / Ad;
  const int d = (int)(a - cell * Ad);
  const float* src = reg + (b * (size_t)Ad * R + (size_t)d * R) * plane + cell;
  float* dst = grad != nullptr ? grad + (b * (size_t)Ad * R + (size_t)d * R) * plane + cell : nullptr;
  for (int c = 0; c < 7; ++c) {
    const float x = src[(size_t)c * plane];
    const float df = x - __ldg(tr + 1 + c);
    const float ad = fabsf(df);
    sr += (double)(ad < 1.f ? 0.5f * df * df : ad - 0.5f);         // F.smooth_l1_loss, beta = 1 (:57)
    float g = ad < 1.f ? df : (df > 0.f ? 1.f : -1.f);
    if (d == 0 && c == 6) g *= 1.f - x * x;                        // x is already tanh(.): chain rule of (:50)
    if (dst != nullptr) dst[(size_t)c * plane] = g;
  }
  if (R > 7) {
    const float o = src[(size_t)7 * plane], ot = __ldg(tr + 8);
    so += (double)(fmaxf(o, 0.f) - o * ot + log1pf(expf(-fabsf(o))));                 // (:59-61)
    if (dst != nullptr) dst[(size_t)7 * plane] = 1.f / (1.f + expf(-o)) - ot;
  }
}

// positives: loss terms, UNSCALED gradient rows (grad_reg is zeroed by the caller) and the list of positive
// anchors.  A warp reads 128 consecutive target rows (1152 floats) as nine coalesced 16-byte loads per lane; the
// lane that holds a row's flag word handles that row.
__global__ void __launch_bounds__(256, 3) k_loss_reg(const float* __restrict__ reg, const float* __restrict__ reg_t,
                                                  int B, int H, int W, int Ad, int R, float* __restrict__ grad,
                                                  unsigned* __restrict__ n_pos, unsigned* __restrict__ list,
                                                  double* __restrict__ partials) {
  __shared__ double s_red[8];
  const size_t plane = (size_t)H * W;
  const size_t A = plane * Ad, n = A * B;
  const int lane = lane_id();
  const size_t nchunks = (n + 127) / 128, e_end = n * 9;          // 128 rows = 1152 floats = 9 float4 per lane
  const bool vec = ((uintptr_t)reg_t % 16) == 0;
  double sr = 0.0, so = 0.0;
  for (size_t ck = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5); ck < nchunks; ck += (size_t)gridDim.x * 8) {
    const size_t e0 = ck * 1152;
    float4 v[9];
    if (vec && e0 + 1152 <= e_end) {                               // nine independent 16-byte loads in flight per lane
#pragma unroll
      for (int j = 0; j < 9; ++j) v[j] = __ldg(reinterpret_cast<const float4*>(reg_t + e0) + lane + 32 * j);
    } else {
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const size_t e = e0 + 4 * (size_t)(lane + 32 * j);
        v[j].x = e < e_end ? __ldg(reg_t + e) : 0.f;
        v[j].y = e + 1 < e_end ? __ldg(reg_t + e + 1) : 0.f;
        v[j].z = e + 2 < e_end ? __ldg(reg_t + e + 2) : 0.f;
        v[j].w = e + 3 < e_end ? __ldg(reg_t + e + 3) : 0.f;
      }
    }
    unsigned long long hit = 0;                                    // bit 4j+k: word 4(lane+32j)+k is a flag word equal to 1
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int w = 4 * (lane + 32 * j), m = w % 9;                // (:54) where(reg_targets[...,0] == 1)
      const float f[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (f[k] == 1.f && (m + k == 0 || m + k == 9)) hit |= 1ull << (4 * j + k);
    }
    while (hit) {
      const int bit = __ffsll((long long)hit) - 1;
      hit &= hit - 1;
      const size_t i = ck * 128 + (4 * (lane + 32 * (bit >> 2)) + (bit & 3)) / 9;
      const unsigned slot = atomicAdd(n_pos, 1u);
      if (list != nullptr) list[slot] = (unsigned)i;
      loss_reg_row(reg, reg_t + i * 9, i, A, plane, Ad, R, grad, sr, so);
    }
  }
  const double a0 = block_sum(sr, s_red);
  const double a1 = block_sum(so, s_red);
  if (threadIdx.x == 0) {
    partials[(size_t)blockIdx.x * 2] = a0;
    partials[(size_t)blockIdx.x * 2 + 1] = a1;
  }
}

// the same pass from a positives list (pp_assign_targets_list): one thread per listed row
__global__ void __launch_bounds__(256) k_loss_reg_list(const float* __restrict__ reg, const int* __restrict__ pos_anchor,
                                                       const float* __restrict__ pos_reg, const int* __restrict__ pos_offsets,
                                                       int B, int H, int W, int Ad, int R, float* __restrict__ grad,
                                                       unsigned* __restrict__ n_pos, unsigned* __restrict__ list,
                                                       double* __restrict__ partials) {
  __shared__ double s_red[8];
  const size_t plane = (size_t)H * W, A = plane * Ad;
  const int n = pos_offsets[B];
  double sr = 0.0, so = 0.0;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < n; e += gridDim.x * 256) {
    const float* tr = pos_reg + (size_t)e * 9;
    if (__ldg(tr) != 1.f) continue;                                // (:54) pos_anchors = where(reg_targets[...,0] == 1)
    const size_t i = (size_t)pos_anchor[e];
    const unsigned slot = atomicAdd(n_pos, 1u);
    if (list != nullptr) list[slot] = (unsigned)i;
    loss_reg_row(reg, tr, i, A, plane, Ad, R, grad, sr, so);
  }
  const double a0 = block_sum(sr, s_red);
  const double a1 = block_sum(so, s_red);
  if (threadIdx.x == 0) {
    partials[(size_t)blockIdx.x * 2] = a0;
    partials[(size_t)blockIdx.x * 2 + 1] = a1;
  }
}

// Every block re-derives the three sums (a few thousand doubles, fixed order), block 0 writes the losses, and
// all blocks scale the listed gradient rows by b_reg / (7 n_pos) and b_ort / n_pos.
__global__ void __launch_bounds__(256) k_loss_finalize(const double* __restrict__ pc, int nc, const double* __restrict__ pr,
                                                       int nr, const unsigned* __restrict__ n_pos,
                                                       const unsigned* __restrict__ list, double n_cls, float b_cls,
                                                       float b_reg, float b_ort, int B, int H, int W, int Ad, int R,
                                                       float* __restrict__ grad, float* __restrict__ losses) {
  __shared__ double s_red[8];
  const unsigned np = *n_pos;
  if (blockIdx.x == 0) {
    double c = 0.0, r = 0.0, o = 0.0;
    for (int i = threadIdx.x; i < nc; i += 256) c += pc[i];
    for (int i = threadIdx.x; i < nr; i += 256) { r += pr[2 * i]; o += pr[2 * i + 1]; }
    const double C = block_sum(c, s_red), Rr = block_sum(r, s_red), O = block_sum(o, s_red);
    if (threadIdx.x == 0) {
      const double nan = __longlong_as_double(0x7ff8000000000000ll);
      const double cls_loss = C / n_cls;
      const double reg_loss = np > 0 ? Rr / (7.0 * np) : nan;     // torch: mean over an empty tensor
      const double ort_loss = np > 0 ? O / np : nan;
      losses[0] = (float)cls_loss;
      losses[1] = (float)reg_loss;
      losses[2] = (float)ort_loss;
      losses[3] = (float)((double)b_cls * cls_loss + (double)b_reg * reg_loss + (double)b_ort * ort_loss);   // (:63)
    }
  }
  if (grad == nullptr || list == nullptr || np == 0) return;
  const float inv_r = b_reg / (7.f * (float)np), inv_o = b_ort / (float)np;
  const size_t plane = (size_t)H * W, A = plane * Ad;
  const int rows = R > 7 ? 8 : 7;
  for (size_t q = (size_t)blockIdx.x * 256 + threadIdx.x; q < (size_t)np * rows; q += (size_t)gridDim.x * 256) {
    const size_t i = list[q / rows];
    const int c = (int)(q % rows);
    const size_t b = i / A, a = i - b * A, cell = a / Ad;
    const int d = (int)(a - cell * Ad);
    float* g = grad + (b * (size_t)Ad * R + (size_t)d * R + c) * plane + cell;
    *g *= c < 7 ? inv_r : inv_o;
  }
}

struct LossWs {
  double* pc;
  double* pr;
  unsigned* n_pos;
  unsigned* list;
};

static int loss_reg_blocks() { return sm_count() * 3; }
constexpr int kLossFinBlocks = 64;

template <class A>
static void loss_layout(A& a, LossWs* ws, int B, int H, int W, int Ad) {
  const size_t nc = (size_t)B * H * ((W + kLossCells - 1) / kLossCells);       // generic kernel's grid; >= #SMs
  auto p0 = a.template take<double>(nc > 1024 ? nc : 1024);
  auto p1 = a.template take<double>((size_t)loss_reg_blocks() * 2);
  auto p2 = a.template take<unsigned>(64);
  auto p3 = a.template take<unsigned>((size_t)B * H * W * Ad);
  if (ws) { ws->pc = p0; ws->pr = p1; ws->n_pos = p2; ws->list = p3; }
}

struct SizeArena4 {
  size_t used = 0;
  template <class T>
  T* take(size_t count) { used += align_up(count * sizeof(T)); return nullptr; }
};

__global__ void __launch_bounds__(256) k_loss_scale(float* __restrict__ a, size_t na, float* __restrict__ b, size_t nb,
                                                    const float* __restrict__ scale, const float* __restrict__ applied) {
  const float r = applied != nullptr ? *scale / *applied : *scale;
  if (r == 1.f) return;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < na + nb; i += (size_t)gridDim.x * 256) {
    if (i < na) a[i] *= r; else b[i - na] *= r;
  }
}

}  // namespace pp

extern "C" {

size_t pp_loss_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t anchors_per_cell) {
  if (B < 1 || H < 1 || W < 1 || anchors_per_cell < 1) return 0;
  pp::SizeArena4 a;
  pp::loss_layout(a, (pp::LossWs*)nullptr, B, H, W, anchors_per_cell);
  return a.used + pp::kAlign;
}

}  // extern "C"

namespace pp {
struct PosListArgs {
  const int* anchor;
  const float* cls;
  const float* reg;
  const int* offsets;
};

static int loss_impl(const float* d_cls_out, float* d_reg_out, const float* d_cls_t, const float* d_reg_t,
                     const PosListArgs* pl, int32_t B, int32_t H, int32_t W, int32_t anchors_per_cell, int32_t num_classes,
                     int32_t reg_dims, float gamma, float alpha_pos, float b_cls, float b_reg, float b_ort, float* d_scores,
                     float* d_grad_cls, float* d_grad_reg, float* d_losses, void* d_workspace, size_t workspace_bytes,
                     pp_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_cls_out || !d_reg_out || !d_losses || B < 1 || H < 1 || W < 1 || anchors_per_cell < 1 || num_classes < 1 ||
      reg_dims < 7)
    return PP_ERR_INVALID_ARG;
  if (pl ? (!pl->anchor || !pl->cls || !pl->reg || !pl->offsets) : (!d_cls_t || !d_reg_t)) return PP_ERR_INVALID_ARG;
  const int CK = anchors_per_cell * num_classes, CR = anchors_per_cell * reg_dims;
  const size_t plane = (size_t)H * W;
  const size_t n_anchors = plane * anchors_per_cell * B;
  if (CK > kLossMaxCh || CR <= 6 || B > 65535 || H > 65535 || n_anchors >= (1ull << 31) ||
      plane * B >= (1ull << 31))
    return PP_ERR_UNSUPPORTED;
  Arena arena(d_workspace, workspace_bytes);
  LossWs ws{};
  loss_layout(arena, &ws, B, H, W, anchors_per_cell);
  if (!arena.ok) return PP_ERR_WORKSPACE;
  const double n_cls = (double)n_anchors * num_classes;
  const float gscale = (float)((double)b_cls / n_cls);
  PP_CUDA(cudaMemsetAsync(ws.n_pos, 0, sizeof(unsigned), st));
  auto aligned16 = [](const void* q) { return q == nullptr || ((uintptr_t)q % 16) == 0; };
  const size_t smem = (size_t)kTmaStages * 2 * CK * kTmaCells * sizeof(float) + (pl ? kMaxListSmem * (sizeof(int) + sizeof(unsigned short)) : 0);
  const bool tma = (g_opt_loss_tma || pl) && tensor_map_encode_fn() != nullptr && plane % 4 == 0 && CK % 2 == 0 &&
                   aligned16(d_cls_out) && aligned16(d_cls_t) && aligned16(d_scores) && aligned16(d_grad_cls) &&
                   smem <= 216 * 1024;
  if (pl && !tma) return PP_ERR_UNSUPPORTED;       // the list form exists for the TMA kernel only: use the dense targets
  int nparts;
  if (tma) {
    const int tiles = B * (int)((plane + kTmaCells - 1) / kTmaCells);
    nparts = tiles < sm_count() ? tiles : sm_count();
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeFn enc = (EncodeFn)tensor_map_encode_fn();
    CUtensorMap tm_cls, tm_grad;
    const cuuint64_t gdim[2] = {(cuuint64_t)plane, (cuuint64_t)B * CK};
    const cuuint64_t gstride[1] = {(cuuint64_t)plane * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kTmaCells, (cuuint32_t)CK};
    const cuuint32_t estr[2] = {1u, 1u};
    auto encode = [&](CUtensorMap* m, const float* base) {
      return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    if (!encode(&tm_cls, d_cls_out) || !encode(&tm_grad, d_grad_cls != nullptr ? d_grad_cls : d_cls_out))
      return PP_ERR_UNSUPPORTED;
    const int threads = (kTmaComputeWarps + 1) * 32;
    if (pl) {
      PosListIn pin{pl->anchor, pl->cls, pl->offsets, anchors_per_cell};
      PP_CUDA(cudaFuncSetAttribute(k_loss_cls_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PP_KERNEL("k_loss_cls_list", st,
                (k_loss_cls_tma<true><<<nparts, threads, smem, st>>>(tm_cls, tm_grad, nullptr, pin, B, (int)plane, CK, gamma,
                                                                     alpha_pos, gscale, d_scores, d_grad_cls != nullptr ? 1 : 0,
                                                                     d_reg_out, CR, ws.pc)));
    } else {
      PosListIn pin{nullptr, nullptr, nullptr, anchors_per_cell};
      PP_CUDA(cudaFuncSetAttribute(k_loss_cls_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PP_KERNEL("k_loss_cls_tma", st,
                (k_loss_cls_tma<false><<<nparts, threads, smem, st>>>(tm_cls, tm_grad, d_cls_t, pin, B, (int)plane, CK, gamma,
                                                                      alpha_pos, gscale, d_scores, d_grad_cls != nullptr ? 1 : 0,
                                                                      d_reg_out, CR, ws.pc)));
    }
  } else {
    const dim3 gc((W + kLossCells - 1) / kLossCells, H, B);
    nparts = (int)((size_t)gc.x * gc.y * gc.z);
    PP_KERNEL("k_loss_cls", st,
              (k_loss_cls<<<gc, 256, 0, st>>>(d_cls_out, d_cls_t, H, W, CK, gamma, alpha_pos, gscale, d_scores, d_grad_cls, ws.pc)));
    PP_KERNEL("k_loss_tanh", st, (k_loss_tanh<<<(int)((B * plane + 255) / 256), 256, 0, st>>>(d_reg_out, B, plane, CR)));
  }
  int nb = loss_reg_blocks();
  if (d_grad_reg != nullptr) PP_CUDA(cudaMemsetAsync(d_grad_reg, 0, (size_t)B * CR * plane * sizeof(float), st));
  if (pl) {
    nb = 32;
    PP_KERNEL("k_loss_reg_list", st,
              (k_loss_reg_list<<<nb, 256, 0, st>>>(d_reg_out, pl->anchor, pl->reg, pl->offsets, B, H, W, anchors_per_cell,
                                                   reg_dims, d_grad_reg, ws.n_pos, d_grad_reg != nullptr ? ws.list : nullptr,
                                                   ws.pr)));
  } else {
    PP_KERNEL("k_loss_reg", st,
              (k_loss_reg<<<nb, 256, 0, st>>>(d_reg_out, d_reg_t, B, H, W, anchors_per_cell, reg_dims, d_grad_reg, ws.n_pos,
                                              d_grad_reg != nullptr ? ws.list : nullptr, ws.pr)));
  }
  PP_KERNEL("k_loss_finalize", st,
            (k_loss_finalize<<<kLossFinBlocks, 256, 0, st>>>(ws.pc, nparts, ws.pr, nb, ws.n_pos, ws.list, n_cls, b_cls, b_reg,
                                                             b_ort, B, H, W, anchors_per_cell, reg_dims, d_grad_reg,
                                                             d_losses)));
  return PP_OK;
}
}  // namespace pp

extern "C" {

int pp_loss(const float* d_cls_out, float* d_reg_out, const float* d_cls_t, const float* d_reg_t, int32_t B,
            int32_t H, int32_t W, int32_t anchors_per_cell, int32_t num_classes, int32_t reg_dims, float gamma,
            float alpha_pos, float b_cls, float b_reg, float b_ort, float* d_scores, float* d_grad_cls,
            float* d_grad_reg, float* d_losses, void* d_workspace, size_t workspace_bytes, pp_stream_t stream) {
  return pp::loss_impl(d_cls_out, d_reg_out, d_cls_t, d_reg_t, nullptr, B, H, W, anchors_per_cell, num_classes, reg_dims,
                       gamma, alpha_pos, b_cls, b_reg, b_ort, d_scores, d_grad_cls, d_grad_reg, d_losses, d_workspace,
                       workspace_bytes, stream);
}

int pp_loss_list(const float* d_cls_out, float* d_reg_out, const int32_t* d_pos_anchor, const float* d_pos_cls,
                 const float* d_pos_reg, const int32_t* d_pos_offsets, int32_t B, int32_t H, int32_t W,
                 int32_t anchors_per_cell, int32_t num_classes, int32_t reg_dims, float gamma, float alpha_pos,
                 float b_cls, float b_reg, float b_ort, float* d_scores, float* d_grad_cls, float* d_grad_reg,
                 float* d_losses, void* d_workspace, size_t workspace_bytes, pp_stream_t stream) {
  pp::PosListArgs pl{d_pos_anchor, d_pos_cls, d_pos_reg, d_pos_offsets};
  return pp::loss_impl(d_cls_out, d_reg_out, nullptr, nullptr, &pl, B, H, W, anchors_per_cell, num_classes, reg_dims,
                       gamma, alpha_pos, b_cls, b_reg, b_ort, d_scores, d_grad_cls, d_grad_reg, d_losses, d_workspace,
                       workspace_bytes, stream);
}

int pp_loss_scale_grads(float* d_grad_cls, size_t n_cls, float* d_grad_reg, size_t n_reg, const float* d_scale,
                        const float* d_applied, pp_stream_t stream) {
  using namespace pp;
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_scale || (!d_grad_cls && n_cls) || (!d_grad_reg && n_reg)) return PP_ERR_INVALID_ARG;
  if (n_cls + n_reg == 0) return PP_OK;
  PP_KERNEL("k_loss_scale", st,
            (k_loss_scale<<<sm_count() * 8, 256, 0, st>>>(d_grad_cls, n_cls, d_grad_reg, n_reg, d_scale, d_applied)));
  return PP_OK;
}

}  // extern "C"
