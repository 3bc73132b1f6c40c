// targets.cu -- K3: rotated anchor-vs-GT IoU and target assignment for sm_100a.
// Replaces data/pillars.cpp:132-172 (iou), :400-427 (make_ious) and utils/box_utils.py:70-109,
// 162-232 (make_target, create_target) of the reference.
//
// IoU: the anchor ring (counter-clockwise) is clipped against the four half-planes of the GT
// ring (clockwise, walked backwards) with Sutherland-Hodgman in fp64; every multiply, add and
// divide is separately rounded (__dmul_rn/__dadd_rn/__ddiv_rn, no FMA contraction) so the value is
// the one a plain IEEE-double CPU evaluation of the same formulae gives.
//
// Assignment never materialises the [A,G] matrix.  A uniform bucket index over anchor centres
// (built once per anchor set) lets each GT visit only the anchors that can pass the centre
// prefilter (|dcx|<=10 and |dcy|<=10, data/pillars.cpp:418-419).  Exact fp64 comparisons and
// first-index tie-breaks of np.argmax are reproduced with 64-bit atomicMax on the IoU bit
// pattern followed by atomicMin on the index among the maximisers:
//   k_iou_pass<0>  one CTA per GT: IoU of its candidates -> atomicMax(best[a]); CTA reduction ->
//                  top_anchor[g] (first anchor among the GT's maximisers; 0 when all IoU are 0)
//   k_iou_pass<1>  same enumeration: candidates whose IoU equals best[a] -> atomicMin(arg[a], g);
//                  positives (best > thresh) set a bit in posmask
//   k_encode       dense, vectorised write of cls[A,K] and reg[A,9] (zeros + positives)
//                  (forced per-GT best-anchor overrides are resolved inline, in the reference's order)
#include <vector>

#include "common.cuh"

namespace pp {

extern int g_opt_encode_bulk;
constexpr double kRadius = 10.0;  // data/pillars.cpp:418-419, hard-coded in the reference
constexpr int kCandCap = 2048;    // per-GT IoU cache entries handed from pass 0 to pass 1 (overflow is recomputed)

struct IndexHeader {  // first 64 bytes of the anchor index
  double x0, y0, cell;
  int nbx, nby;
  long long A;
  int max_bucket;
  int magic;
  long long geom_off;   // byte offset of the bucket-ordered geometry copy (0: none): [A][10] doubles = centre x,y + 4 corners
  long long pad;
};
static_assert(sizeof(IndexHeader) == 64, "index header must be 64 bytes");
constexpr int kIndexMagic = 0x50504958;

// ---- geometry ------------------------------------------------------------------------------
__device__ __forceinline__ double ring_area_ccw(const double* v, int n) {
  if (n < 3) return 0.0;
  double s = 0.0;
  for (int i = 0; i < n; ++i) {
    const int j = (i + 1 == n) ? 0 : i + 1;
    s = __dadd_rn(s, __dsub_rn(__dmul_rn(v[2 * i], v[2 * j + 1]), __dmul_rn(v[2 * j], v[2 * i + 1])));
  }
  return __dmul_rn(0.5, s);
}

// a: 4 corners counter-clockwise; g: 4 corners clockwise (both open rings, x,y interleaved)
__device__ double quad_iou(const double* __restrict__ a, const double* __restrict__ g) {
  double buf0[16], buf1[16];
  double* subj = buf0;
  double* nxt = buf1;
  int n = 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) subj[i] = a[i];
  for (int e = 0; e < 4 && n > 0; ++e) {
    const double c0x = g[2 * (3 - e)], c0y = g[2 * (3 - e) + 1];
    const int e1 = (6 - e) & 3;
    const double ex = __dsub_rn(g[2 * e1], c0x), ey = __dsub_rn(g[2 * e1 + 1], c0y);
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const int ip = (i == 0) ? n - 1 : i - 1;
      const double cx = subj[2 * i], cy = subj[2 * i + 1];
      const double px = subj[2 * ip], py = subj[2 * ip + 1];
      const double dc = __dsub_rn(__dmul_rn(ex, __dsub_rn(cy, c0y)), __dmul_rn(ey, __dsub_rn(cx, c0x)));
      const double dp = __dsub_rn(__dmul_rn(ex, __dsub_rn(py, c0y)), __dmul_rn(ey, __dsub_rn(px, c0x)));
      const bool cin = dc >= 0.0, pin = dp >= 0.0;
      if (cin != pin) {
        const double t = __ddiv_rn(dp, __dsub_rn(dp, dc));
        nxt[2 * m] = __dadd_rn(px, __dmul_rn(t, __dsub_rn(cx, px)));
        nxt[2 * m + 1] = __dadd_rn(py, __dmul_rn(t, __dsub_rn(cy, py)));
        ++m;
      }
      if (cin) {
        nxt[2 * m] = cx;
        nxt[2 * m + 1] = cy;
        ++m;
      }
    }
    double* t2 = subj; subj = nxt; nxt = t2;
    n = m;
  }
  if (n < 3) return 0.0;
  // the reference takes bg::area of a clockwise-typed output polygon (data/pillars.cpp:15,164):
  // walk the intersection ring clockwise and flip the sign
  for (int i = 0; i < n; ++i) {
    nxt[2 * i] = subj[2 * (n - 1 - i)];
    nxt[2 * i + 1] = subj[2 * (n - 1 - i) + 1];
  }
  const double inter = -ring_area_ccw(nxt, n);
  if (!(inter > 0.0)) return 0.0;
  const double area_a = ring_area_ccw(a, 4);
  const double area_g = -ring_area_ccw(g, 4);
  return __ddiv_rn(inter, __dsub_rn(__dadd_rn(area_a, area_g), inter));
}

__device__ __forceinline__ bool prefilter_far(const double* __restrict__ ac, const double* __restrict__ gc) {
  return (fabs(__dsub_rn(ac[0], gc[0])) > kRadius) || (fabs(__dsub_rn(ac[1], gc[1])) > kRadius);
}

// ---- dense make_ious ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_make_ious(const double* __restrict__ a_corners,
                                                   const double* __restrict__ g_corners,
                                                   const double* __restrict__ a_centers,
                                                   const double* __restrict__ g_centers,
                                                   long long A, long long G,
                                                   double* __restrict__ ious,
                                                   int* __restrict__ status) {
  const long long total = A * G;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / G, j = idx % G;
    double v = 0.0;
    if (!prefilter_far(a_centers + i * 3, g_centers + j * 3)) {
      double a[8], g[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { a[k] = a_corners[i * 8 + k]; g[k] = g_corners[j * 8 + k]; }
      v = quad_iou(a, g);
      if (v < 0.0) atomicOr(status, PP_STATUS_NEG_IOU);
    }
    ious[idx] = v;
  }
}

// ---- sparse assignment ----------------------------------------------------------------------
struct GtParams {
  int n_sweeps;
  long long off[PP_MAX_SWEEPS + 1];
};

__device__ __forceinline__ int find_gt_sweep(const GtParams& gp, long long g) {
  int lo = 0, hi = gp.n_sweeps - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (gp.off[mid] <= g) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// PASS 0: best[a] = max IoU (as bits), top_anchor[g].  PASS 1: arg[a], posmask, counters.
// (256, 3): at most 85 registers, so that the 400 CTAs of a batch-4 step (one per GT) are one resident wave on
// 148 SMs; at 100 registers only two CTAs fit an SM and the kernel ran as 1.35 waves
template <int PASS>
__global__ void __launch_bounds__(256, 3) k_iou_pass(
    const double* __restrict__ a_corners, const double* __restrict__ a_centers,
    const unsigned char* __restrict__ index, long long A, const double* __restrict__ g_corners,
    const double* __restrict__ g_centers, GtParams gp, double pos_thresh,
    unsigned long long* __restrict__ best, int* __restrict__ arg, unsigned* __restrict__ posmask,
    unsigned* __restrict__ forcedmask, int* __restrict__ top_anchor, int* __restrict__ counts,
    double* __restrict__ cand_iou, int* __restrict__ status) {
  const long long gg = blockIdx.x;  // global GT row
  const int b = find_gt_sweep(gp, gg);
  const int gl = (int)(gg - gp.off[b]);  // index of the GT inside its sweep
  const IndexHeader* hdr = reinterpret_cast<const IndexHeader*>(index);
  const int* bucket_start = reinterpret_cast<const int*>(index + sizeof(IndexHeader));
  const int nb = hdr->nbx * hdr->nby;
  const int* ids = bucket_start + nb + 1;
  // bucket-ordered copy of the anchor geometry (centre x,y + corners): candidate k of a bucket row reads
  // 80 contiguous bytes next to its neighbours' instead of two dependent gathers from the [A,...] arrays
  const double* geom = hdr->geom_off != 0 ? reinterpret_cast<const double*>(index + hdr->geom_off) : nullptr;

  double g[8], gc[2];
#pragma unroll
  for (int k = 0; k < 8; ++k) g[k] = g_corners[gg * 8 + k];
  gc[0] = g_centers[gg * 3];
  gc[1] = g_centers[gg * 3 + 1];

  // bucket window that can contain anchors passing the prefilter (one bucket of slack each side)
  int bx0 = (int)floor((gc[0] - kRadius - hdr->x0) / hdr->cell) - 1;
  int bx1 = (int)floor((gc[0] + kRadius - hdr->x0) / hdr->cell) + 1;
  int by0 = (int)floor((gc[1] - kRadius - hdr->y0) / hdr->cell) - 1;
  int by1 = (int)floor((gc[1] + kRadius - hdr->y0) / hdr->cell) + 1;
  bx0 = max(bx0, 0); by0 = max(by0, 0);
  bx1 = min(bx1, hdr->nbx - 1); by1 = min(by1, hdr->nby - 1);
  if (!(gc[0] == gc[0]) || !(gc[1] == gc[1])) { bx1 = -1; by1 = -1; }  // NaN centre: out of contract

  unsigned long long my_best = 0ull;
  int my_a = 0x7fffffff;
  unsigned long long* best_b = best + (size_t)b * A;
  if (PASS == 1 && threadIdx.x == 0) {
    // per-GT best anchor (computed by pass 0); 0 means "dropped" (utils/box_utils.py:204)
    const int t = top_anchor[gg];
    if (t != 0) {
      atomicOr(&forcedmask[(size_t)b * ((A + 31) / 32) + (t >> 5)], 1u << (t & 31));
      atomicAdd(&counts[b * 4 + 1], 1);
    }
  }

  double* cache = cand_iou + (size_t)gg * kCandCap;
  // GT bounding box: an anchor whose bounding box is strictly separated from it cannot intersect the
  // GT polygon, the reference's bg::intersection output is empty and its IoU is exactly 0
  // (data/pillars.cpp:161-163) -- no clipping needed for those
  double gx0 = g[0], gx1 = g[0], gy0 = g[1], gy1 = g[1];
#pragma unroll
  for (int q = 1; q < 4; ++q) {
    gx0 = fmin(gx0, g[2 * q]); gx1 = fmax(gx1, g[2 * q]);
    gy0 = fmin(gy0, g[2 * q + 1]); gy1 = fmax(gy1, g[2 * q + 1]);
  }
  // PASS 0 runs in two phases so that every lane clips a polygon pair: phase 1 classifies the window's
  // candidates (centre prefilter + bounding boxes; ~3 of 4 drop out) and queues the survivors in
  // shared memory, phase 2 deals the queue round-robin to the threads.  Clipping inline left most
  // lanes of a warp idle behind the few that passed (ncu r1h: 22 % issue utilisation, 82 us).
  constexpr int kQueue = 2048;
  __shared__ int q_a[PASS == 0 ? kQueue : 1];
  __shared__ int q_slot[PASS == 0 ? kQueue : 1];
  __shared__ int q_k[PASS == 0 ? kQueue : 1];
  __shared__ int q_n;
  if (PASS == 0) {
    if (threadIdx.x == 0) q_n = 0;
    __syncthreads();
  }
  auto survives = [&](int a, int k) -> bool {
    const double* ac = geom != nullptr ? geom + (size_t)k * 10 : a_centers + (size_t)a * 3;
    if (prefilter_far(ac, gc)) return false;
    const double* ar = geom != nullptr ? geom + (size_t)k * 10 + 2 : a_corners + (size_t)a * 8;
    double ax0 = ar[0], ax1 = ar[0], ay0 = ar[1], ay1 = ar[1];
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      ax0 = fmin(ax0, ar[2 * q]); ax1 = fmax(ax1, ar[2 * q]);
      ay0 = fmin(ay0, ar[2 * q + 1]); ay1 = fmax(ay1, ar[2 * q + 1]);
    }
    return !(ax1 < gx0 || gx1 < ax0 || ay1 < gy0 || gy1 < ay0);
  };
  auto clip = [&](int a, int k) -> double {
    double ar[8];
    const double* src = geom != nullptr ? geom + (size_t)k * 10 + 2 : a_corners + (size_t)a * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) ar[q] = src[q];
    double v = quad_iou(ar, g);
    if (v < 0.0) { atomicOr(status, PP_STATUS_NEG_IOU); v = 0.0; }
    return v;
  };
  auto account0 = [&](int a, double v) {        // PASS 0 bookkeeping of one positive IoU
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    atomicMax(&best_b[a], bits);
    if (bits > my_best || (bits == my_best && a < my_a)) { my_best = bits; my_a = a; }
  };
  int rowbase = 0;                                   // enumeration index of the first candidate of this bucket row
  for (int by = by0; by <= by1; ++by) {
    if (bx1 < bx0) break;
    const int s = bucket_start[by * hdr->nbx + bx0];
    const int e = bucket_start[by * hdr->nbx + bx1 + 1];
    for (int k = s + (int)threadIdx.x; k < e; k += blockDim.x) {
      const int a = ids[k];
      const int slot = rowbase + (k - s);
      if (PASS == 0) {
        if (survives(a, k)) {
          const int pos = atomicAdd(&q_n, 1);
          if (pos < kQueue) {
            q_a[pos] = a;
            q_slot[pos] = slot;
            q_k[pos] = k;
          } else {                                   // queue full (cannot happen with the reference's anchor grid): inline
            const double v = clip(a, k);
            if (slot < kCandCap) cache[slot] = v;
            if (v > 0.0) account0(a, v);
          }
        } else if (slot < kCandCap) {
          cache[slot] = 0.0;
        }
      } else {
        double v;
        if (slot < kCandCap) v = cache[slot];        // computed by pass 0
        else v = survives(a, k) ? clip(a, k) : 0.0;
        if (!(v > 0.0)) continue;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
        if (bits == best_b[a]) {
          atomicMin(&arg[(size_t)b * A + a], gl);
          if (v > pos_thresh) {
            const unsigned bit = 1u << (a & 31);
            const unsigned old = atomicOr(&posmask[(size_t)b * ((A + 31) / 32) + (a >> 5)], bit);
            if (!(old & bit)) atomicAdd(&counts[b * 4 + 0], 1);
          }
          if (fabs(v - pos_thresh) < 1e-6) atomicAdd(&counts[b * 4 + 2], 1);
        }
      }
    }
    rowbase += e - s;
  }
  if (PASS == 0) {
    __syncthreads();
    const int nq = min(q_n, kQueue);
    for (int i = (int)threadIdx.x; i < nq; i += blockDim.x) {
      const int a = q_a[i];
      const double v = clip(a, q_k[i]);
      if (q_slot[i] < kCandCap) cache[q_slot[i]] = v;
      if (v > 0.0) account0(a, v);
    }
  }

  if (PASS == 0) {
    // CTA reduction: max IoU bits, then min anchor index (np.argmax first-index rule, :200)
    __shared__ unsigned long long s_best[8];
    __shared__ int s_a[8];
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ob = __shfl_xor_sync(0xffffffffu, my_best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, my_a, o);
      if (ob > my_best || (ob == my_best && oa < my_a)) { my_best = ob; my_a = oa; }
    }
    if (lane_id() == 0) { s_best[threadIdx.x >> 5] = my_best; s_a[threadIdx.x >> 5] = my_a; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
        if (s_best[w] > my_best || (s_best[w] == my_best && s_a[w] < my_a)) { my_best = s_best[w]; my_a = s_a[w]; }
      }
      // all-zero column -> np.argmax returns 0 -> dropped by np.nonzero (utils/box_utils.py:204)
      top_anchor[gg] = (my_best == 0ull) ? 0 : my_a;
    }
  }
}

// utils/box_utils.py:70-109 in fp64.  gc is the IMAGE-space GT centre (y already flipped, which is
// what :83 recomputes); gyaw is the unflipped yaw.
__device__ void make_target(const double* __restrict__ ac, const double* __restrict__ awlh, double at,
                            const double* __restrict__ gc, const double* __restrict__ gwlh,
                            double gyaw, double out[9]) {
  const double aw = awlh[0], al = awlh[1], ah = awlh[2];
  const double ad = sqrt(__dadd_rn(__dmul_rn(aw, aw), __dmul_rn(al, al)));
  out[0] = 1.0;
  out[1] = __ddiv_rn(__dsub_rn(gc[0], ac[0]), ad);
  out[2] = __ddiv_rn(__dsub_rn(gc[1], ac[1]), ad);
  out[3] = __ddiv_rn(__dsub_rn(gc[2], ac[2]), ah);
  out[4] = log(__ddiv_rn(gwlh[0], aw));
  out[5] = log(__ddiv_rn(gwlh[1], al));
  out[6] = log(__ddiv_rn(gwlh[2], ah));
  const double kPi = 3.141592653589793;  // np.pi
  double gt = gyaw;
  if (gt <= kPi && gt >= kPi / 2) gt = __dsub_rn(gt, kPi);
  else if (gt >= -kPi && gt <= -kPi / 2) gt = __dadd_rn(gt, kPi);
  const double dth = __dsub_rn(gt, at);
  out[7] = sin(dth);
  out[8] = ((dth <= kPi && dth >= kPi / 2) || (dth >= -kPi && dth <= -kPi / 2)) ? 1.0 : 0.0;
}

struct EncodeArgs {
  const double *a_centers, *a_wlh, *a_yaw, *g_centers, *g_wlh, *g_yaw;
  const int* g_cls;
  const int* arg;
  const unsigned* posmask;
  const unsigned* forcedmask;
  const int* top_anchor;
};

// bit 0: positive by threshold, bit 1: forced (best anchor of some kept GT)
__device__ __forceinline__ unsigned anchor_flags(const EncodeArgs& ea, int b, long long A, long long a) {
  const size_t w = (size_t)b * ((A + 31) / 32) + (a >> 5);
  return ((ea.posmask[w] >> (a & 31)) & 1u) | (((ea.forcedmask[w] >> (a & 31)) & 1u) << 1);
}

__device__ float encode_value(const EncodeArgs& ea, const GtParams& gp, int b, long long A, long long a,
                              int col, bool is_reg, unsigned flags) {
  long long gg;
  if (flags & 2u) {
    // forced match (utils/box_utils.py:212-213,226-228): the row is cleared, every kept GT whose best
    // anchor is `a` sets its class bit, the LAST such GT writes the regression row
    gg = -1;
    bool hit = false;
    for (long long h = gp.off[b]; h < gp.off[b + 1]; ++h) {
      if (ea.top_anchor[h] == (int)a) {
        gg = h;
        hit = hit || (ea.g_cls[h] == col);
      }
    }
    if (!is_reg) return hit ? 1.f : 0.f;
  } else {
    gg = gp.off[b] + ea.arg[(size_t)b * A + a];
    if (!is_reg) return ea.g_cls[gg] == col ? 1.f : 0.f;   // utils/box_utils.py:211
  }
  double t[9];
  make_target(ea.a_centers + a * 3, ea.a_wlh + a * 3, ea.a_yaw[a], ea.g_centers + gg * 3,
              ea.g_wlh + gg * 3, ea.g_yaw[gg], t);
  return (float)t[col];                                   // utils/box_utils.py:219-221, then .float()
}

// dense write of one [A, ncol] float matrix per sweep (cls when !IS_REG, reg when IS_REG).
// A CTA owns tiles of 1024 anchors.  The positive / forced bitmasks say whether the tile holds any
// non-zero row: >95 % of tiles do not and are pure 16-byte zero stores with no per-element logic.
// NCOL > 0 fixes the row width at compile time (9 for both outputs of the reference config) so the
// element -> (anchor, column) split in the rare path is a multiply-shift, not a 64-bit division.
constexpr int kEncTile = 1024;   // anchors per tile (32 mask words)
template <bool IS_REG, int NCOL>
__global__ void __launch_bounds__(256) k_encode(EncodeArgs ea, GtParams gp, long long A, int ncol_rt,
                                                bool vec_ok, float* __restrict__ out) {
  const unsigned ncol = NCOL > 0 ? (unsigned)NCOL : (unsigned)ncol_rt;
  const int b = blockIdx.y;
  const unsigned nA = (unsigned)A;
  const unsigned total = nA * ncol;                   // host guarantees A*ncol < 2^31
  float* ob = out + (size_t)b * total;
  const unsigned ntiles = (nA + kEncTile - 1) / kEncTile;
  const size_t mw = (size_t)b * ((A + 31) / 32);
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const unsigned a0 = tile * kEncTile;
    const unsigned a1 = min(a0 + kEncTile, nA);
    // any flagged anchor in this tile?  (32 mask words, read by the first warp's lanes)
    unsigned w = 0;
    if (threadIdx.x < 32) {
      const unsigned wi = (a0 >> 5) + threadIdx.x;
      if (wi * 32 < a1) w = ea.posmask[mw + wi] | ea.forcedmask[mw + wi];
    }
    const int any = __syncthreads_or(w != 0u);
    const unsigned e0 = a0 * ncol, e1 = a1 * ncol;    // element range of the tile
    if (!any && vec_ok && (e0 & 3u) == 0u) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      float4* o4 = reinterpret_cast<float4*>(ob + e0);
      const unsigned n4 = (e1 - e0) >> 2;
      for (unsigned i = threadIdx.x; i < n4; i += blockDim.x) __stcs(o4 + i, z);
      for (unsigned e = e0 + (n4 << 2) + threadIdx.x; e < e1; e += blockDim.x) ob[e] = 0.f;
    } else {
      for (unsigned e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const unsigned a = e / ncol;
        float v = 0.f;
        const unsigned fl = any ? anchor_flags(ea, b, A, a) : 0u;
        if (fl) v = encode_value(ea, gp, b, A, a, (int)(e - a * ncol), IS_REG, fl);
        ob[e] = v;
      }
    }
  }
}

// Both output rows of one flagged anchor (same rules as encode_value, evaluated once per row).
__device__ void encode_row(const EncodeArgs& ea, const GtParams& gp, int b, long long A, long long a,
                           unsigned flags, float crow[9], float rrow[9]) {
#pragma unroll
  for (int c = 0; c < 9; ++c) crow[c] = 0.f;
  long long gg;
  if (flags & 2u) {
    // forced match (utils/box_utils.py:212-213,226-228): the row is cleared, every kept GT whose best
    // anchor is `a` sets its class bit, the LAST such GT writes the regression row
    // the per-GT best anchors are read eight at a time: one dependent L2 round trip per GT made this
    // loop ~60k cycles for 100 GT boxes
    gg = -1;
    const long long h_end = gp.off[b + 1];
    for (long long h0 = gp.off[b]; h0 < h_end; h0 += 8) {
      int ta[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) ta[j] = (h0 + j < h_end) ? ea.top_anchor[h0 + j] : -1;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (ta[j] == (int)a) {
          gg = h0 + j;
          const int k = ea.g_cls[gg];
          if (k >= 0 && k < 9) crow[k] = 1.f;
        }
      }
    }
  } else {
    gg = gp.off[b] + ea.arg[(size_t)b * A + a];
    const int k = ea.g_cls[gg];
    if (k >= 0 && k < 9) crow[k] = 1.f;                    // utils/box_utils.py:211
  }
  double t[9];
  make_target(ea.a_centers + a * 3, ea.a_wlh + a * 3, ea.a_yaw[a], ea.g_centers + gg * 3,
              ea.g_wlh + gg * 3, ea.g_yaw[gg], t);
#pragma unroll
  for (int c = 0; c < 9; ++c) rrow[c] = (float)t[c];      // utils/box_utils.py:219-221, then .float()
}

// cls [B,A,9] and reg [B,A,9]: both tensors are > 99.9 % zeros, so they are written as one plain zero
// stream (k_encode_zero: grid-stride float4 stores, memset speed) followed by k_encode_patch, which
// walks the positive / forced bit masks and writes the two 9-float rows of each flagged anchor
// (~150 per sweep).  History: per-element logic in the store loop made flagged tiles instruction-bound
// (38 % of the HBM peak); zero-fill + in-CTA patch left 250 threads of a tile waiting at a barrier for
// the few that evaluate make_target (ncu r1u: 61 % of the samples in that barrier, still 2.6 TB/s).
__global__ void __launch_bounds__(256) k_encode_zero(float4* __restrict__ cls4, float4* __restrict__ reg4, size_t n4) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    __stcs(cls4 + i, z);
    __stcs(reg4 + i, z);
  }
}

// The same zero stream as bulk copies: each block zeroes a 32 KB shared-memory tile once and one thread streams it
// out with cp.async.bulk (shared -> global, up to 8 copies of 32 KB in flight): no per-element store instructions,
// full 32 KB bursts per request.  Both byte counts are multiples of 16.
constexpr int kZeroTile = 32768;
__global__ void __launch_bounds__(128) k_encode_zero_bulk(unsigned char* __restrict__ cls, unsigned char* __restrict__ reg,
                                                          size_t bytes_each) {
  __shared__ __align__(128) unsigned char s_zero[kZeroTile];
  for (int i = threadIdx.x; i < kZeroTile / 16; i += blockDim.x) reinterpret_cast<uint4*>(s_zero)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  const size_t chunks_each = (bytes_each + kZeroTile - 1) / kZeroTile;
  const unsigned src = (unsigned)__cvta_generic_to_shared(s_zero);
  for (size_t c = blockIdx.x; c < 2 * chunks_each; c += gridDim.x) {
    unsigned char* base = c < chunks_each ? cls : reg;
    const size_t off = (c < chunks_each ? c : c - chunks_each) * (size_t)kZeroTile;
    const unsigned n = (unsigned)(bytes_each - off < (size_t)kZeroTile ? bytes_each - off : (size_t)kZeroTile);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off), "r"(src), "r"(n) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(128) k_encode_patch(EncodeArgs ea, GtParams gp, long long A, int B,
                                                      float* __restrict__ cls, float* __restrict__ reg) {
  const size_t words_per_sweep = (size_t)((A + 31) / 32);
  const size_t nwords = words_per_sweep * B;
  const size_t nw_round = (nwords + 31) / 32 * 32;     // whole warps iterate together (ballot / shuffle below)
  for (size_t wi = (size_t)blockIdx.x * blockDim.x + threadIdx.x; wi < nw_round; wi += (size_t)gridDim.x * blockDim.x) {
    const unsigned w = wi < nwords ? (ea.posmask[wi] | ea.forcedmask[wi]) : 0u;
    // positives cluster (the anchors around one GT share a mask word): the 32 anchors of a flagged
    // word are dealt to the 32 lanes instead of being walked by the one lane that loaded the word
    unsigned todo = __ballot_sync(0xffffffffu, w != 0u);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const unsigned ws = __shfl_sync(0xffffffffu, w, src);
      const size_t wsrc = wi - lane_id() + src;
      if ((ws >> lane_id()) & 1u) {
        const int b = (int)(wsrc / words_per_sweep);
        const long long a = (long long)(wsrc - (size_t)b * words_per_sweep) * 32 + lane_id();
        if (a < A) {
          const unsigned fl = anchor_flags(ea, b, A, a);
          float crow[9], rrow[9];
          encode_row(ea, gp, b, A, a, fl, crow, rrow);
          float* oc = cls + ((size_t)b * A + a) * 9;
          float* orr = reg + ((size_t)b * A + a) * 9;
#pragma unroll
          for (int c = 0; c < 9; ++c) { oc[c] = crow[c]; orr[c] = rrow[c]; }
        }
      }
    }
  }
}

// Positives list ("emit positives list + implicit zeros"): the flagged anchors of all sweeps in
// ascending (sweep, anchor) order with their two 9-float rows, instead of the two dense [B,A,9] tensors.
// Slots come from a prefix sum over the popcounts of the positive / forced mask words (k_pos_count: one sum per
// block of 256 words; k_pos_emit: block base + in-block scan), so the order is deterministic; the 32 anchors of
// a flagged word are dealt to the lanes of a warp as in k_encode_patch.
__global__ void __launch_bounds__(256) k_pos_count(const unsigned* __restrict__ posmask,
                                                   const unsigned* __restrict__ forcedmask, size_t nwords,
                                                   int* __restrict__ blocksum) {
  __shared__ int s_red[8];
  const size_t wi = (size_t)blockIdx.x * 256 + threadIdx.x;
  int c = wi < nwords ? __popc(posmask[wi] | forcedmask[wi]) : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane_id() == 0) s_red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    blocksum[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) k_pos_emit(EncodeArgs ea, GtParams gp, long long A, int B,
                                                  const int* __restrict__ blocksum, int* __restrict__ pos_anchor,
                                                  float* __restrict__ pos_cls, float* __restrict__ pos_reg,
                                                  int* __restrict__ pos_offsets, int cap, int* __restrict__ status) {
  __shared__ int s_red[8];
  __shared__ int s_base;
  const int lane = (int)lane_id(), warp = threadIdx.x >> 5;
  int part = 0;
  for (int j = threadIdx.x; j < (int)blockIdx.x; j += 256) part += blocksum[j];
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) s_red[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    s_base = t;
  }
  __syncthreads();
  const size_t words_per_sweep = (size_t)((A + 31) / 32);
  const size_t nwords = words_per_sweep * B;
  const size_t wi = (size_t)blockIdx.x * 256 + threadIdx.x;
  const unsigned w = wi < nwords ? (ea.posmask[wi] | ea.forcedmask[wi]) : 0u;
  const int c = __popc(w);
  int incl = c;                                                  // inclusive scan inside the warp
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  __syncthreads();                                               // s_red is reused
  if (lane == 31) s_red[warp] = incl;
  __syncthreads();
  int wbase = s_base + incl - c;
  for (int q = 0; q < warp; ++q) wbase += s_red[q];
  if (wi < nwords) {
    // offsets never exceed the capacity: rows beyond it are dropped (PP_STATUS_CAND_OVERFLOW) and the consumers
    // (pp_loss_list) index the list arrays up to offsets[B]
    if (wi % words_per_sweep == 0) pos_offsets[wi / words_per_sweep] = min(wbase, cap);
    if (wi == nwords - 1) pos_offsets[B] = min(wbase + c, cap);
  }
  unsigned todo = __ballot_sync(0xffffffffu, w != 0u);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const unsigned ws = __shfl_sync(0xffffffffu, w, src);
    const int wb = __shfl_sync(0xffffffffu, wbase, src);
    const size_t wsrc = wi - lane + src;
    if ((ws >> lane) & 1u) {
      const int b = (int)(wsrc / words_per_sweep);
      const long long a = (long long)(wsrc - (size_t)b * words_per_sweep) * 32 + lane;
      const int slot = wb + __popc(ws & ((1u << lane) - 1u));
      if (a < A) {
        if (slot < cap) {
          const unsigned fl = anchor_flags(ea, b, A, a);
          float crow[9], rrow[9];
          encode_row(ea, gp, b, A, a, fl, crow, rrow);
          pos_anchor[slot] = (int)((long long)b * A + a);
#pragma unroll
          for (int k = 0; k < 9; ++k) { pos_cls[(size_t)slot * 9 + k] = crow[k]; pos_reg[(size_t)slot * 9 + k] = rrow[k]; }
        } else {
          atomicOr(status, PP_STATUS_CAND_OVERFLOW);
        }
      }
    }
  }
}

struct TargetWs {
  double* cand_iou;          // [Gt, kCandCap] pass-0 -> pass-1 IoU cache
  unsigned long long* best;  // [B, A]   zero-init
  unsigned* posmask;         // [B, ceil(A/32)] zero-init
  unsigned* forcedmask;      // [B, ceil(A/32)] zero-init
  int* arg;                  // [B, A]   0x7f-init
  int* blocksum;             // [ceil(B*ceil(A/32)/256)] positives per block of mask words (list output)
  size_t zero_bytes;
};

template <class AR>
static void targets_layout(AR& a, TargetWs* ws, int B, long long A, long long Gt) {
  auto pc = a.template take<double>((size_t)(Gt > 0 ? Gt : 1) * kCandCap);
  if (ws) ws->cand_iou = pc;
  size_t z0 = a.used;
  auto p0 = a.template take<unsigned long long>((size_t)B * A);
  auto p1 = a.template take<unsigned>((size_t)B * ((A + 31) / 32));
  auto p3 = a.template take<unsigned>((size_t)B * ((A + 31) / 32));
  size_t z1 = a.used;
  auto p2 = a.template take<int>((size_t)B * A);
  auto p4 = a.template take<int>(((size_t)B * ((A + 31) / 32) + 255) / 256 + 1);
  if (ws) { ws->best = p0; ws->posmask = p1; ws->forcedmask = p3; ws->arg = p2; ws->blocksum = p4; ws->zero_bytes = z1 - z0; }
}

struct SizeArena3 {
  size_t used = 0;
  template <class T>
  T* take(size_t count) { used += align_up(count * sizeof(T)); return nullptr; }
};

struct HostIndex {
  IndexHeader hdr;
  std::vector<int> start, ids;
};

static size_t geom_offset(const HostIndex& hi);
static bool build_host_index(const double* c, long long A, HostIndex& hi) {
  if (c == nullptr || A < 1 || A > 0x7fffffffll) return false;
  double x0 = 0, x1 = 0, y0 = 0, y1 = 0;
  bool any = false;
  for (long long i = 0; i < A; ++i) {
    const double x = c[i * 3], y = c[i * 3 + 1];
    if (!(x == x) || !(y == y) || x > 1e300 || x < -1e300 || y > 1e300 || y < -1e300) continue;
    if (!any) { x0 = x1 = x; y0 = y1 = y; any = true; }
    if (x < x0) x0 = x;
    if (x > x1) x1 = x;
    if (y < y0) y0 = y;
    if (y > y1) y1 = y;
  }
  double cell = 4.0;
  const double ext = (x1 - x0) > (y1 - y0) ? (x1 - x0) : (y1 - y0);
  if (ext / cell > 2048.0) cell = ext / 2048.0;
  IndexHeader& h = hi.hdr;
  h.x0 = x0; h.y0 = y0; h.cell = cell;
  h.nbx = (int)floor((x1 - x0) / cell) + 1;
  h.nby = (int)floor((y1 - y0) / cell) + 1;
  h.A = A; h.magic = kIndexMagic; h.geom_off = 0; h.pad = 0;
  const int nb = h.nbx * h.nby;
  hi.start.assign((size_t)nb + 1, 0);
  std::vector<int> bucket((size_t)A);
  for (long long i = 0; i < A; ++i) {
    const double x = c[i * 3], y = c[i * 3 + 1];
    int bx = 0, by = 0;
    if (x == x && y == y && x <= 1e300 && x >= -1e300 && y <= 1e300 && y >= -1e300) {
      bx = (int)floor((x - x0) / cell);
      by = (int)floor((y - y0) / cell);
      if (bx < 0) bx = 0; if (bx >= h.nbx) bx = h.nbx - 1;
      if (by < 0) by = 0; if (by >= h.nby) by = h.nby - 1;
    }
    bucket[(size_t)i] = by * h.nbx + bx;
    hi.start[(size_t)bucket[(size_t)i] + 1]++;
  }
  int mx = 0;
  for (int k = 0; k < nb; ++k) {
    if (hi.start[(size_t)k + 1] > mx) mx = hi.start[(size_t)k + 1];
    hi.start[(size_t)k + 1] += hi.start[(size_t)k];
  }
  h.max_bucket = mx;
  hi.ids.assign((size_t)A, 0);
  std::vector<int> cur(hi.start.begin(), hi.start.end() - 1);
  for (long long i = 0; i < A; ++i) hi.ids[(size_t)cur[(size_t)bucket[(size_t)i]]++] = (int)i;  // ascending ids per bucket
  return true;
}

static size_t geom_offset(const HostIndex& hi) {
  const size_t ints = sizeof(IndexHeader) + (hi.start.size() + hi.ids.size()) * sizeof(int);
  return (ints + 15) / 16 * 16;
}

}  // namespace pp

extern "C" {

int pp_make_ious(const double* d_a_corners, const double* d_g_corners, const double* d_a_centers,
                 const double* d_g_centers, int64_t A, int64_t G, double* d_ious,
                 int32_t* d_status, pp_stream_t stream) {
  using namespace pp;
  if (A < 0 || G < 0 || d_status == nullptr) return PP_ERR_INVALID_ARG;
  if (A == 0 || G == 0) return PP_OK;
  if (!d_a_corners || !d_g_corners || !d_a_centers || !d_g_centers || !d_ious) return PP_ERR_INVALID_ARG;
  const long long total = (long long)A * G;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 64;
  if (blocks > cap) blocks = cap;
  PP_KERNEL("k_make_ious", (cudaStream_t)stream, k_make_ious<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(d_a_corners, d_g_corners, d_a_centers,
                                                            d_g_centers, A, G, d_ious, d_status));
  return PP_OK;
}

size_t pp_anchor_index_bytes(const double* h_a_centers, int64_t A, int32_t with_geometry) {
  pp::HostIndex hi;
  if (!pp::build_host_index(h_a_centers, A, hi)) return 0;
  return pp::geom_offset(hi) + (with_geometry ? (size_t)A * 10 * sizeof(double) : 0) + pp::kAlign;
}

int pp_anchor_index_build(const double* h_a_centers, const double* h_a_corners, int64_t A, void* d_index,
                          size_t index_bytes, pp_stream_t stream) {
  using namespace pp;
  HostIndex hi;
  if (d_index == nullptr || !build_host_index(h_a_centers, A, hi)) return PP_ERR_INVALID_ARG;
  const size_t goff = geom_offset(hi);
  const size_t need = goff + (h_a_corners != nullptr ? (size_t)A * 10 * sizeof(double) : 0);
  if (index_bytes < need) return PP_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* d = (unsigned char*)d_index;
  std::vector<double> geom;
  if (h_a_corners != nullptr) {
    hi.hdr.geom_off = (long long)goff;
    geom.resize((size_t)A * 10);
    for (long long k = 0; k < A; ++k) {
      const long long a = hi.ids[(size_t)k];
      geom[(size_t)k * 10 + 0] = h_a_centers[a * 3];
      geom[(size_t)k * 10 + 1] = h_a_centers[a * 3 + 1];
      for (int q = 0; q < 8; ++q) geom[(size_t)k * 10 + 2 + q] = h_a_corners[a * 8 + q];
    }
    PP_CUDA(cudaMemcpyAsync(d + goff, geom.data(), geom.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  PP_CUDA(cudaMemcpyAsync(d, &hi.hdr, sizeof(IndexHeader), cudaMemcpyHostToDevice, st));
  PP_CUDA(cudaMemcpyAsync(d + sizeof(IndexHeader), hi.start.data(), hi.start.size() * sizeof(int),
                          cudaMemcpyHostToDevice, st));
  PP_CUDA(cudaMemcpyAsync(d + sizeof(IndexHeader) + hi.start.size() * sizeof(int), hi.ids.data(),
                          hi.ids.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  PP_CUDA(cudaStreamSynchronize(st));  // host vectors die at return; one-off call
  return PP_OK;
}

size_t pp_assign_targets_workspace_bytes(int32_t n_sweeps, int64_t A, int64_t total_gt,
                                         const void* h_index_header) {
  (void)h_index_header;
  if (n_sweeps < 1 || n_sweeps > PP_MAX_SWEEPS || A < 1) return 0;
  pp::SizeArena3 a;
  pp::targets_layout(a, (pp::TargetWs*)nullptr, n_sweeps, A, total_gt);
  return a.used + pp::kAlign;
}

}  // extern "C"

namespace pp {
struct PosListOut {
  int* anchor;
  float* cls;
  float* reg;
  int* offsets;
  int cap;
};

static int assign_impl(const double* d_a_corners, const double* d_a_centers, const double* d_a_wlh,
                       const double* d_a_yaw, const void* d_anchor_index, int64_t A,
                       const double* d_g_corners, const double* d_g_centers, const double* d_g_wlh,
                       const double* d_g_yaw, const int32_t* d_g_cls, const int64_t* h_gt_offsets,
                       int32_t n_sweeps, int32_t num_classes, double pos_thresh, float* d_cls,
                       float* d_reg, const PosListOut* pl, int32_t* d_top_anchor, int32_t* d_counts, int32_t* d_status,
                       void* d_workspace, size_t workspace_bytes, pp_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const bool dense = d_cls != nullptr || d_reg != nullptr;
  if (!d_a_corners || !d_a_centers || !d_a_wlh || !d_a_yaw || !d_anchor_index || A < 1 ||
      A > 0x7fffffffll || !h_gt_offsets || n_sweeps < 1 || n_sweeps > PP_MAX_SWEEPS ||
      num_classes < 1 || (dense && (!d_cls || !d_reg)) || (!dense && !pl) || !d_counts || !d_status ||
      (long long)A * (num_classes > 9 ? num_classes : 9) > 0x7fffffffll)
    return PP_ERR_INVALID_ARG;
  if (pl && (!pl->anchor || !pl->cls || !pl->reg || !pl->offsets || pl->cap < 1 || num_classes != 9 ||
             (long long)A * n_sweeps > 0x7fffffffll))
    return PP_ERR_INVALID_ARG;
  GtParams gp;
  gp.n_sweeps = n_sweeps;
  for (int s = 0; s <= n_sweeps; ++s) {
    gp.off[s] = h_gt_offsets[s];
    if (s > 0 && h_gt_offsets[s] < h_gt_offsets[s - 1]) return PP_ERR_INVALID_ARG;
  }
  if (gp.off[0] != 0) return PP_ERR_INVALID_ARG;
  const long long Gt = gp.off[n_sweeps];
  if (Gt > 0 && (!d_g_corners || !d_g_centers || !d_g_wlh || !d_g_yaw || !d_g_cls || !d_top_anchor))
    return PP_ERR_INVALID_ARG;
  if (Gt > 0x7fffffffll) return PP_ERR_INVALID_ARG;
  Arena arena(d_workspace, workspace_bytes);
  TargetWs ws{};
  targets_layout(arena, &ws, n_sweeps, A, Gt);
  if (!arena.ok) return PP_ERR_WORKSPACE;

  PP_CUDA(cudaMemsetAsync(ws.best, 0, ws.zero_bytes, st));
  PP_CUDA(cudaMemsetAsync(ws.arg, 0x7f, (size_t)n_sweeps * A * sizeof(int), st));
  PP_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)n_sweeps * 4 * sizeof(int), st));
  const unsigned char* idx = (const unsigned char*)d_anchor_index;
  if (Gt > 0) {
    PP_KERNEL("k_iou_pass0", st, k_iou_pass<0><<<(int)Gt, 256, 0, st>>>(d_a_corners, d_a_centers, idx, A, d_g_corners,
                                           d_g_centers, gp, pos_thresh, ws.best, ws.arg, ws.posmask,
                                           ws.forcedmask, d_top_anchor, d_counts, ws.cand_iou, d_status));
    PP_KERNEL("k_iou_pass1", st, k_iou_pass<1><<<(int)Gt, 256, 0, st>>>(d_a_corners, d_a_centers, idx, A, d_g_corners,
                                           d_g_centers, gp, pos_thresh, ws.best, ws.arg, ws.posmask,
                                           ws.forcedmask, d_top_anchor, d_counts, ws.cand_iou, d_status));
  }
  EncodeArgs ea{d_a_centers, d_a_wlh, d_a_yaw, d_g_centers, d_g_wlh, d_g_yaw, d_g_cls, ws.arg,
                ws.posmask, ws.forcedmask, d_top_anchor};
  if (pl) {
    const size_t nwords = (size_t)((A + 31) / 32) * n_sweeps;
    const int nblk = (int)((nwords + 255) / 256);
    PP_KERNEL("k_pos_count", st, (k_pos_count<<<nblk, 256, 0, st>>>(ws.posmask, ws.forcedmask, nwords, ws.blocksum)));
    PP_KERNEL("k_pos_emit", st,
              (k_pos_emit<<<nblk, 256, 0, st>>>(ea, gp, A, n_sweeps, ws.blocksum, pl->anchor, pl->cls, pl->reg, pl->offsets,
                                                pl->cap, d_status)));
    if (!dense) return PP_OK;
  }
  const bool both = num_classes == 9 && ((long long)A * 9 % 4 == 0) && ((uintptr_t)d_cls % 16 == 0) &&
                    ((uintptr_t)d_reg % 16 == 0);
  if (both) {
    const size_t n4 = (size_t)n_sweeps * A * 9 / 4;
    if (g_opt_encode_bulk) {
      PP_KERNEL("k_encode_zero", st,
                (k_encode_zero_bulk<<<sm_count() * 4, 128, 0, st>>>((unsigned char*)d_cls, (unsigned char*)d_reg, n4 * 16)));
    } else {
      PP_KERNEL("k_encode_zero", st,
                (k_encode_zero<<<sm_count() * 8, 256, 0, st>>>((float4*)d_cls, (float4*)d_reg, n4)));
    }
    const size_t nwords = (size_t)((A + 31) / 32) * n_sweeps;
    PP_KERNEL("k_encode_patch", st,
              (k_encode_patch<<<(int)((nwords + 127) / 128), 128, 0, st>>>(ea, gp, A, n_sweeps, d_cls, d_reg)));
    return PP_OK;
  }
  {
    const long long total = (long long)A * num_classes;
    const bool vec_ok = (total % 4 == 0) && ((uintptr_t)d_cls % 16 == 0);
    long long blocks = (A + kEncTile - 1) / kEncTile;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)n_sweeps);
    if (num_classes == 9) {
      PP_KERNEL("k_encode", st, (k_encode<false, 9><<<grid, 256, 0, st>>>(ea, gp, A, 9, vec_ok, d_cls)));
    } else {
      PP_KERNEL("k_encode", st, (k_encode<false, 0><<<grid, 256, 0, st>>>(ea, gp, A, num_classes, vec_ok, d_cls)));
    }
  }
  {
    const long long total = (long long)A * 9;
    const bool vec_ok = (total % 4 == 0) && ((uintptr_t)d_reg % 16 == 0);
    long long blocks = (A + kEncTile - 1) / kEncTile;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)n_sweeps);
    PP_KERNEL("k_encode", st, (k_encode<true, 9><<<grid, 256, 0, st>>>(ea, gp, A, 9, vec_ok, d_reg)));
  }
  return PP_OK;
}
}  // namespace pp

extern "C" {

int pp_assign_targets(const double* d_a_corners, const double* d_a_centers, const double* d_a_wlh,
                      const double* d_a_yaw, const void* d_anchor_index, int64_t A,
                      const double* d_g_corners, const double* d_g_centers, const double* d_g_wlh,
                      const double* d_g_yaw, const int32_t* d_g_cls, const int64_t* h_gt_offsets,
                      int32_t n_sweeps, int32_t num_classes, double pos_thresh, float* d_cls,
                      float* d_reg, int32_t* d_top_anchor, int32_t* d_counts, int32_t* d_status,
                      void* d_workspace, size_t workspace_bytes, pp_stream_t stream) {
  if (!d_cls || !d_reg) return PP_ERR_INVALID_ARG;
  return pp::assign_impl(d_a_corners, d_a_centers, d_a_wlh, d_a_yaw, d_anchor_index, A, d_g_corners, d_g_centers, d_g_wlh,
                         d_g_yaw, d_g_cls, h_gt_offsets, n_sweeps, num_classes, pos_thresh, d_cls, d_reg, nullptr,
                         d_top_anchor, d_counts, d_status, d_workspace, workspace_bytes, stream);
}

int pp_assign_targets_list(const double* d_a_corners, const double* d_a_centers, const double* d_a_wlh,
                           const double* d_a_yaw, const void* d_anchor_index, int64_t A,
                           const double* d_g_corners, const double* d_g_centers, const double* d_g_wlh,
                           const double* d_g_yaw, const int32_t* d_g_cls, const int64_t* h_gt_offsets,
                           int32_t n_sweeps, int32_t num_classes, double pos_thresh, int32_t* d_pos_anchor,
                           float* d_pos_cls, float* d_pos_reg, int32_t* d_pos_offsets, int32_t capacity,
                           float* d_cls, float* d_reg, int32_t* d_top_anchor, int32_t* d_counts, int32_t* d_status,
                           void* d_workspace, size_t workspace_bytes, pp_stream_t stream) {
  pp::PosListOut pl{d_pos_anchor, d_pos_cls, d_pos_reg, d_pos_offsets, capacity};
  return pp::assign_impl(d_a_corners, d_a_centers, d_a_wlh, d_a_yaw, d_anchor_index, A, d_g_corners, d_g_centers, d_g_wlh,
                         d_g_yaw, d_g_cls, h_gt_offsets, n_sweeps, num_classes, pos_thresh, d_cls, d_reg, &pl,
                         d_top_anchor, d_counts, d_status, d_workspace, workspace_bytes, stream);
}

}  // extern "C"
