// pillarize.cu -- K1: point -> pillar binning, order-exact compaction, decoration and per-slot
// mean subtraction for sm_100a.  Replaces data/pillars.cpp:236-398 (create_pillars) and the
// tensor glue of data/dataset.py:99-106 of the reference.
//
// Order-exactness: pillars are numbered by the input index of their first
// in-range point; inside a pillar points keep input order; the first N are emitted; the mean is
// the reference's sequential running mean over ALL in-range points of the pillar.  atomics only
// ever decide things that do not depend on order (min index, counts, list placement before the
// rank pass), so the result is deterministic and identical to the sequential CPU algorithm.
//
// Stages (one launch each, all sweeps of the batch in the same launch):
//   k_bin        per point: range filter + floor binning in fp64, atomicMin(first index of cell),
//                atomicAdd(count of cell)
//   k_tilecount  per 1024-point tile: number of first-touch points
//   k_assign     per tile: exclusive scan of first-touch flags -> pillar slot of each cell,
//                list segment of each kept pillar
//   k_scatter    per point: append its index to its pillar's segment (unordered)
//   k_rank       per point: rank = number of smaller indices in the segment; the point's running-mean
//                terms (four IEEE divisions) go to its place in the ordered segment
//   k_rank_big   pillars with more than kBig points: ordered compaction by a block scan instead
//   k_mean       per pillar (one thread): sequential multiply-add chain of the running mean, indices row
//   k_emit_*     dense [B,9,P,N] float with fused "- data_mean", or compact fp64 rows
#include <type_traits>

#include "internal.cuh"

namespace pp {

constexpr int kTile = 1024;  // points per scan tile == threads per block of the scan kernels
constexpr int kBig = 1024;   // pillars with more points than this take the block-scan rank path

struct GridDev {
  double x_step, y_step, x_min, y_min, z_min, x_max, y_max, z_max, canvas_height;
  int nx, ny, ncell;
};

__device__ __forceinline__ int find_sweep(const SweepParams& sw, long long i) {
  int lo = 0, hi = sw.n_sweeps - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (sw.off[mid] <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ int find_tile_sweep(const SweepParams& sw, int t) {
  int lo = 0, hi = sw.n_sweeps - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (sw.tile_start[mid] <= t) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <typename T>
__device__ __forceinline__ void load_xyz(const T* __restrict__ pts, long long i, long long sp,
                                         long long sc, bool vec4, double& x, double& y, double& z,
                                         double& r) {
  if (sizeof(T) == 4 && vec4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(pts) + i);
    x = v.x; y = v.y; z = v.z; r = v.w;
  } else {
    const T* p = pts + i * sp;
    x = (double)__ldg(p);
    y = (double)__ldg(p + sc);
    z = (double)__ldg(p + 2 * sc);
    r = (double)__ldg(p + 3 * sc);
  }
}

// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_bin(const T* __restrict__ pts, long long sp, long long sc,
                                             bool vec4, SweepParams sw, GridDev g,
                                             int* __restrict__ cell_of_point,
                                             int* __restrict__ cell_first,
                                             int* __restrict__ cell_count,
                                             int* __restrict__ status) {
  const long long total = sw.off[sw.n_sweeps];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = find_sweep(sw, i);
    double x, y, z, r;
    load_xyz(pts, i, sp, sc, vec4, x, y, z, r);
    int cell = -1;
    // data/pillars.cpp:271-275 (half-open box; written as the reference writes it)
    if (!((x >= g.x_max) || (x < g.x_min) || (y >= g.y_max) || (y < g.y_min) ||
          (z >= g.z_max) || (z < g.z_min))) {
      // data/pillars.cpp:278-279, IEEE double subtract / divide / floor
      const double fx = floor(__ddiv_rn(__dsub_rn(x, g.x_min), g.x_step));
      const double fy = floor(__ddiv_rn(__dsub_rn(y, g.y_min), g.y_step));
      if (fx >= 0.0 && fx < (double)g.nx && fy >= 0.0 && fy < (double)g.ny) {
        cell = (int)fy * g.nx + (int)fx;
      } else {
        atomicOr(status, PP_STATUS_BAD_POINT);  // NaN/Inf: out of contract, dropped
      }
    }
    cell_of_point[i] = cell;
    if (cell >= 0) {
      const int il = (int)(i - sw.off[b]);
      atomicMin(&cell_first[(size_t)b * g.ncell + cell], il);
      atomicAdd(&cell_count[(size_t)b * g.ncell + cell], 1);
    }
  }
}

__device__ __forceinline__ bool first_touch_flag(const SweepParams& sw, const GridDev& g, int t,
                                                 int& b, int& il, int& cell,
                                                 const int* __restrict__ cell_of_point,
                                                 const int* __restrict__ cell_first) {
  b = find_tile_sweep(sw, t);
  il = (t - sw.tile_start[b]) * kTile + (int)threadIdx.x;
  const long long n_b = sw.off[b + 1] - sw.off[b];
  cell = -1;
  if (il < n_b) cell = cell_of_point[sw.off[b] + il];
  return cell >= 0 && cell_first[(size_t)b * g.ncell + cell] == il;
}

__global__ void __launch_bounds__(kTile) k_tilecount(SweepParams sw, GridDev g,
                                                     const int* __restrict__ cell_of_point,
                                                     const int* __restrict__ cell_first,
                                                     int* __restrict__ tile_count) {
  __shared__ int warp_cnt[kTile / 32];
  int b, il, cell;
  const bool flag = first_touch_flag(sw, g, blockIdx.x, b, il, cell, cell_of_point, cell_first);
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (lane_id() == 0) warp_cnt[threadIdx.x >> 5] = __popc(m);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = warp_cnt[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) tile_count[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(kTile) k_assign(
    SweepParams sw, GridDev g, int P, const int* __restrict__ cell_of_point,
    const int* __restrict__ cell_first, const int* __restrict__ cell_count,
    const int* __restrict__ tile_count, int* __restrict__ cell_slot, int* __restrict__ pil_cnt,
    int* __restrict__ pil_off, int* __restrict__ pil_cell, int* __restrict__ list_cursor,
    int* __restrict__ big_count, int* __restrict__ big_list, int* __restrict__ num_pillars) {
  __shared__ int warp_sum[kTile / 32];
  __shared__ int s_base;
  const int t = blockIdx.x;
  int b, il, cell;
  const bool flag = first_touch_flag(sw, g, t, b, il, cell, cell_of_point, cell_first);

  // base = number of first-touch points in the earlier tiles of this sweep
  int part = 0;
  for (int k = sw.tile_start[b] + (int)threadIdx.x; k < t; k += kTile) part += tile_count[k];
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane_id() == 0) warp_sum[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = warp_sum[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) s_base = v;
  }
  __syncthreads();
  const int base = s_base;

  // exclusive scan of the flags inside the tile
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  const int in_warp = __popc(m & ((1u << lane_id()) - 1u));
  __syncthreads();  // warp_sum reuse
  if (lane_id() == 0) warp_sum[threadIdx.x >> 5] = __popc(m);
  __syncthreads();
  int before = 0, total = 0;
  for (int w = 0; w < kTile / 32; ++w) {
    const int v = warp_sum[w];
    if (w < (int)(threadIdx.x >> 5)) before += v;
    total += v;
  }
  // kept pillars reserve a list segment of `cnt` entries: block-scan the counts and take ONE
  // atomicAdd per tile (segments only have to be disjoint, not ordered)
  int slot = -1, cnt = 0;
  size_t ci = 0;
  if (flag) {
    slot = base + before + in_warp;
    ci = (size_t)b * g.ncell + cell;
    if (slot < P) cnt = cell_count[ci];
  }
  int incl = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)lane_id() >= o) incl += v;
  }
  __shared__ int warp_cnt[kTile / 32];
  __shared__ int s_seg_base;
  if (lane_id() == 31) warp_cnt[threadIdx.x >> 5] = incl;
  __syncthreads();
  int cnt_before = 0, cnt_total = 0;
  for (int w = 0; w < kTile / 32; ++w) {
    const int v = warp_cnt[w];
    if (w < (int)(threadIdx.x >> 5)) cnt_before += v;
    cnt_total += v;
  }
  if (threadIdx.x == 0) s_seg_base = cnt_total > 0 ? atomicAdd(&list_cursor[b], cnt_total) : 0;
  __syncthreads();
  if (flag) {
    if (slot < P) {
      const int off = s_seg_base + cnt_before + incl - cnt;
      const size_t pi = (size_t)b * P + slot;
      cell_slot[ci] = slot;
      pil_cnt[pi] = cnt;
      pil_off[pi] = off;
      pil_cell[pi] = cell;
      if (cnt > kBig) big_list[atomicAdd(big_count, 1)] = (int)pi;
    } else {
      cell_slot[ci] = -1;  // pillar beyond the max_pillars cap: data/pillars.cpp:339
    }
  }
  const int ntiles_b = sw.tile_start[b + 1] - sw.tile_start[b];
  if (threadIdx.x == 0 && t - sw.tile_start[b] == ntiles_b - 1) {
    const int np = base + total;
    num_pillars[b] = np < P ? np : P;
  }
}

__global__ void __launch_bounds__(256) k_scatter(SweepParams sw, GridDev g, int P,
                                                 const int* __restrict__ cell_of_point,
                                                 const int* __restrict__ cell_slot,
                                                 const int* __restrict__ pil_off,
                                                 int* __restrict__ pil_cursor,
                                                 int* __restrict__ list_u) {
  const long long total = sw.off[sw.n_sweeps];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cell = cell_of_point[i];
    if (cell < 0) continue;
    const int b = find_sweep(sw, i);
    const int slot = cell_slot[(size_t)b * g.ncell + cell];
    if (slot < 0) continue;
    const size_t pi = (size_t)b * P + slot;
    const int pos = atomicAdd(&pil_cursor[pi], 1);
    list_u[sw.off[b] + pil_off[pi] + pos] = (int)(i - sw.off[b]);
  }
}

// Per-point terms of the reference's running mean (data/pillars.cpp:311-328)
//   m <- m*(n/(n+1)) + x/(n+1)
// for the point of rank n: {a = n/(n+1), x/(n+1), y/(n+1), z/(n+1)} (the first point initialises the
// mean: a = 0, terms = x, y, z).  The four IEEE divisions do not depend on m, so they are done here,
// one point per thread with every lane busy, and k_mean is left with the bare multiply-add chain.
template <typename T>
__device__ __forceinline__ void mean_terms(const T* __restrict__ pts, long long gi, long long sp, long long sc,
                                           bool vec4, int rank, double4* __restrict__ out) {
  double x, y, z, r;
  load_xyz(pts, gi, sp, sc, vec4, x, y, z, r);
  double4 t;
  if (rank == 0) {
    t = make_double4(0.0, x, y, z);
  } else {
    const double n1 = __dadd_rn((double)rank, 1.0);
    t = make_double4(__ddiv_rn((double)rank, n1), __ddiv_rn(x, n1), __ddiv_rn(y, n1), __ddiv_rn(z, n1));
  }
  *out = t;
}

template <typename T>
__global__ void __launch_bounds__(256) k_rank(const T* __restrict__ pts, long long sp, long long sc, bool vec4,
                                              SweepParams sw, GridDev g, int P,
                                              const int* __restrict__ cell_of_point,
                                              const int* __restrict__ cell_slot,
                                              const int* __restrict__ pil_cnt,
                                              const int* __restrict__ pil_off,
                                              const int* __restrict__ list_u,
                                              double4* __restrict__ terms,
                                              int* __restrict__ rank_of_point) {
  const long long total = sw.off[sw.n_sweeps];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cell = cell_of_point[i];
    int rank = -1;
    if (cell >= 0) {
      const int b = find_sweep(sw, i);
      const int slot = cell_slot[(size_t)b * g.ncell + cell];
      if (slot >= 0) {
        const size_t pi = (size_t)b * P + slot;
        const int c = pil_cnt[pi];
        if (c <= kBig) {
          const int il = (int)(i - sw.off[b]);
          const int* seg = list_u + sw.off[b] + pil_off[pi];
          rank = 0;
          for (int k = 0; k < c; ++k) rank += (seg[k] < il) ? 1 : 0;
          mean_terms(pts, i, sp, sc, vec4, rank, terms + sw.off[b] + pil_off[pi] + rank);
        } else {
          rank = -2;  // filled by k_rank_big
        }
      }
    }
    if (rank != -2) rank_of_point[i] = rank;
  }
}

// One block per big pillar: ordered stream compaction of the sweep's points that fall in its cell.
template <typename T>
__global__ void __launch_bounds__(kTile) k_rank_big(const T* __restrict__ pts, long long sp, long long sc, bool vec4,
                                                    SweepParams sw, int P,
                                                    const int* __restrict__ cell_of_point,
                                                    const int* __restrict__ pil_off,
                                                    const int* __restrict__ pil_cell,
                                                    const int* __restrict__ big_count,
                                                    const int* __restrict__ big_list,
                                                    double4* __restrict__ terms,
                                                    int* __restrict__ rank_of_point) {
  __shared__ int warp_sum[kTile / 32];
  const int nbig = *big_count;
  for (int k = blockIdx.x; k < nbig; k += gridDim.x) {
    const int pi = big_list[k];
    const int b = pi / P;
    const int cell = pil_cell[pi];
    const long long n_b = sw.off[b + 1] - sw.off[b];
    double4* out = terms + sw.off[b] + pil_off[pi];
    int running = 0;
    for (long long s = 0; s < n_b; s += kTile) {
      const long long il = s + threadIdx.x;
      const bool flag = il < n_b && cell_of_point[sw.off[b] + il] == cell;
      const unsigned m = __ballot_sync(0xffffffffu, flag);
      const int in_warp = __popc(m & ((1u << lane_id()) - 1u));
      __syncthreads();
      if (lane_id() == 0) warp_sum[threadIdx.x >> 5] = __popc(m);
      __syncthreads();
      int before = 0, total = 0;
      for (int w = 0; w < kTile / 32; ++w) {
        const int v = warp_sum[w];
        if (w < (int)(threadIdx.x >> 5)) before += v;
        total += v;
      }
      if (flag) {
        const int rank = running + before + in_warp;
        mean_terms(pts, sw.off[b] + il, sp, sc, vec4, rank, out + rank);
        rank_of_point[sw.off[b] + il] = rank;
      }
      running += total;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ double4 ld_terms(const double4* __restrict__ p) {
  const double2* q = reinterpret_cast<const double2*>(p);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

// One thread per pillar slot: the sequential multiply-add chain of the reference's running mean over
// the terms k_rank prepared (32-byte records, contiguous per pillar, L2-resident), evaluated in input
// order with separately rounded IEEE operations.  The records of the next steps are loaded ahead of
// the chain (4-deep), so a pillar costs ~2 dependent fp64 operations per point (median pillar: 2
// points, longest of a Lyft-shaped sweep: ~300).
__global__ void __launch_bounds__(128) k_mean(SweepParams sw, GridDev g, int P,
                                              const int* __restrict__ num_pillars,
                                              const int* __restrict__ pil_cnt,
                                              const int* __restrict__ pil_off,
                                              const int* __restrict__ pil_cell,
                                              const double4* __restrict__ terms,
                                              double* __restrict__ pil_mean,
                                              long long* __restrict__ indices) {
  const long long grp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (grp >= (long long)sw.n_sweeps * P) return;
  const int b = (int)(grp / P);
  const int slot = (int)(grp - (long long)b * P);
  if (slot >= num_pillars[b]) {
    if (indices != nullptr) { indices[grp * 3 + 0] = 0; indices[grp * 3 + 1] = 0; indices[grp * 3 + 2] = 0; }
    return;
  }
  const int c = pil_cnt[grp];
  const double4* seg = terms + sw.off[b] + pil_off[grp];
  double4 t = ld_terms(seg);                          // rank 0: a = 0, terms = the point itself
  double m0 = t.y, m1 = t.z, m2 = t.w;             // data/pillars.cpp:313-317
  constexpr int kAhead = 4;
  int k = 1;
  if (k + kAhead <= c) {
    double4 u[kAhead], v[kAhead];
#pragma unroll
    for (int j = 0; j < kAhead; ++j) u[j] = ld_terms(seg + k + j);
    for (; k + kAhead <= c; k += kAhead) {
      const bool more = k + 2 * kAhead <= c;
      // long pillars are bound by the L2 round trip of their record stream, not by the arithmetic
      // chain: pull the 128-byte line eight batches ahead into L1
      if (k + 9 * kAhead <= c) asm volatile("prefetch.global.L1 [%0];" ::"l"(seg + k + 8 * kAhead));
      if (more) {
#pragma unroll
        for (int j = 0; j < kAhead; ++j) v[j] = ld_terms(seg + k + kAhead + j);   // in flight during the chain below
      }
#pragma unroll
      for (int j = 0; j < kAhead; ++j) {
        m0 = __dadd_rn(__dmul_rn(m0, u[j].x), u[j].y);   // data/pillars.cpp:324-326
        m1 = __dadd_rn(__dmul_rn(m1, u[j].x), u[j].z);
        m2 = __dadd_rn(__dmul_rn(m2, u[j].x), u[j].w);
      }
      if (more) {
#pragma unroll
        for (int j = 0; j < kAhead; ++j) u[j] = v[j];
      }
    }
  }
  for (; k < c; ++k) {
    const double4 u = ld_terms(seg + k);
    m0 = __dadd_rn(__dmul_rn(m0, u.x), u.y);
    m1 = __dadd_rn(__dmul_rn(m1, u.x), u.z);
    m2 = __dadd_rn(__dmul_rn(m2, u.x), u.w);
  }
  pil_mean[grp * 3 + 0] = m0;
  pil_mean[grp * 3 + 1] = m1;
  pil_mean[grp * 3 + 2] = m2;
  if (indices != nullptr) {
    const int cell = pil_cell[grp];
    const double cx = (double)(cell % g.nx);
    const double cy = __dsub_rn(__dsub_rn(g.canvas_height, 1.0), (double)(cell / g.nx));
    indices[grp * 3 + 0] = 1;                 // data/pillars.cpp:390-392, then .long()
    indices[grp * 3 + 1] = (long long)cx;
    indices[grp * 3 + 2] = (long long)cy;
  }
}

// Features of one point: data/pillars.cpp:30-31,48-56,381-383 in fp64.
__device__ __forceinline__ void point_features(double x, double y, double z, double r, double cx,
                                               double cy, const double* __restrict__ mean,
                                               double f[9]) {
  f[0] = x; f[1] = y; f[2] = z; f[3] = r;
  f[4] = __dsub_rn(cx, x);
  f[5] = __dsub_rn(cy, y);
  f[6] = __dsub_rn(mean[0], x);
  f[7] = __dsub_rn(mean[1], y);
  f[8] = __dsub_rn(mean[2], z);
}

constexpr int kFeatGroup = 3;   // features per thread in k_emit_dense (9 = 3 groups -> grid.y)

// Per kept point (rank < N inside a kept pillar): the nine decorated features in fp64, rounded once
// to fp32 (torch .float(), data/dataset.py:101), stored compactly at the point's position in its
// pillar's ordered segment.  Keeps all fp64 math out of the streaming kernel below.
template <typename T>
__global__ void __launch_bounds__(256) k_feat(const T* __restrict__ pts, long long sp, long long sc,
                                              bool vec4, SweepParams sw, GridDev g, int P, int N,
                                              const int* __restrict__ cell_of_point,
                                              const int* __restrict__ cell_slot,
                                              const int* __restrict__ rank_of_point,
                                              const int* __restrict__ pil_off,
                                              const double* __restrict__ pil_mean,
                                              float* __restrict__ feat_c) {
  const long long total = sw.off[sw.n_sweeps];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cell = cell_of_point[i];
    if (cell < 0) continue;
    const int b = find_sweep(sw, i);
    const int slot = cell_slot[(size_t)b * g.ncell + cell];
    if (slot < 0) continue;
    const int rank = rank_of_point[i];
    if (rank >= N) continue;                       // data/pillars.cpp:371 first-N cap
    const size_t pi = (size_t)b * P + slot;
    double x, y, z, r, ft[9];
    load_xyz(pts, i, sp, sc, vec4, x, y, z, r);
    const double cx = (double)(cell % g.nx);
    const double cy = __dsub_rn(__dsub_rn(g.canvas_height, 1.0), (double)(cell / g.nx));
    point_features(x, y, z, r, cx, cy, pil_mean + pi * 3, ft);
    float* o = feat_c + (size_t)(sw.off[b] + pil_off[pi] + rank) * 9;
#pragma unroll
    for (int d = 0; d < 9; ++d) o[d] = (float)ft[d];
  }
}

// Dense emit: x[b,d,p,n] = float(feature) - data_mean[d,p,n] for every slot (data/dataset.py:99-105).
// One thread owns VEC consecutive n of one pillar for kFeatGroup features and all sweeps, so that
// data_mean is read once per batch and every access is a full-width coalesced vector.  ~98.7 % of
// the groups hold no point and stream "0 - mean"; occupied ones pick their features up from k_feat.
template <int VEC>
__global__ void __launch_bounds__(256) k_emit_dense(
    SweepParams sw, int P, int N, const float* __restrict__ data_mean, float* __restrict__ xout,
    const int* __restrict__ num_pillars, const int* __restrict__ pil_cnt,
    const int* __restrict__ pil_off, const float* __restrict__ feat_c) {
  typedef typename std::conditional<VEC == 4, float4, float>::type vec_t;
  const long long PN = (long long)P * N;
  const unsigned groups = (unsigned)(PN / VEC);
  const int d0 = blockIdx.y * kFeatGroup;
  for (unsigned gi = blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += gridDim.x * blockDim.x) {
    const unsigned e = gi * VEC;
    const int p = (int)(e / (unsigned)N);
    const int n0 = (int)(e - (unsigned)p * (unsigned)N);
    vec_t m[kFeatGroup];
#pragma unroll
    for (int d = 0; d < kFeatGroup; ++d) {
      if (data_mean != nullptr) {
        m[d] = __ldg(reinterpret_cast<const vec_t*>(data_mean + (d0 + d) * PN + e));
      } else {
        float* mf = reinterpret_cast<float*>(&m[d]);
#pragma unroll
        for (int k = 0; k < VEC; ++k) mf[k] = 0.f;
      }
    }
    for (int b = 0; b < sw.n_sweeps; ++b) {
      const size_t pi = (size_t)b * P + p;
      int c = 0;
      if (p < num_pillars[b]) c = min(pil_cnt[pi], N);  // data/pillars.cpp:371 first-N cap
      float* ob = xout + (size_t)b * 9 * PN + e;
      const float* fc = (n0 < c) ? feat_c + (size_t)(sw.off[b] + pil_off[pi] + n0) * 9 + d0 : nullptr;
#pragma unroll
      for (int d = 0; d < kFeatGroup; ++d) {
        vec_t v;
        float* vf = reinterpret_cast<float*>(&v);
        const float* mf = reinterpret_cast<const float*>(&m[d]);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float f = (fc != nullptr && n0 + k < c) ? fc[k * 9 + d] : 0.f;
          vf[k] = __fsub_rn(f, mf[k]);            // feature (or 0) - mean, like the reference
        }
        __stcs(reinterpret_cast<vec_t*>(ob + (d0 + d) * PN), v);
      }
    }
  }
}

// Compact emit for the numpy-signature drop-in: fp64 rows of the kept points only.
template <typename T>
__global__ void __launch_bounds__(256) k_emit_compact(
    const T* __restrict__ pts, long long sp, long long sc, bool vec4, SweepParams sw, GridDev g,
    int P, int N, const int* __restrict__ cell_of_point, const int* __restrict__ cell_slot,
    const int* __restrict__ rank_of_point, const int* __restrict__ pil_off,
    const double* __restrict__ pil_mean, double* __restrict__ rows, int* __restrict__ slot_out,
    int* __restrict__ n_rows) {
  const long long total = sw.off[1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cell = cell_of_point[i];
    if (cell < 0) continue;
    const int slot = cell_slot[cell];
    if (slot < 0) continue;
    const int rank = rank_of_point[i];
    if (rank >= N) continue;
    const long long q = (long long)pil_off[slot] + rank;
    double x, y, z, r, ft[9];
    load_xyz(pts, i, sp, sc, vec4, x, y, z, r);
    const double cx = (double)(cell % g.nx);
    const double cy = __dsub_rn(__dsub_rn(g.canvas_height, 1.0), (double)(cell / g.nx));
    point_features(x, y, z, r, cx, cy, pil_mean + (size_t)slot * 3, ft);
#pragma unroll
    for (int d = 0; d < 9; ++d) rows[q * 9 + d] = ft[d];
    slot_out[q] = slot * N + rank;
    atomicAdd(n_rows, 1);
  }
}

__global__ void k_pillar_xy(GridDev g, int P, const int* __restrict__ num_pillars,
                            const int* __restrict__ pil_cell, int* __restrict__ pillar_xy,
                            int* __restrict__ counts) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = num_pillars[0];
  if (p == 0) counts[0] = np;
  if (p >= P) return;
  if (p < np) {
    const int cell = pil_cell[p];
    pillar_xy[2 * p] = cell % g.nx;
    pillar_xy[2 * p + 1] = (int)__dsub_rn(__dsub_rn(g.canvas_height, 1.0), (double)(cell / g.nx));
  } else {
    pillar_xy[2 * p] = 0;
    pillar_xy[2 * p + 1] = 0;
  }
}

// ------------------------------------------------------------------------------------------
struct PillarWs {
  int *cell_first, *cell_slot, *cell_of_point, *rank_of_point, *tile_count, *list_u;
  double4* terms;   // per point, in its pillar's ordered segment: {n/(n+1), x/(n+1), y/(n+1), z/(n+1)}
  int *pil_cnt, *pil_off, *pil_cell, *big_list, *num_pillars_scratch;
  double* pil_mean;
  float* feat_c;
  // zero-initialised block (one memset)
  int* zero_begin;
  int *cell_count, *pil_cursor, *list_cursor, *big_count, *n_rows;
  size_t zero_bytes;
};

static bool make_grid(const pp_grid* grid, GridDev& g) {
  if (grid == nullptr) return false;
  if (!(grid->x_step > 0) || !(grid->y_step > 0)) return false;
  if (!(grid->x_max > grid->x_min) || !(grid->y_max > grid->y_min)) return false;
  g.x_step = grid->x_step; g.y_step = grid->y_step;
  g.x_min = grid->x_min; g.y_min = grid->y_min; g.z_min = grid->z_min;
  g.x_max = grid->x_max; g.y_max = grid->y_max; g.z_max = grid->z_max;
  g.canvas_height = grid->canvas_height;
  const double fx = floor((grid->x_max - grid->x_min) / grid->x_step) + 1.0;
  const double fy = floor((grid->y_max - grid->y_min) / grid->y_step) + 1.0;
  if (!(fx >= 1 && fx <= 32768.0 && fy >= 1 && fy <= 32768.0)) return false;
  g.nx = (int)fx; g.ny = (int)fy;
  const long long nc = (long long)g.nx * g.ny;
  if (nc > (1ll << 28)) return false;
  g.ncell = (int)nc;
  return true;
}

template <class A>
static void layout(A& a, PillarWs* ws, int B, long long T, long long ntiles, int ncell, int P) {
  const size_t nc = (size_t)B * ncell, np = (size_t)B * P, t = (size_t)(T > 0 ? T : 1);
#define TAKE(field, type, count)                    \
  do {                                              \
    auto _p = a.template take<type>(count);         \
    if (ws) ws->field = (decltype(ws->field))_p;    \
  } while (0)
  // --- zero block start
  size_t z0 = a.used;
  TAKE(cell_count, int, nc);
  TAKE(pil_cursor, int, np);
  TAKE(list_cursor, int, PP_MAX_SWEEPS);
  TAKE(big_count, int, 1);
  TAKE(n_rows, int, 1);
  TAKE(num_pillars_scratch, int, PP_MAX_SWEEPS);
  if (ws) { ws->zero_bytes = a.used - z0; }
  // --- rest
  TAKE(cell_first, int, nc);
  TAKE(cell_slot, int, nc);
  TAKE(cell_of_point, int, t);
  TAKE(rank_of_point, int, t);
  TAKE(tile_count, int, (size_t)ntiles + 1);
  TAKE(list_u, int, t);
  TAKE(terms, double4, t);
  TAKE(pil_cnt, int, np);
  TAKE(pil_off, int, np);
  TAKE(pil_cell, int, np);
  TAKE(big_list, int, t / kBig + 2);
  TAKE(pil_mean, double, np * 3);
  TAKE(feat_c, float, t * 9);
#undef TAKE
}

struct SizeArena {
  size_t used = 0;
  template <class T>
  T* take(size_t count) { used += align_up(count * sizeof(T)); return nullptr; }
};

template <typename T>
static int run_stages(const T* pts, long long sp, long long sc, bool vec4, const SweepParams& sw,
                      const GridDev& g, int P, PillarWs& ws, int* d_num_pillars,
                      long long* d_indices, int* d_status, cudaStream_t st) {
  const long long total = sw.off[sw.n_sweeps];
  const int ntiles = sw.tile_start[sw.n_sweeps];
  const int B = sw.n_sweeps;
  PP_CUDA(cudaMemsetAsync(ws.cell_count, 0, ws.zero_bytes, st));
  PP_CUDA(cudaMemsetAsync(ws.cell_first, 0x7f, (size_t)B * g.ncell * sizeof(int), st));
  PP_CUDA(cudaMemsetAsync(d_num_pillars, 0, (size_t)B * sizeof(int), st));
  const int pt_blocks = (int)((total + 255) / 256);
  if (total > 0) {
    PP_KERNEL("k_bin", st, k_bin<T><<<pt_blocks, 256, 0, st>>>(pts, sp, sc, vec4, sw, g, ws.cell_of_point, ws.cell_first,
                                        ws.cell_count, d_status));
    PP_KERNEL("k_tilecount", st, k_tilecount<<<ntiles, kTile, 0, st>>>(sw, g, ws.cell_of_point, ws.cell_first, ws.tile_count));
    PP_KERNEL("k_assign", st, k_assign<<<ntiles, kTile, 0, st>>>(sw, g, P, ws.cell_of_point, ws.cell_first, ws.cell_count,
                                       ws.tile_count, ws.cell_slot, ws.pil_cnt, ws.pil_off,
                                       ws.pil_cell, ws.list_cursor, ws.big_count, ws.big_list,
                                       d_num_pillars));
    PP_KERNEL("k_scatter", st, k_scatter<<<pt_blocks, 256, 0, st>>>(sw, g, P, ws.cell_of_point, ws.cell_slot, ws.pil_off,
                                         ws.pil_cursor, ws.list_u));
    PP_KERNEL("k_rank", st, k_rank<T><<<pt_blocks, 256, 0, st>>>(pts, sp, sc, vec4, sw, g, P, ws.cell_of_point, ws.cell_slot,
                                         ws.pil_cnt, ws.pil_off, ws.list_u, ws.terms, ws.rank_of_point));
    PP_KERNEL("k_rank_big", st, k_rank_big<T><<<64, kTile, 0, st>>>(pts, sp, sc, vec4, sw, P, ws.cell_of_point, ws.pil_off,
                                        ws.pil_cell, ws.big_count, ws.big_list, ws.terms, ws.rank_of_point));
  }
  const long long groups = (long long)B * P;
  PP_KERNEL("k_mean", st, k_mean<<<(int)((groups + 127) / 128), 128, 0, st>>>(sw, g, P, d_num_pillars, ws.pil_cnt, ws.pil_off,
                                         ws.pil_cell, ws.terms, ws.pil_mean, d_indices));
  return PP_OK;
}

static int make_sweeps(const int64_t* h_off, int n_sweeps, SweepParams& sw) {
  if (h_off == nullptr || n_sweeps < 1 || n_sweeps > PP_MAX_SWEEPS) return PP_ERR_INVALID_ARG;
  sw.n_sweeps = n_sweeps;
  int tiles = 0;
  for (int s = 0; s <= n_sweeps; ++s) {
    sw.off[s] = h_off[s];
    if (s > 0) {
      const long long n = h_off[s] - h_off[s - 1];
      if (n < 0 || n > 0x7f000000ll) return PP_ERR_INVALID_ARG;
      sw.tile_start[s - 1] = tiles;
      tiles += (int)((n + kTile - 1) / kTile);
    }
  }
  sw.tile_start[n_sweeps] = tiles;
  if (h_off[0] != 0 || h_off[n_sweeps] > 0x7f000000ll) return PP_ERR_INVALID_ARG;
  return PP_OK;
}

static int emit_dense(const SweepParams& sw, int P, int N, const float* d_mean, float* d_x,
                      const int32_t* d_num_pillars, const PillarWs& ws, cudaStream_t st) {
  const long long PN = (long long)P * N;
  const bool v4 = (N % 4 == 0) && ((uintptr_t)d_x % 16) == 0 &&
                  (d_mean == nullptr || ((uintptr_t)d_mean % 16) == 0);
  const long long groups = v4 ? PN / 4 : PN;
  long long blocks = (groups + 255) / 256;
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  const dim3 egrid((unsigned)blocks, 9 / kFeatGroup);
  if (v4) {
    PP_KERNEL("k_emit_dense", st,
              k_emit_dense<4><<<egrid, 256, 0, st>>>(sw, P, N, d_mean, d_x, d_num_pillars, ws.pil_cnt,
                                                     ws.pil_off, ws.feat_c));
  } else {
    PP_KERNEL("k_emit_dense", st,
              k_emit_dense<1><<<egrid, 256, 0, st>>>(sw, P, N, d_mean, d_x, d_num_pillars, ws.pil_cnt,
                                                     ws.pil_off, ws.feat_c));
  }
  return PP_OK;
}

template <typename T>
static int pillarize_impl(const T* pts, long long sp, long long sc, const int64_t* h_off, int B,
                          const pp_grid* grid, int N, int P, const float* d_mean, float* d_x,
                          int64_t* d_indices, int32_t* d_num_pillars, int32_t* d_status, void* d_ws,
                          size_t ws_bytes, cudaStream_t st) {
  GridDev g;
  SweepParams sw;
  if (!make_grid(grid, g)) return PP_ERR_INVALID_ARG;
  int rc = make_sweeps(h_off, B, sw);
  if (rc != PP_OK) return rc;
  if (N < 1 || P < 1 || d_x == nullptr || d_indices == nullptr || d_num_pillars == nullptr ||
      d_status == nullptr || (pts == nullptr && sw.off[B] > 0) || (long long)P * N > 0x7fffffffll)
    return PP_ERR_INVALID_ARG;
  if ((long long)B * g.ncell > 0x7fffffffll || (long long)B * P > 0x3fffffffll) return PP_ERR_INVALID_ARG;
  Arena arena(d_ws, ws_bytes);
  PillarWs ws{};
  layout(arena, &ws, B, sw.off[B], sw.tile_start[B], g.ncell, P);
  if (!arena.ok) return PP_ERR_WORKSPACE;
  const bool vec4 = sizeof(T) == 4 && sp == 4 && sc == 1 && ((uintptr_t)pts % 16) == 0;
  rc = run_stages<T>(pts, sp, sc, vec4, sw, g, P, ws, d_num_pillars, (long long*)d_indices,
                     d_status, st);
  if (rc != PP_OK) return rc;
  const long long total = sw.off[B];
  if (total > 0) {
    PP_KERNEL("k_feat", st,
              k_feat<T><<<(int)((total + 255) / 256), 256, 0, st>>>(
                  pts, sp, sc, vec4, sw, g, P, N, ws.cell_of_point, ws.cell_slot, ws.rank_of_point,
                  ws.pil_off, ws.pil_mean, ws.feat_c));
  }
  rc = emit_dense(sw, P, N, d_mean, d_x, d_num_pillars, ws, st);
  if (rc != PP_OK) return rc;
  return PP_OK;
}

// pp_input_path: K1 stages + (optional) dense emit + sparse PFN + canvas
template <typename T>
static int input_path_impl(const T* pts, long long sp, long long sc, const int64_t* h_off, int B,
                           const pp_grid* grid, int N, int P, const float* d_mean, const void* d_mean_prep, int C,
                           const PfnParams& prm, int H, int W, float* d_canvas, float* d_x, int64_t* d_indices,
                           int32_t* d_num_pillars, int32_t* d_status, void* d_ws, size_t ws_bytes,
                           int stages, cudaStream_t st) {
  GridDev g;
  SweepParams sw;
  if (!make_grid(grid, g)) return PP_ERR_INVALID_ARG;
  int rc = make_sweeps(h_off, B, sw);
  if (rc != PP_OK) return rc;
  if ((stages & 3) == 0) return PP_ERR_INVALID_ARG;
  if (N < 1 || P < 1 || d_canvas == nullptr || d_indices == nullptr || d_num_pillars == nullptr ||
      d_status == nullptr || (pts == nullptr && sw.off[B] > 0) || (long long)P * N > 0x7fffffffll || H < 1 || W < 1)
    return PP_ERR_INVALID_ARG;
  if ((long long)B * g.ncell > 0x7fffffffll || (long long)B * P > 0x3fffffffll) return PP_ERR_INVALID_ARG;
  if (!pfn_sparse_supported(B, P, N, C, d_mean)) return PP_ERR_UNSUPPORTED;
  Arena arena(d_ws, ws_bytes);
  PillarWs ws{};
  layout(arena, &ws, B, sw.off[B], sw.tile_start[B], g.ncell, P);
  if (!arena.ok) return PP_ERR_WORKSPACE;
  const size_t k1_bytes = arena.used;
  if (stages & 1) {
    const bool vec4 = sizeof(T) == 4 && sp == 4 && sc == 1 && ((uintptr_t)pts % 16) == 0;
    rc = run_stages<T>(pts, sp, sc, vec4, sw, g, P, ws, d_num_pillars, (long long*)d_indices, d_status, st);
    if (rc != PP_OK) return rc;
    const long long total = sw.off[B];
    if (total > 0) {
      PP_KERNEL("k_feat", st,
                k_feat<T><<<(int)((total + 255) / 256), 256, 0, st>>>(
                    pts, sp, sc, vec4, sw, g, P, N, ws.cell_of_point, ws.cell_slot, ws.rank_of_point,
                    ws.pil_off, ws.pil_mean, ws.feat_c));
    }
    if (d_x != nullptr) {
      rc = emit_dense(sw, P, N, d_mean, d_x, d_num_pillars, ws, st);
      if (rc != PP_OK) return rc;
    }
  }
  if (!(stages & 2)) return PP_OK;
  CompactPillars cp;
  cp.sw = sw; cp.P = P; cp.N = N;
  cp.feat_c = ws.feat_c; cp.pil_cnt = ws.pil_cnt; cp.pil_off = ws.pil_off;
  cp.num_pillars = d_num_pillars; cp.data_mean = d_mean; cp.mean_prepared = d_mean != nullptr ? d_mean_prep : nullptr;
  return pfn_sparse_scatter(cp, d_indices, C, prm, H, W, d_canvas, d_status, (char*)d_ws + k1_bytes,
                            ws_bytes - k1_bytes, st);
}

template <typename T>
static int compact_impl(const T* pts, long long sp, long long sc, int64_t n_points,
                        const pp_grid* grid, int N, int P, double* d_rows, int32_t* d_slot,
                        int32_t* d_pillar_xy, int32_t* d_counts, int32_t* d_status, void* d_ws,
                        size_t ws_bytes, cudaStream_t st) {
  GridDev g;
  SweepParams sw;
  if (!make_grid(grid, g)) return PP_ERR_INVALID_ARG;
  const int64_t h_off[2] = {0, n_points};
  int rc = make_sweeps(h_off, 1, sw);
  if (rc != PP_OK) return rc;
  if (N < 1 || P < 1 || d_rows == nullptr || d_slot == nullptr || d_pillar_xy == nullptr ||
      d_counts == nullptr || d_status == nullptr || (pts == nullptr && n_points > 0))
    return PP_ERR_INVALID_ARG;
  if ((long long)P * N > 0x7fffffffll) return PP_ERR_INVALID_ARG;
  Arena arena(d_ws, ws_bytes);
  PillarWs ws{};
  layout(arena, &ws, 1, n_points, sw.tile_start[1], g.ncell, P);
  if (!arena.ok) return PP_ERR_WORKSPACE;
  const bool vec4 = sizeof(T) == 4 && sp == 4 && sc == 1 && ((uintptr_t)pts % 16) == 0;
  rc = run_stages<T>(pts, sp, sc, vec4, sw, g, P, ws, ws.num_pillars_scratch, nullptr, d_status, st);
  if (rc != PP_OK) return rc;
  if (n_points > 0) {
    PP_CUDA(cudaMemsetAsync(d_slot, 0xff, (size_t)n_points * sizeof(int), st));
    const int blocks = (int)((n_points + 255) / 256);
    PP_KERNEL("k_emit_compact", st,
              k_emit_compact<T><<<blocks, 256, 0, st>>>(pts, sp, sc, vec4, sw, g, P, N,
                                                        ws.cell_of_point, ws.cell_slot,
                                                        ws.rank_of_point, ws.pil_off, ws.pil_mean,
                                                        d_rows, d_slot, ws.n_rows));
  }
  PP_KERNEL("k_pillar_xy", st, k_pillar_xy<<<(P + 255) / 256, 256, 0, st>>>(g, P, ws.num_pillars_scratch, ws.pil_cell,
                                               d_pillar_xy, d_counts));
  PP_CUDA(cudaMemcpyAsync(d_counts + 1, ws.n_rows, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return PP_OK;
}

}  // namespace pp

extern "C" {

size_t pp_pillarize_workspace_bytes(int32_t n_sweeps, int64_t total_points, const pp_grid* grid,
                                    int32_t max_pillars) {
  pp::GridDev g;
  if (!pp::make_grid(grid, g) || n_sweeps < 1 || n_sweeps > PP_MAX_SWEEPS || total_points < 0 ||
      max_pillars < 1)
    return 0;
  pp::SizeArena a;
  const long long ntiles = total_points / pp::kTile + n_sweeps + 1;
  pp::layout(a, (pp::PillarWs*)nullptr, n_sweeps, total_points, ntiles, g.ncell, max_pillars);
  return a.used + pp::kAlign;
}

int pp_pillarize(const void* d_points, int32_t point_dtype, int64_t stride_point,
                 int64_t stride_col, const int64_t* h_sweep_offsets, int32_t n_sweeps,
                 const pp_grid* grid, int32_t max_points_per_pillar, int32_t max_pillars,
                 const float* d_data_mean, float* d_x, int64_t* d_indices, int32_t* d_num_pillars,
                 int32_t* d_status, void* d_workspace, size_t workspace_bytes, pp_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (point_dtype == PP_F32)
    return pp::pillarize_impl<float>((const float*)d_points, stride_point, stride_col,
                                     h_sweep_offsets, n_sweeps, grid, max_points_per_pillar,
                                     max_pillars, d_data_mean, d_x, d_indices, d_num_pillars,
                                     d_status, d_workspace, workspace_bytes, st);
  if (point_dtype == PP_F64)
    return pp::pillarize_impl<double>((const double*)d_points, stride_point, stride_col,
                                      h_sweep_offsets, n_sweeps, grid, max_points_per_pillar,
                                      max_pillars, d_data_mean, d_x, d_indices, d_num_pillars,
                                      d_status, d_workspace, workspace_bytes, st);
  return PP_ERR_INVALID_ARG;
}

size_t pp_input_path_workspace_bytes(int32_t n_sweeps, int64_t total_points, const pp_grid* grid,
                                     int32_t max_points_per_pillar, int32_t max_pillars, int32_t C, int32_t canvas_h,
                                     int32_t canvas_w, int32_t prepare_in_workspace) {
  const size_t k1 = pp_pillarize_workspace_bytes(n_sweeps, total_points, grid, max_pillars);
  if (k1 == 0 || C < 1 || canvas_h < 1 || canvas_w < 1 || max_points_per_pillar < 1) return 0;
  return k1 + pp::pfn_sparse_workspace_bytes(n_sweeps, max_pillars, max_points_per_pillar, C, canvas_h, canvas_w,
                                             prepare_in_workspace != 0);
}

size_t pp_mean_prepared_bytes(int32_t max_pillars, int32_t max_points_per_pillar) {
  if (max_pillars < 1 || max_points_per_pillar < 1) return 0;
  return pp::mean_prepared_bytes(max_pillars, max_points_per_pillar);
}

int pp_mean_prepare(const float* d_data_mean, int32_t max_pillars, int32_t max_points_per_pillar, void* d_prepared,
                    size_t prepared_bytes, pp_stream_t stream) {
  return pp::mean_prepare(d_data_mean, max_pillars, max_points_per_pillar, d_prepared, prepared_bytes, (cudaStream_t)stream);
}

int pp_input_path(const void* d_points, int32_t point_dtype, int64_t stride_point, int64_t stride_col,
                  const int64_t* h_sweep_offsets, int32_t n_sweeps, const pp_grid* grid,
                  int32_t max_points_per_pillar, int32_t max_pillars, const float* d_data_mean,
                  const void* d_mean_prepared, int32_t C,
                  const float* d_conv_w, const float* d_conv_b, const float* d_bn_w, const float* d_bn_b,
                  float* d_running_mean, float* d_running_var, int64_t* d_num_batches_tracked,
                  int32_t training, float momentum, float eps, int32_t canvas_h, int32_t canvas_w,
                  float* d_canvas, float* d_x, int64_t* d_indices, int32_t* d_num_pillars,
                  int32_t* d_status, void* d_workspace, size_t workspace_bytes, int32_t stages,
                  pp_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_conv_w || !d_conv_b || !d_bn_w || !d_bn_b || !d_running_mean || !d_running_var)
    return PP_ERR_INVALID_ARG;
  pp::PfnParams prm{d_conv_w, d_conv_b, d_bn_w, d_bn_b, d_running_mean, d_running_var,
                    d_num_batches_tracked, training, momentum, eps};
  if (point_dtype == PP_F32)
    return pp::input_path_impl<float>((const float*)d_points, stride_point, stride_col, h_sweep_offsets,
                                      n_sweeps, grid, max_points_per_pillar, max_pillars, d_data_mean, d_mean_prepared, C, prm,
                                      canvas_h, canvas_w, d_canvas, d_x, d_indices, d_num_pillars, d_status,
                                      d_workspace, workspace_bytes, stages, st);
  if (point_dtype == PP_F64)
    return pp::input_path_impl<double>((const double*)d_points, stride_point, stride_col, h_sweep_offsets,
                                       n_sweeps, grid, max_points_per_pillar, max_pillars, d_data_mean, d_mean_prepared, C, prm,
                                       canvas_h, canvas_w, d_canvas, d_x, d_indices, d_num_pillars, d_status,
                                       d_workspace, workspace_bytes, stages, st);
  return PP_ERR_INVALID_ARG;
}

size_t pp_input_path_backward_workspace_bytes(int32_t n_sweeps, int32_t max_pillars, int32_t C) {
  if (n_sweeps < 1 || n_sweeps > PP_MAX_SWEEPS || max_pillars < 1 || C != 64) return 0;
  return pp::pfn_sparse_backward_workspace_bytes(n_sweeps, max_pillars);
}

int pp_input_path_backward(const int64_t* h_sweep_offsets, int32_t n_sweeps, const pp_grid* grid,
                           int32_t max_points_per_pillar, int32_t max_pillars, const float* d_data_mean, int32_t C,
                           const float* d_conv_w, const float* d_conv_b, const float* d_bn_w,
                           const float* d_running_mean, const float* d_running_var, int32_t training, float eps,
                           int32_t canvas_h, int32_t canvas_w, const float* d_grad_canvas, const int64_t* d_indices,
                           const int32_t* d_num_pillars, float* d_grad_conv_w, float* d_grad_conv_b,
                           float* d_grad_bn_w, float* d_grad_bn_b, const void* d_forward_workspace,
                           size_t forward_workspace_bytes, void* d_workspace, size_t workspace_bytes,
                           pp_stream_t stream) {
  using namespace pp;
  cudaStream_t st = (cudaStream_t)stream;
  GridDev g;
  SweepParams sw;
  if (!make_grid(grid, g)) return PP_ERR_INVALID_ARG;
  int rc = make_sweeps(h_sweep_offsets, n_sweeps, sw);
  if (rc != PP_OK) return rc;
  const int N = max_points_per_pillar, P = max_pillars;
  if (N < 1 || P < 1 || !d_indices || !d_num_pillars || !d_forward_workspace || canvas_h < 1 || canvas_w < 1)
    return PP_ERR_INVALID_ARG;
  if (!pfn_sparse_supported(n_sweeps, P, N, C, d_data_mean)) return PP_ERR_UNSUPPORTED;
  // the forward's K1 state, found by laying the same workspace out again (pp_input_path, stage 1)
  Arena arena(const_cast<void*>(d_forward_workspace), forward_workspace_bytes);
  PillarWs ws{};
  layout(arena, &ws, n_sweeps, sw.off[n_sweeps], sw.tile_start[n_sweeps], g.ncell, P);
  if (!arena.ok) return PP_ERR_WORKSPACE;
  CompactPillars cp;
  cp.sw = sw; cp.P = P; cp.N = N;
  cp.feat_c = ws.feat_c; cp.pil_cnt = ws.pil_cnt; cp.pil_off = ws.pil_off;
  cp.num_pillars = d_num_pillars; cp.data_mean = d_data_mean; cp.mean_prepared = nullptr;
  return pfn_sparse_backward(cp, d_indices, C, d_conv_w, d_conv_b, d_bn_w, d_running_mean, d_running_var, training, eps,
                             canvas_h, canvas_w, d_grad_canvas, d_grad_conv_w, d_grad_conv_b, d_grad_bn_w, d_grad_bn_b,
                             d_workspace, workspace_bytes, st);
}

int pp_pillarize_compact(const void* d_points, int32_t point_dtype, int64_t stride_point,
                         int64_t stride_col, int64_t n_points, const pp_grid* grid,
                         int32_t max_points_per_pillar, int32_t max_pillars, double* d_rows,
                         int32_t* d_slot, int32_t* d_pillar_xy, int32_t* d_counts,
                         int32_t* d_status, void* d_workspace, size_t workspace_bytes,
                         pp_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_points < 0) return PP_ERR_INVALID_ARG;
  if (point_dtype == PP_F32)
    return pp::compact_impl<float>((const float*)d_points, stride_point, stride_col, n_points, grid,
                                   max_points_per_pillar, max_pillars, d_rows, d_slot, d_pillar_xy,
                                   d_counts, d_status, d_workspace, workspace_bytes, st);
  if (point_dtype == PP_F64)
    return pp::compact_impl<double>((const double*)d_points, stride_point, stride_col, n_points,
                                    grid, max_points_per_pillar, max_pillars, d_rows, d_slot,
                                    d_pillar_xy, d_counts, d_status, d_workspace, workspace_bytes,
                                    st);
  return PP_ERR_INVALID_ARG;
}

}  // extern "C"
