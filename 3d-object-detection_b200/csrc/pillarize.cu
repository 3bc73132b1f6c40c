// pillarize.cu -- K1: point -> pillar binning, order-exact compaction, decoration and per-slot
// mean subtraction for sm_100a.  Replaces data/pillars.cpp:236-398 (create_pillars) and the
// tensor glue of data/dataset.py:99-106 of the reference.
//
// Order-exactness: pillars are numbered by the input index of their first
// in-range point; inside a pillar points keep input order; the first N are emitted; the mean is
// the reference's sequential running mean over ALL in-range points of the pillar.  atomics only
// ever decide things that do not depend on order (first index, counts, list placement before the
// rank pass), so the result is deterministic and identical to the sequential CPU algorithm.
//
// Stages (one launch each after ONE memset of the workspace's zero block; all sweeps of the batch in the same launch):
//   k_bin      per point: range filter + floor binning in fp64; first index (atomicMax of the inverted index) and
//              count of its cell, one pair of atomics per distinct cell of a warp
//   k_assign   per 1024-point tile (ticket order): first-touch count published to the later tiles, exclusive scan of
//              the first-touch flags -> pillar slot of each cell, list segment of each kept pillar, long / big lists
//   k_scatter  per point: append its index to its pillar's segment (unordered); point -> pillar record
//   k_rank     rank = number of smaller indices in the segment (per point / 128 threads per long pillar / block scan
//              for pillars with more than kBig points); the point's running-mean terms (four IEEE divisions) go to
//              its place in the ordered segment
//   k_mean     sequential multiply-add chain of the running mean (one thread per pillar, one warp per long pillar);
//              indices row and canvas cell map
//   k_feat     per kept point: the nine decorated features, fp64 -> one rounding to fp32
//   k_emit_*   dense [B,9,P,N] float with fused "- data_mean", or compact fp64 rows
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
// The stages are latency-bound single waves of dependent, mostly uncoalesced accesses (~10 MB of traffic in
// all): what they cost is launch + drain latency and the number of memory WAVEFRONTS the SM's load/store unit
// has to take (a fully divergent warp instruction = 32 of them), not bytes.  Hence: warp-aggregated atomics, one
// coalesced 16-byte point -> pillar record (point_seg) instead of the cell -> slot -> pillar chain in every stage,
// 48-byte feature records written as three aligned float4, long running-mean chains on their own warps.
// (One persistent cooperative kernel with grid barriers instead of six launches was tried: 68 us instead of ~85,
// but it owns every SM while it spins, so target assignment on the side stream no longer ran next to it and the
// step got slower, 474 us per batch instead of 357.)
template <typename T>
struct K1 {
  const T* pts;
  long long sp, sc;
  bool vec4;
  SweepParams sw;
  GridDev g;
  int P, N;
  int *cell_of_point, *cell_first, *cell_count, *tile_count, *cell_slot;
  int *pil_cnt, *pil_off, *pil_cell, *pil_cursor, *list_cursor, *list_u, *rank_of_point;
  int *big_count, *big_list, *long_count, *long_list, *tile_ticket;
  int* map;                // [B, map_h * map_w] canvas cell -> slot + 1 (null: not wanted)
  int map_h, map_w;
  int2* pil_oc;            // per pillar {segment offset, count} (one 8-byte read in st_scatter)
  int4* point_seg;         // per point {absolute segment start, pillar index or -1, count, 0}
  double4* terms;
  double* pil_mean;
  long long* indices;      // may be null (compact path)
  int* num_pillars;
  float* feat_c;           // null: no decoration stage (compact path emits fp64 rows itself)
  int* status;
};

constexpr int kMeanLong = 32;   // pillars with more points run their running-mean chain warp-cooperatively
constexpr unsigned kTileReady = 0x80000000u;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- stage: range filter + floor binning in fp64, first index and count of every cell
template <typename T>
__device__ __forceinline__ void st_bin(const K1<T>& a, long long i0, long long stride) {
  const SweepParams& sw = a.sw;
  const GridDev& g = a.g;
  const long long total = sw.off[sw.n_sweeps];
  const int lane = (int)lane_id();
  for (long long base = i0 - lane; base < total; base += stride) {      // warp-uniform trip count
    const long long i = base + lane;
    int cell = -1, b = 0;
    if (i < total) {
      b = find_sweep(sw, i);
      double x, y, z, r;
      load_xyz(a.pts, i, a.sp, a.sc, a.vec4, x, y, z, r);
      // data/pillars.cpp:271-275 (half-open box; written as the reference writes it)
      if (!((x >= g.x_max) || (x < g.x_min) || (y >= g.y_max) || (y < g.y_min) ||
            (z >= g.z_max) || (z < g.z_min))) {
        // data/pillars.cpp:278-279, IEEE double subtract / divide / floor
        const double fx = floor(__ddiv_rn(__dsub_rn(x, g.x_min), g.x_step));
        const double fy = floor(__ddiv_rn(__dsub_rn(y, g.y_min), g.y_step));
        if (fx >= 0.0 && fx < (double)g.nx && fy >= 0.0 && fy < (double)g.ny) {
          cell = (int)fy * g.nx + (int)fx;
        } else {
          atomicOr(a.status, PP_STATUS_BAD_POINT);  // NaN/Inf: out of contract, dropped
        }
      }
      a.cell_of_point[i] = cell;
    }
    // neighbours in the input often share a cell: one pair of atomics per distinct (sweep, cell) of the warp.
    // The group's lowest lane holds its smallest index (lanes are in input order).
    const int key = cell >= 0 ? b * g.ncell + cell : -1 - lane;          // < 2^31 (checked by the host)
    const unsigned grp = __match_any_sync(0xffffffffu, key);
    if (cell >= 0 && (grp & ((1u << lane) - 1u)) == 0u) {
      const int il = (int)(i - sw.off[b]);
      // first index as an inverted maximum: the table starts at zero like the counters (one memset)
      atomicMax(&a.cell_first[(size_t)key], 0x7fffffff - il);
      atomicAdd(&a.cell_count[(size_t)key], __popc(grp));
    }
  }
}

template <typename T>
__device__ __forceinline__ bool first_touch_flag(const K1<T>& a, int t, int& b, int& il, int& cell) {
  const SweepParams& sw = a.sw;
  b = find_tile_sweep(sw, t);
  il = (t - sw.tile_start[b]) * kTile + (int)threadIdx.x;
  const long long n_b = sw.off[b + 1] - sw.off[b];
  cell = -1;
  if (il < n_b) cell = a.cell_of_point[sw.off[b] + il];
  return cell >= 0 && a.cell_first[(size_t)b * a.g.ncell + cell] == 0x7fffffff - il;
}

// ---- stage: number of first-touch points of tile t, published with a ready bit (st_assign of a later tile
// of the same sweep polls it: every block publishes all its tiles before it waits for anything)
template <typename T>
__device__ __forceinline__ void st_tilecount(const K1<T>& a, int t, bool flag) {
  __shared__ int warp_cnt[kTile / 32];
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (lane_id() == 0) warp_cnt[threadIdx.x >> 5] = __popc(m);
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = warp_cnt[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) {
      __threadfence();
      atomicExch(reinterpret_cast<unsigned*>(&a.tile_count[t]), (unsigned)v | kTileReady);
    }
  }
  __syncthreads();
}

// ---- stage: pillar slot of every first-touch point (exclusive scan in input order), list segments
template <typename T>
__device__ __forceinline__ void st_assign(const K1<T>& a, int t, bool flag, int b, int cell, size_t ci, int cnt_cell) {
  __shared__ int warp_sum[kTile / 32];
  __shared__ int warp_cnt[kTile / 32];
  __shared__ int s_base, s_seg_base;
  const SweepParams& sw = a.sw;
  const GridDev& g = a.g;
  const int P = a.P;
  // base = number of first-touch points in the earlier tiles of this sweep
  int part = 0;
  for (int k = sw.tile_start[b] + (int)threadIdx.x; k < t; k += kTile) {
    unsigned v;
    const long long t0 = clock64();
    while (((v = ld_acquire(reinterpret_cast<const unsigned*>(&a.tile_count[k]))) & kTileReady) == 0u) {
      if (clock64() - t0 > 4000000000ll) __trap();
    }
    part += (int)(v & ~kTileReady);
  }
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
  if (flag) {
    slot = base + before + in_warp;
    if (slot < P) cnt = cnt_cell;
  }
  int incl = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)lane_id() >= o) incl += v;
  }
  if (lane_id() == 31) warp_cnt[threadIdx.x >> 5] = incl;
  __syncthreads();
  int cnt_before = 0, cnt_total = 0;
  for (int w = 0; w < kTile / 32; ++w) {
    const int v = warp_cnt[w];
    if (w < (int)(threadIdx.x >> 5)) cnt_before += v;
    cnt_total += v;
  }
  if (threadIdx.x == 0) s_seg_base = cnt_total > 0 ? atomicAdd(&a.list_cursor[b], cnt_total) : 0;
  __syncthreads();
  if (flag) {
    if (slot < P) {
      const int off = s_seg_base + cnt_before + incl - cnt;
      const size_t pi = (size_t)b * P + slot;
      a.cell_slot[ci] = slot;
      a.pil_cnt[pi] = cnt;
      a.pil_off[pi] = off;
      a.pil_oc[pi] = make_int2(off, cnt);
      a.pil_cell[pi] = cell;
      if (cnt > kBig) a.big_list[atomicAdd(a.big_count, 1)] = (int)pi;
      else if (cnt > kMeanLong) a.long_list[atomicAdd(a.long_count, 1)] = (int)pi;
    } else {
      a.cell_slot[ci] = -1;  // pillar beyond the max_pillars cap: data/pillars.cpp:339
    }
  }
  const int ntiles_b = sw.tile_start[b + 1] - sw.tile_start[b];
  if (threadIdx.x == 0 && t - sw.tile_start[b] == ntiles_b - 1) {
    const int np = base + total;
    a.num_pillars[b] = np < P ? np : P;
  }
  __syncthreads();   // shared scratch is reused by the block's next tile
}

// ---- stage: append every kept point to its pillar's segment (unordered).  The point's pillar, segment and count
// go to point_seg with one coalesced 16-byte store: the later per-point stages start from there instead of
// repeating the cell -> slot -> pillar chain of random 4-byte reads (these stages are bound by the number of
// memory wavefronts the SM's load/store unit takes for uncoalesced accesses, not by bytes).
template <typename T>
__device__ __forceinline__ void st_scatter(const K1<T>& a, long long i0, long long stride) {
  const SweepParams& sw = a.sw;
  const long long total = sw.off[sw.n_sweeps];
  const int lane = (int)lane_id();
  for (long long base = i0 - lane; base < total; base += stride) {      // warp-uniform trip count
    const long long i = base + lane;
    int pi = -1, b = 0;
    if (i < total) {
      const int cell = a.cell_of_point[i];
      if (cell >= 0) {
        b = find_sweep(sw, i);
        const int slot = a.cell_slot[(size_t)b * a.g.ncell + cell];
        if (slot >= 0) pi = b * a.P + slot;
      }
    }
    // one cursor atomic per pillar of the warp
    const unsigned grp = __match_any_sync(0xffffffffu, pi >= 0 ? pi : -1 - lane);
    const unsigned lower = grp & ((1u << lane) - 1u);
    int pos = 0;
    if (pi >= 0 && lower == 0u) pos = atomicAdd(&a.pil_cursor[pi], __popc(grp));
    pos = __shfl_sync(0xffffffffu, pos, __ffs(grp) - 1) + __popc(lower);
    if (i < total) {
      int4 ps = make_int4(0, -1, 0, 0);
      if (pi >= 0) {
        const int2 oc = a.pil_oc[pi];
        const int segbase = (int)sw.off[b] + oc.x;
        a.list_u[segbase + pos] = (int)(i - sw.off[b]);
        ps = make_int4(segbase, pi, oc.y, 0);
      }
      a.point_seg[i] = ps;
    }
  }
}

// one 256-bit access per 32-byte record (sm_100: LDG.256 / STG.256): a record fetched by a single lane costs one
// memory wavefront instead of two
__device__ __forceinline__ double4 ld_terms(const double4* __restrict__ p) {
  double4 v;
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_terms(double4* p, const double4& v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}

// Per-point terms of the reference's running mean (data/pillars.cpp:311-328)
//   m <- m*(n/(n+1)) + x/(n+1)
// for the point of rank n: {a = n/(n+1), x/(n+1), y/(n+1), z/(n+1)} (the first point initialises the
// mean: a = 0, terms = x, y, z).  The four IEEE divisions do not depend on m, so they are done here,
// one point per thread with every lane busy, and the mean stage is left with the bare multiply-add chain.
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
  st_terms(out, t);
}

// ---- stage: rank of every kept point inside its pillar = number of smaller indices in the segment.
// One thread per point for pillars up to kMeanLong points; longer ones by st_rank_long / st_rank_big.
template <typename T>
__device__ __forceinline__ void st_rank(const K1<T>& a, long long i0, long long stride) {
  const SweepParams& sw = a.sw;
  const long long total = sw.off[sw.n_sweeps];
  for (long long i = i0; i < total; i += stride) {
    const int4 ps = a.point_seg[i];
    if (ps.y < 0) { a.rank_of_point[i] = -1; continue; }
    const int c = ps.z;
    if (c > kMeanLong) continue;
    const int il = (int)(i - sw.off[find_sweep(sw, i)]);
    const int* seg = a.list_u + ps.x;
    int rank = 0;
    for (int k = 0; k < c; ++k) rank += (seg[k] < il) ? 1 : 0;
    mean_terms(a.pts, i, a.sp, a.sc, a.vec4, rank, a.terms + ps.x + rank);
    a.rank_of_point[i] = rank;
  }
}

// Pillars with kMeanLong < c <= kBig points: 128 threads each.  The segment's indices go to shared memory once and
// every thread ranks its element against broadcast 16-byte reads (one wavefront serves a whole warp).  As one
// thread per point every lane of a warp walked a different pillar's segment: c wavefronts per point, 200 us of
// the dense-cloud configuration (10-sweep clouds, most pillars hold 100+ points).  (One WARP per pillar left a
// 12k-instruction serial tail on the longest pillar: +10 us on the default configuration.)
constexpr int kRankBlock = 256;
constexpr int kRankGroup = 128;      // threads per long pillar: two pillars per block, each behind its own named barrier
template <typename T>
__device__ __forceinline__ void st_rank_long(const K1<T>& a, int long_blocks) {
  __shared__ __align__(16) int s_seg_all[kRankBlock / kRankGroup][kBig + 4];
  const SweepParams& sw = a.sw;
  const int grp_id = (int)threadIdx.x / kRankGroup, t = (int)threadIdx.x % kRankGroup;
  int* s_seg = s_seg_all[grp_id];
  const int ngroups = long_blocks * (kRankBlock / kRankGroup);
  const int nl = __ldcg(a.long_count);
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp_id), "r"(kRankGroup) : "memory"); };
  for (int w = (int)blockIdx.x + grp_id * long_blocks; w < nl; w += ngroups) {
    const long long grp = a.long_list[w];
    const int b = (int)(grp / a.P);
    const int2 oc = a.pil_oc[grp];
    const int c = oc.y, c4 = (c + 3) & ~3;
    const long long segbase = sw.off[b] + oc.x;
    group_sync();                                        // the previous pillar's readers are done
    for (int k = t; k < c4; k += kRankGroup) s_seg[k] = k < c ? a.list_u[segbase + k] : 0x7fffffff;
    group_sync();
    for (int e = t; e < c; e += kRankGroup) {
      const int il = s_seg[e];
      int rank = 0;
#pragma unroll 4
      for (int k = 0; k < c4; k += 4) {
        const int4 v = *reinterpret_cast<const int4*>(s_seg + k);
        rank += (v.x < il) + (v.y < il) + (v.z < il) + (v.w < il);
      }
      const long long gi = sw.off[b] + il;
      mean_terms(a.pts, gi, a.sp, a.sc, a.vec4, rank, a.terms + segbase + rank);
      a.rank_of_point[gi] = rank;
    }
  }
}

// One block per big pillar: ordered stream compaction of the sweep's points that fall in its cell.
template <typename T>
__device__ __forceinline__ void st_rank_big(const K1<T>& a) {
  __shared__ int warp_sum[32];
  const SweepParams& sw = a.sw;
  const int nbig = __ldcg(a.big_count);
  const int nwarps = (int)(blockDim.x >> 5);
  for (int k = blockIdx.x; k < nbig; k += gridDim.x) {
    const int pi = a.big_list[k];
    const int b = pi / a.P;
    const int cell = a.pil_cell[pi];
    const long long n_b = sw.off[b + 1] - sw.off[b];
    double4* out = a.terms + sw.off[b] + a.pil_off[pi];
    int running = 0;
    for (long long s = 0; s < n_b; s += blockDim.x) {
      const long long il = s + threadIdx.x;
      const bool flag = il < n_b && a.cell_of_point[sw.off[b] + il] == cell;
      const unsigned m = __ballot_sync(0xffffffffu, flag);
      const int in_warp = __popc(m & ((1u << lane_id()) - 1u));
      __syncthreads();
      if (lane_id() == 0) warp_sum[threadIdx.x >> 5] = __popc(m);
      __syncthreads();
      int before = 0, total = 0;
      for (int w = 0; w < nwarps; ++w) {
        const int v = warp_sum[w];
        if (w < (int)(threadIdx.x >> 5)) before += v;
        total += v;
      }
      if (flag) {
        const int rank = running + before + in_warp;
        mean_terms(a.pts, sw.off[b] + il, a.sp, a.sc, a.vec4, rank, out + rank);
        a.rank_of_point[sw.off[b] + il] = rank;
      }
      running += total;
    }
    __syncthreads();
  }
}


template <typename T>
__device__ __forceinline__ void mean_store(const K1<T>& a, long long grp, double m0, double m1, double m2) {
  st_terms(reinterpret_cast<double4*>(a.pil_mean) + grp, make_double4(m0, m1, m2, 0.0));   // 32-byte record per pillar
  if (a.indices != nullptr) {
    const int cell = a.pil_cell[grp];
    const double cx = (double)(cell % a.g.nx);
    const double cy = __dsub_rn(__dsub_rn(a.g.canvas_height, 1.0), (double)(cell / a.g.nx));
    a.indices[grp * 3 + 0] = 1;                 // data/pillars.cpp:390-392, then .long()
    a.indices[grp * 3 + 1] = (long long)cx;
    a.indices[grp * 3 + 2] = (long long)cy;
    if (a.map != nullptr) {                     // model/model.py:59-61: out[b, :, y_inds, x_inds]; one pillar per cell
      const long long ix = (long long)cx, iy = (long long)cy;
      if (ix < 0 || ix >= a.map_w || iy < 0 || iy >= a.map_h) {
        atomicOr(a.status, PP_STATUS_BAD_INDEX);
      } else {
        const long long b = grp / a.P;
        a.map[(size_t)b * a.map_h * a.map_w + (size_t)iy * a.map_w + (size_t)ix] = (int)(grp - b * a.P) + 1;
      }
    }
  }
}

// ---- stage: the sequential multiply-add chain of the reference's running mean over the terms st_rank
// prepared (32-byte records, contiguous per pillar, L2-resident), in input order with separately rounded IEEE
// operations.  One thread per pillar slot up to kMeanLong points (records loaded four steps ahead of the
// chain); longer pillars (0.4 % of them, up to a few hundred points) take a whole warp (st_mean_long).  As one
// thread per pillar the longest pillar alone set the stage's duration (31 us).
constexpr int kMeanBlock = 256;

template <typename T>
__device__ __forceinline__ void st_mean_slots(const K1<T>& a, long long grp) {
  const SweepParams& sw = a.sw;
  if (grp >= (long long)sw.n_sweeps * a.P) return;
  const int b = (int)(grp / a.P);
  const int slot = (int)(grp - (long long)b * a.P);
  if (slot >= a.num_pillars[b]) {
    if (a.indices != nullptr) { a.indices[grp * 3 + 0] = 0; a.indices[grp * 3 + 1] = 0; a.indices[grp * 3 + 2] = 0; }
    return;
  }
  const int2 oc = a.pil_oc[grp];
  const int c = oc.y;
  if (c > kMeanLong) return;
  const double4* seg = a.terms + sw.off[b] + oc.x;
  double4 t = ld_terms(seg);                          // rank 0: a = 0, terms = the point itself
  double m0 = t.y, m1 = t.z, m2 = t.w;             // data/pillars.cpp:313-317
  constexpr int kAhead = 4;
  double4 u[kAhead];
#pragma unroll
  for (int j = 0; j < kAhead; ++j)
    if (1 + j < c) u[j] = ld_terms(seg + 1 + j);
  for (int k = 1; k < c; k += kAhead) {
#pragma unroll
    for (int j = 0; j < kAhead; ++j) {
      if (k + j < c) {
        const double4 v = u[j];
        if (k + j + kAhead < c) u[j] = ld_terms(seg + k + j + kAhead);
        m0 = __dadd_rn(__dmul_rn(m0, v.x), v.y);   // data/pillars.cpp:324-326
        m1 = __dadd_rn(__dmul_rn(m1, v.x), v.z);
        m2 = __dadd_rn(__dmul_rn(m2, v.x), v.w);
      }
    }
  }
  mean_store(a, grp, m0, m1, m2);
}

// Long pillars (more than kMeanLong points; the big ones of st_rank_big included): one warp each.  The lanes load
// 32 records at once, park them in shared memory and every lane walks the chain over broadcast reads: two LDS.128
// and six fp64 operations per step (~47 cycles) instead of an L2 round trip every four steps.
template <typename T>
__device__ __forceinline__ void st_mean_long(const K1<T>& a, int w0, int wstride) {
  __shared__ double4 s_rec[kMeanBlock / 32][32];
  const SweepParams& sw = a.sw;
  const unsigned lane = lane_id();
  double4* rec = s_rec[threadIdx.x >> 5];
  const int nl0 = __ldcg(a.long_count), nlong = nl0 + __ldcg(a.big_count);
  for (int w = w0; w < nlong; w += wstride) {
    const long long grp = w < nl0 ? a.long_list[w] : a.big_list[w - nl0];
    const int b = (int)(grp / a.P);
    const int2 oc = a.pil_oc[grp];
    const int c = oc.y;
    const double4* seg = a.terms + sw.off[b] + oc.x;
    double4 cur = make_double4(0.0, 0.0, 0.0, 0.0);
    if ((int)lane < c) cur = ld_terms(seg + lane);
    double m0 = 0.0, m1 = 0.0, m2 = 0.0;
    for (int k0 = 0; k0 < c; k0 += 32) {
      rec[lane] = cur;
      __syncwarp();
      if (k0 + 32 + (int)lane < c) cur = ld_terms(seg + k0 + 32 + lane);      // in flight during the chain below
      const int n = c - k0 < 32 ? c - k0 : 32;
      int j = 0;
      if (k0 == 0) {                                    // rank 0: the point itself (data/pillars.cpp:313-317)
        const double4 v = rec[0];
        m0 = v.y; m1 = v.z; m2 = v.w;
        j = 1;
      }
#pragma unroll 4
      for (; j < n; ++j) {
        const double4 v = rec[j];
        m0 = __dadd_rn(__dmul_rn(m0, v.x), v.y);       // data/pillars.cpp:324-326
        m1 = __dadd_rn(__dmul_rn(m1, v.x), v.z);
        m2 = __dadd_rn(__dmul_rn(m2, v.x), v.w);
      }
      __syncwarp();                                     // all lanes have read the chunk before it is overwritten
    }
    if (lane == 0) mean_store(a, grp, m0, m1, m2);
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

// ---- stage: per kept point (rank < N inside a kept pillar) the nine decorated features in fp64, rounded once
// to fp32 (torch .float(), data/dataset.py:101), stored compactly at the point's position in its pillar's
// ordered segment.  Keeps all fp64 math out of the streaming kernels downstream.
template <typename T>
__device__ __forceinline__ void st_feat(const K1<T>& a, long long i0, long long stride) {
  const SweepParams& sw = a.sw;
  const GridDev& g = a.g;
  const long long total = sw.off[sw.n_sweeps];
  for (long long i = i0; i < total; i += stride) {
    const int4 ps = a.point_seg[i];
    if (ps.y < 0) continue;
    const int rank = a.rank_of_point[i];
    if (rank >= a.N) continue;                       // data/pillars.cpp:371 first-N cap
    const int cell = a.cell_of_point[i];
    double x, y, z, r, ft[9], mean[3];
    load_xyz(a.pts, i, a.sp, a.sc, a.vec4, x, y, z, r);
    const double cx = (double)(cell % g.nx);
    const double cy = __dsub_rn(__dsub_rn(g.canvas_height, 1.0), (double)(cell / g.nx));
    const double4 pm = ld_terms(reinterpret_cast<const double4*>(a.pil_mean) + ps.y);
    mean[0] = pm.x; mean[1] = pm.y; mean[2] = pm.z;
    point_features(x, y, z, r, cx, cy, mean, ft);
    // 48-byte records: three aligned 16-byte stores instead of nine scattered 4-byte ones
    float4* o = reinterpret_cast<float4*>(a.feat_c + (size_t)(ps.x + rank) * kFeatStride);
    o[0] = make_float4((float)ft[0], (float)ft[1], (float)ft[2], (float)ft[3]);
    o[1] = make_float4((float)ft[4], (float)ft[5], (float)ft[6], (float)ft[7]);
    o[2] = make_float4((float)ft[8], 0.f, 0.f, 0.f);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_bin(const K1<T> a) {
  const SweepParams& sw = a.sw;
  if (blockIdx.x == 0 && (int)threadIdx.x < sw.n_sweeps && sw.off[threadIdx.x + 1] == sw.off[threadIdx.x])
    a.num_pillars[threadIdx.x] = 0;                  // a sweep without points has no tile to report its count
  st_bin(a, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

// One block per 1024-point tile.  Tiles are taken in the order the blocks start (ticket), so the tiles a block
// waits for in st_assign belong to blocks that are already running, and those publish their count before they
// wait for anything themselves.
template <typename T>
__global__ void __launch_bounds__(kTile) k_assign(const K1<T> a) {
  __shared__ int s_t;
  if (threadIdx.x == 0) s_t = atomicAdd(a.tile_ticket, 1);
  __syncthreads();
  const int t = s_t;
  int b, il, cell;
  const bool flag = first_touch_flag(a, t, b, il, cell);    // one random read per point, shared by both stages
  // the cell's point count is needed once the slot is known: requested now, under the count exchange between tiles
  const size_t ci = flag ? (size_t)b * a.g.ncell + cell : 0;
  const int cnt_cell = flag ? a.cell_count[ci] : 0;
  st_tilecount(a, t, flag);
  st_assign(a, t, flag, b, cell, ci, cnt_cell);
}

template <typename T>
__global__ void __launch_bounds__(256) k_scatter(const K1<T> a) {
  st_scatter(a, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

// blocks [0, long_blocks): one block per long pillar; the others: one thread per point; all: big pillars
template <typename T>
__global__ void __launch_bounds__(kRankBlock) k_rank(const K1<T> a, int long_blocks) {
  if ((int)blockIdx.x < long_blocks)
    st_rank_long(a, long_blocks);
  else
    st_rank(a, (long long)(blockIdx.x - long_blocks) * blockDim.x + threadIdx.x, (long long)(gridDim.x - long_blocks) * blockDim.x);
  st_rank_big(a);
}

// blocks [0, long_blocks): long pillars, dealt block-minor so that consecutive entries go to different SMs (the
// chains of one SM share its fp64 pipe); they come first in the grid and start with the launch.  The other blocks:
// one thread per pillar slot.
template <typename T>
__global__ void __launch_bounds__(kMeanBlock) k_mean(const K1<T> a, int long_blocks) {
  if ((int)blockIdx.x < long_blocks)
    st_mean_long(a, (int)blockIdx.x + (int)(threadIdx.x >> 5) * long_blocks, (kMeanBlock / 32) * long_blocks);
  else
    st_mean_slots(a, (long long)(blockIdx.x - long_blocks) * kMeanBlock + threadIdx.x);
}

template <typename T>
__global__ void __launch_bounds__(256) k_feat(const K1<T> a) {
  st_feat(a, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
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
      const float* fc = (n0 < c) ? feat_c + (size_t)(sw.off[b] + pil_off[pi] + n0) * kFeatStride + d0 : nullptr;
#pragma unroll
      for (int d = 0; d < kFeatGroup; ++d) {
        vec_t v;
        float* vf = reinterpret_cast<float*>(&v);
        const float* mf = reinterpret_cast<const float*>(&m[d]);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float f = (fc != nullptr && n0 + k < c) ? fc[k * kFeatStride + d] : 0.f;
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
    point_features(x, y, z, r, cx, cy, pil_mean + (size_t)slot * 4, ft);
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
  int *pil_cnt, *pil_off, *pil_cell, *big_list, *long_list, *num_pillars_scratch;
  int2* pil_oc;
  int4* point_seg;
  int* map;                // zero block
  double* pil_mean;
  float* feat_c;
  // zero-initialised block (one memset), starts at cell_count
  int *cell_count, *pil_cursor, *list_cursor, *big_count, *long_count, *tile_ticket, *n_rows;
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
static void layout(A& a, PillarWs* ws, int B, long long T, long long ntiles, int ncell, int P, long long map_cells = 0) {
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
  TAKE(long_count, int, 1);
  TAKE(n_rows, int, 1);
  TAKE(tile_ticket, int, 1);
  TAKE(num_pillars_scratch, int, PP_MAX_SWEEPS);
  TAKE(tile_count, int, (size_t)ntiles + 1);
  TAKE(cell_first, int, nc);
  TAKE(map, int, (size_t)B * (size_t)map_cells);
  if (ws) { ws->zero_bytes = a.used - z0; if (map_cells == 0) ws->map = nullptr; }
  // --- rest
  TAKE(cell_slot, int, nc);
  TAKE(cell_of_point, int, t);
  TAKE(rank_of_point, int, t);
  TAKE(list_u, int, t);
  TAKE(terms, double4, t);
  TAKE(pil_cnt, int, np);
  TAKE(pil_off, int, np);
  TAKE(pil_cell, int, np);
  TAKE(pil_oc, int2, np);
  TAKE(point_seg, int4, t);
  TAKE(big_list, int, t / kBig + 2);
  TAKE(long_list, int, t / kMeanLong + 2);
  TAKE(pil_mean, double, np * 4);
  TAKE(feat_c, float, t * kFeatStride);
#undef TAKE
}

struct SizeArena {
  size_t used = 0;
  template <class T>
  T* take(size_t count) { used += align_up(count * sizeof(T)); return nullptr; }
};

template <typename T>
static int run_stages(const T* pts, long long sp, long long sc, bool vec4, const SweepParams& sw,
                      const GridDev& g, int P, int N, PillarWs& ws, int* d_num_pillars,
                      long long* d_indices, float* d_feat, int* d_status, cudaStream_t st, int map_h = 0, int map_w = 0) {
  const long long total = sw.off[sw.n_sweeps];
  const int ntiles = sw.tile_start[sw.n_sweeps];
  PP_CUDA(cudaMemsetAsync(ws.cell_count, 0, ws.zero_bytes, st));
  K1<T> a;
  a.pts = pts; a.sp = sp; a.sc = sc; a.vec4 = vec4; a.sw = sw; a.g = g; a.P = P; a.N = N;
  a.cell_of_point = ws.cell_of_point; a.cell_first = ws.cell_first; a.cell_count = ws.cell_count;
  a.tile_count = ws.tile_count; a.cell_slot = ws.cell_slot; a.pil_cnt = ws.pil_cnt; a.pil_off = ws.pil_off;
  a.pil_cell = ws.pil_cell; a.pil_cursor = ws.pil_cursor; a.list_cursor = ws.list_cursor; a.list_u = ws.list_u;
  a.rank_of_point = ws.rank_of_point; a.big_count = ws.big_count; a.big_list = ws.big_list;
  a.long_count = ws.long_count; a.long_list = ws.long_list; a.tile_ticket = ws.tile_ticket;
  a.pil_oc = ws.pil_oc; a.point_seg = ws.point_seg; a.terms = ws.terms; a.pil_mean = ws.pil_mean;
  a.indices = d_indices; a.num_pillars = d_num_pillars; a.feat_c = d_feat; a.status = d_status;
  a.map = (map_h > 0 && map_w > 0) ? ws.map : nullptr; a.map_h = map_h; a.map_w = map_w;
  const int pt_blocks = (int)((total + 255) / 256);
  PP_KERNEL("k_bin", st, k_bin<T><<<pt_blocks > 0 ? pt_blocks : 1, 256, 0, st>>>(a));
  if (total > 0) {
    PP_KERNEL("k_assign", st, k_assign<T><<<ntiles, kTile, 0, st>>>(a));
    PP_KERNEL("k_scatter", st, k_scatter<T><<<pt_blocks, 256, 0, st>>>(a));
    PP_KERNEL("k_rank", st, k_rank<T><<<2 * sm_count() + pt_blocks, kRankBlock, 0, st>>>(a, 2 * sm_count()));
  }
  const long long groups = (long long)sw.n_sweeps * P;
  const int long_blocks = total > 0 ? 2 * sm_count() : 0;   // 16 chains per SM keep its fp64 pipe busy without queueing
  PP_KERNEL("k_mean", st, k_mean<T><<<long_blocks + (int)((groups + kMeanBlock - 1) / kMeanBlock), kMeanBlock, 0, st>>>(a, long_blocks));
  if (d_feat != nullptr && total > 0) PP_KERNEL("k_feat", st, k_feat<T><<<pt_blocks, 256, 0, st>>>(a));
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
  rc = run_stages<T>(pts, sp, sc, vec4, sw, g, P, N, ws, d_num_pillars, (long long*)d_indices, ws.feat_c,
                     d_status, st);
  if (rc != PP_OK) return rc;
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
  layout(arena, &ws, B, sw.off[B], sw.tile_start[B], g.ncell, P, (long long)H * W);
  if (!arena.ok) return PP_ERR_WORKSPACE;
  const size_t k1_bytes = arena.used;
  if (stages & 1) {
    const bool vec4 = sizeof(T) == 4 && sp == 4 && sc == 1 && ((uintptr_t)pts % 16) == 0;
    rc = run_stages<T>(pts, sp, sc, vec4, sw, g, P, N, ws, d_num_pillars, (long long*)d_indices, ws.feat_c, d_status, st, H, W);
    if (rc != PP_OK) return rc;
    if (d_x != nullptr) {
      rc = emit_dense(sw, P, N, d_mean, d_x, d_num_pillars, ws, st);
      if (rc != PP_OK) return rc;
    }
  }
  if (!(stages & 2)) return PP_OK;
  CompactPillars cp{};
  cp.sw = sw; cp.P = P; cp.N = N;
  cp.feat_c = ws.feat_c; cp.pil_cnt = ws.pil_cnt; cp.pil_off = ws.pil_off; cp.cell_map = ws.map;
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
  rc = run_stages<T>(pts, sp, sc, vec4, sw, g, P, N, ws, ws.num_pillars_scratch, nullptr, nullptr, d_status, st);
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

static size_t k1_workspace_bytes(int32_t n_sweeps, int64_t total_points, const pp_grid* grid, int32_t max_pillars,
                                 long long map_cells) {
  pp::GridDev g;
  if (!pp::make_grid(grid, g) || n_sweeps < 1 || n_sweeps > PP_MAX_SWEEPS || total_points < 0 ||
      max_pillars < 1)
    return 0;
  pp::SizeArena a;
  const long long ntiles = total_points / pp::kTile + n_sweeps + 1;
  pp::layout(a, (pp::PillarWs*)nullptr, n_sweeps, total_points, ntiles, g.ncell, max_pillars, map_cells);
  return a.used + pp::kAlign;
}

size_t pp_pillarize_workspace_bytes(int32_t n_sweeps, int64_t total_points, const pp_grid* grid,
                                    int32_t max_pillars) {
  return k1_workspace_bytes(n_sweeps, total_points, grid, max_pillars, 0);
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
  if (C < 1 || canvas_h < 1 || canvas_w < 1 || max_points_per_pillar < 1) return 0;
  const size_t k1 = k1_workspace_bytes(n_sweeps, total_points, grid, max_pillars, (long long)canvas_h * canvas_w);
  if (k1 == 0) return 0;
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
  layout(arena, &ws, n_sweeps, sw.off[n_sweeps], sw.tile_start[n_sweeps], g.ncell, P, (long long)canvas_h * canvas_w);
  if (!arena.ok) return PP_ERR_WORKSPACE;
  CompactPillars cp{};
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
