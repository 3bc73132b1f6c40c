// aggregate.cu -- multi-sweep aggregation in front of K1: data/dataset.py:54-88 with the
// Lyft SDK's LidarPointCloud.transform / remove_close, on the raw [n,5] float32 rows that are already in HBM.
//
// In place: every point's x,y,z become float32(M_file[:3,:] . [x,y,z,1]) (float64 product, sequential order,
// one rounding to float32 -- the SDK stores the float64 product back into its float32 array), and a point
// that remove_close drops (|x| < r and |y| < r in float32) is overwritten with a finite out-of-range sentinel,
// which K1's range filter (data/pillars.cpp:271-276) rejects.  The surviving points keep their order, so
// pillar membership, in-pillar order and slot order equal those of the compacted cloud, and the per-sample
// offsets stay the host-known file boundaries: no compaction pass, no device-to-host count.
#include "common.cuh"

namespace pp {

constexpr float kDroppedCoord = 3.0e38f;
constexpr int kAggMaxFilesSmem = 1024;

__global__ void __launch_bounds__(256) k_aggregate(float* __restrict__ pts, long long T, int S,
                                                   const long long* __restrict__ file_offsets, int F,
                                                   const double* __restrict__ xf, float radius, int* __restrict__ kept) {
  __shared__ long long s_off[kAggMaxFilesSmem + 1];
  const bool in_smem = F <= kAggMaxFilesSmem;
  if (in_smem) {
    for (int i = threadIdx.x; i <= F; i += 256) s_off[i] = file_offsets[i];
    __syncthreads();
  }
  const long long* off = in_smem ? s_off : file_offsets;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < T; i += (long long)gridDim.x * 256) {
    int lo = 0, hi = F;                                            // file f: off[f] <= i < off[f+1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (off[mid] <= i) lo = mid; else hi = mid;
    }
    const double* M = xf + (size_t)lo * 12;
    float* p = pts + (size_t)i * S;
    const double x = (double)p[0], y = (double)p[1], z = (double)p[2];
    float o[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(__ldg(M + 4 * r), x), __dmul_rn(__ldg(M + 4 * r + 1), y)),
                                           __dmul_rn(__ldg(M + 4 * r + 2), z)),
                                 __ldg(M + 4 * r + 3));
      o[r] = __double2float_rn(v);
    }
    const bool close = fabsf(o[0]) < radius && fabsf(o[1]) < radius;          // remove_close
    if (close) {
      o[0] = o[1] = o[2] = kDroppedCoord;
    } else if (kept != nullptr) {
      atomicAdd(kept + lo, 1);
    }
    p[0] = o[0]; p[1] = o[1]; p[2] = o[2];
  }
}

}  // namespace pp

extern "C" {

int pp_aggregate_sweeps(float* d_points, int64_t n_points, int32_t point_stride, const int64_t* d_file_offsets,
                        int32_t n_files, const double* d_transforms, float min_dist, int32_t* d_kept,
                        pp_stream_t stream) {
  using namespace pp;
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_file_offsets || !d_transforms || n_points < 0 || point_stride < 3 || n_files < 1 || (!d_points && n_points > 0))
    return PP_ERR_INVALID_ARG;
  if (d_kept != nullptr) PP_CUDA(cudaMemsetAsync(d_kept, 0, (size_t)n_files * sizeof(int32_t), st));
  if (n_points == 0) return PP_OK;
  const long long want = (n_points + 255) / 256;
  const int blocks = (int)(want < (long long)sm_count() * 8 ? want : (long long)sm_count() * 8);
  PP_KERNEL("k_aggregate", st,
            (k_aggregate<<<blocks, 256, 0, st>>>(d_points, (long long)n_points, point_stride,
                                                 (const long long*)d_file_offsets, n_files, d_transforms, min_dist, d_kept)));
  return PP_OK;
}

}  // extern "C"
