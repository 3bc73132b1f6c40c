// internal.cuh -- types shared between the K1 (pillarize.cu) and K2 (pfn*.cu) translation units for the
// fused input path (pp_input_path): K1's compact per-point state is consumed directly, the dense
// [B,9,P,N] network input is never materialised.
#pragma once
#include "common.cuh"

namespace pp {

constexpr int kFeatStride = 12;   // floats per feat_c record: 9 features padded to three aligned float4

struct SweepParams {
  int n_sweeps;
  int tile_start[PP_MAX_SWEEPS + 1];
  long long off[PP_MAX_SWEEPS + 1];
};

// K1 state handed to the sparse PFN (all device pointers, valid until the workspace is reused)
struct CompactPillars {
  SweepParams sw;            // point offsets of the sweeps (host copy, passed by value to kernels)
  int P, N;
  const int* cell_map;       // [B, H*W]: slot + 1 of the pillar on each canvas cell, 0 = none (written by K1's mean stage
                             // when the canvas size is known; saves the index -> map kernel and its memset), or null
  const float* feat_c;       // [T, kFeatStride] float (9 used): decorated features of the kept points, before "- data_mean",
                             //   at sw.off[b] + pil_off[b*P+p] + rank
  const int* pil_cnt;        // [B*P] in-range points of the pillar (may exceed N)
  const int* pil_off;        // [B*P] first entry of the pillar's segment (relative to the sweep)
  const int* num_pillars;    // [B]
  const float* data_mean;    // [9*P*N] or nullptr
  const void* mean_prepared; // pp_mean_prepare's output for data_mean, or nullptr (prepared per call in the workspace)
};

struct PfnParams {
  const float *conv_w, *conv_b, *bn_w, *bn_b;
  float *running_mean, *running_var;
  int64_t* num_batches_tracked;
  int training;
  float momentum, eps;
};

size_t pfn_sparse_workspace_bytes(int B, int P, int N, int C, int H, int W, bool own_prep);
size_t mean_prepared_bytes(int P, int N);
int mean_prepare(const float* d_mean, int P, int N, void* d_prep, size_t bytes, cudaStream_t st);
// canvas [B,C,H,W] from K1's compact state; d_inds is K1's [B,P,3] indices output
int pfn_sparse_scatter(const CompactPillars& cp, const int64_t* d_inds, int C, const PfnParams& prm, int H, int W,
                       float* d_canvas, int32_t* d_status, void* d_ws, size_t ws_bytes, cudaStream_t st);
bool pfn_sparse_supported(int B, int P, int N, int C, const void* data_mean);

// backward of the same path (pfn_bwd.cu): conv1 / bn1 gradients from the compact state and the canvas gradient
size_t pfn_sparse_backward_workspace_bytes(int B, int P);
int pfn_sparse_backward(const CompactPillars& cp, const int64_t* d_inds, int C, const float* conv_w, const float* conv_b,
                        const float* bn_w, const float* running_mean, const float* running_var, int training, float eps,
                        int H, int W, const float* d_grad_canvas, float* g_w, float* g_b, float* g_gamma, float* g_beta,
                        void* d_ws, size_t ws_bytes, cudaStream_t st);

}  // namespace pp
