/*
 * pp_b200.h -- C ABI of libpp_b200.so: the B200 (sm_100a) PointPillars input path.
 *
 * This is the drop-in boundary for the per-sweep hot path of mr3543/3d-Object-Detection
 * (paths below are under the reference checkout):
 *
 *   pp_pillarize          replaces  data/pillars.cpp:236-398  create_pillars  (pybind11 export :433)
 *                         + the tensor glue of                data/dataset.py:99-106
 *   pp_aggregate_sweeps   replaces  data/dataset.py:54-88     sweep aggregation (SDK transform + remove_close)
 *   pp_pfn_forward        replaces  model/model.py:31-40      PPFeatureNet.forward
 *   pp_scatter            replaces  model/model.py:53-62      PPScatter.forward
 *   pp_pfn_scatter        the two above fused (canvas written straight from the pillar maxima)
 *   pp_make_ious          replaces  data/pillars.cpp:400-427  make_ious       (pybind11 export :432)
 *   pp_assign_targets     replaces  utils/box_utils.py:162-232 create_target + :70-109 make_target
 *
 * Conventions
 *   - Every pointer named d_* is DEVICE memory of the current CUDA device; h_* is host memory.
 *     No torch / pybind types appear here: plain pointers, sizes and a stream handle.
 *   - Every function returns PP_OK (0) or a PP_ERR_* code; nothing throws, nothing calls exit().
 *     (The reference kills the process on a negative IoU, data/pillars.cpp:166-169; here the
 *     kernels raise bit PP_STATUS_NEG_IOU in the caller's d_status word instead.)
 *   - All work is enqueued on `stream` and is asynchronous with respect to the host.  No hidden
 *     allocation, no global state: scratch memory is a caller-provided workspace whose size the
 *     matching *_workspace_bytes() function reports.  Calls are re-entrant per (device, stream,
 *     workspace).
 *   - There is no CPU fallback.  If no CUDA device is usable the calls return PP_ERR_CUDA.
 */
#ifndef PP_B200_H_
#define PP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_B200_VERSION 100 /* major*100 + minor */

typedef void* pp_stream_t; /* a cudaStream_t */

enum {
  PP_OK = 0,
  PP_ERR_INVALID_ARG = 1, /* bad size / NULL pointer / unsupported combination */
  PP_ERR_WORKSPACE = 2,   /* workspace too small or misaligned (needs 256-byte alignment) */
  PP_ERR_CUDA = 3,        /* a CUDA runtime call failed; see pp_last_cuda_error() */
  PP_ERR_UNSUPPORTED = 4
};

/* bits of the device-side status word (d_status), OR-ed in by kernels */
enum {
  PP_STATUS_NEG_IOU = 1,       /* wrong corner winding: the reference would exit(1) */
  PP_STATUS_BAD_POINT = 2,     /* NaN/Inf coordinate met the reference's range filter (out of contract); point dropped */
  PP_STATUS_BAD_INDEX = 4,     /* scatter index outside the canvas; row skipped */
  PP_STATUS_CAND_OVERFLOW = 8, /* more candidate anchors for one GT than the anchor index promised */
  PP_STATUS_RANGE = 16         /* pp_input_path: a data_mean value or conv weight left the fp16 range of the
                                  tensor-core padding pass (|v| >= 2^15, |256 w| >= 2^15); canvas invalid,
                                  use pp_pillarize + pp_pfn_scatter */
};

enum { PP_F32 = 0, PP_F64 = 1, PP_I64 = 2 };

#define PP_MAX_SWEEPS 64   /* sweeps per call (batch); larger batches are split by the caller */
#define PP_NUM_FEATURES 9  /* x,y,z,r,xp,yp,xc,yc,zc : data/pillars.cpp:48-56 */
#define PP_REG_DIMS 9      /* [1,dx,dy,dz,dw,dl,dh,sin(dyaw),ort] : utils/box_utils.py:109 */

/* The nine doubles create_pillars receives positionally (data/pillars.cpp:241-249). */
typedef struct pp_grid {
  double x_step, y_step;
  double x_min, y_min, z_min;
  double x_max, y_max, z_max;
  double canvas_height;
} pp_grid;

int pp_version(void);
const char* pp_error_string(int code);
/* cudaError_t of the most recent failing CUDA call made by this thread inside the library. */
int pp_last_cuda_error(void);

/* Fused input path: pp_pillarize's stages, then PPFeatureNet + PPScatter evaluated straight from the
 * compact per-point state -- the dense network input x [B,9,P,N] (data/dataset.py:99-105) is never
 * materialised unless d_x is non-NULL.  Same canvas as pp_pillarize + pp_pfn_scatter up to fp32
 * summation order (model/model.py:31-40,53-62; BatchNorm statistics over all B*P*N slots, padding
 * included).  It uses that a padding slot holds 0 - data_mean[d,p,n] in every sweep: the padding
 * slots are evaluated once per (p,n) on the tensor cores, whatever the batch size (suffix extremes at the
 * slot boundaries 4 / 16 / 48 + BatchNorm sums), the ~1.3 % of slots that hold a point -- and the few
 * padding slots between a pillar's count and the next boundary -- are evaluated separately.
 * Supported: 1 <= n_sweeps <= PP_MAX_SWEEPS, C = 64,
 * max_points_per_pillar in [16, 255] and a multiple of 8, max_pillars even; otherwise PP_ERR_UNSUPPORTED
 * (call pp_pillarize + pp_pfn_scatter).  d_indices [B,P,3] int64 and d_num_pillars [B] int32 are
 * outputs as in pp_pillarize.  A data_mean value or weight outside the fp16 range of the padding
 * pass raises PP_STATUS_RANGE in *d_status.
 * d_mean_prepared: the output of pp_mean_prepare for d_data_mean (pillar_means.pkl is a constant of the
 * dataset, make_means.py: prepare it once).  NULL with a non-NULL d_data_mean: the operand is prepared inside
 * the workspace on every call (ask pp_input_path_workspace_bytes with prepare_in_workspace = 1); correct, but
 * it costs a streaming pass over data_mean per call.
 * stages: PP_STAGE_PILLARIZE | PP_STAGE_ENCODE (3) runs everything.  A streaming caller may issue the
 * two stages separately (same arguments, same workspace) -- e.g. the pillarize stage of batch k+1 on
 * one stream while the encode stage of batch k runs on another; with training != 0 the encode stages
 * of successive batches must be ordered by the caller (they update the running statistics). */
enum { PP_STAGE_PILLARIZE = 1, PP_STAGE_ENCODE = 2 };
size_t pp_input_path_workspace_bytes(int32_t n_sweeps, int64_t total_points, const pp_grid* grid,
                                     int32_t max_points_per_pillar, int32_t max_pillars, int32_t C, int32_t canvas_h,
                                     int32_t canvas_w, int32_t prepare_in_workspace);
int pp_input_path(const void* d_points, int32_t point_dtype, int64_t stride_point, int64_t stride_col,
                  const int64_t* h_sweep_offsets, int32_t n_sweeps, const pp_grid* grid,
                  int32_t max_points_per_pillar, int32_t max_pillars, const float* d_data_mean,
                  const void* d_mean_prepared, int32_t C,
                  const float* d_conv_w, const float* d_conv_b, const float* d_bn_w, const float* d_bn_b,
                  float* d_running_mean, float* d_running_var, int64_t* d_num_batches_tracked,
                  int32_t training, float momentum, float eps, int32_t canvas_h, int32_t canvas_w,
                  float* d_canvas, float* d_x, int64_t* d_indices, int32_t* d_num_pillars,
                  int32_t* d_status, void* d_workspace, size_t workspace_bytes, int32_t stages,
                  pp_stream_t stream);

/* The per-slot normalisation constant of data/dataset.py:99-105 (pillar_means.pkl, make_means.py:28-37) in the
 * form the padding pass of pp_input_path consumes: x = 0 - data_mean[d,p,n] split into two fp16 pieces and laid
 * out as the tensor-core operand ([P][3][N][8] halves, 48 bytes per slot), plus the first and second moments of x
 * over all (p,n) in float64 (they turn the padding slots' BatchNorm sums into sums of |y| and y|y|) and a range
 * flag (|x| >= 2^15).  d_prepared: pp_mean_prepared_bytes() bytes, 256-byte aligned, owned by the caller; valid
 * for this data_mean, max_pillars and max_points_per_pillar until overwritten.  Deterministic. */
size_t pp_mean_prepared_bytes(int32_t max_pillars, int32_t max_points_per_pillar);
int pp_mean_prepare(const float* d_data_mean, int32_t max_pillars, int32_t max_points_per_pillar, void* d_prepared,
                    size_t prepared_bytes, pp_stream_t stream);

/* Backward of pp_input_path with respect to conv1 / bn1 (the fused path made trainable): the same gradients
 * pp_pfn_backward produces from the dense x, computed from the forward's compact per-point state -- x is never
 * materialised.  Training-mode BatchNorm makes the weight gradient dense over all B*P*N slots; the padding slots
 * hold 0 - data_mean[d,p,n] in every sweep, so their moment sums are taken once over [P,N] (k_pfn_bwd_pad) and
 * multiplied by n_sweeps, and only the live pillars are visited per sweep (k_pfn_bwd_live: real-minus-padding
 * corrections and the arg-max routing of the canvas gradient).
 *   d_forward_workspace: the workspace pp_input_path ran in (all stages), NOT reused since; h_sweep_offsets,
 *   grid, N, P, d_data_mean as in that call; d_indices / d_num_pillars its outputs; d_grad_canvas [B,C,H,W].
 *   training == 0: running statistics normalise (no padding pass needed).
 * Outputs (any may be NULL): d_grad_conv_w [C,9], d_grad_conv_b, d_grad_bn_w, d_grad_bn_b [C]. */
size_t pp_input_path_backward_workspace_bytes(int32_t n_sweeps, int32_t max_pillars, int32_t C);
int pp_input_path_backward(const int64_t* h_sweep_offsets, int32_t n_sweeps, const pp_grid* grid,
                           int32_t max_points_per_pillar, int32_t max_pillars, const float* d_data_mean, int32_t C,
                           const float* d_conv_w, const float* d_conv_b, const float* d_bn_w,
                           const float* d_running_mean, const float* d_running_var, int32_t training, float eps,
                           int32_t canvas_h, int32_t canvas_w, const float* d_grad_canvas, const int64_t* d_indices,
                           const int32_t* d_num_pillars, float* d_grad_conv_w, float* d_grad_conv_b,
                           float* d_grad_bn_w, float* d_grad_bn_b, const void* d_forward_workspace,
                           size_t forward_workspace_bytes, void* d_workspace, size_t workspace_bytes,
                           pp_stream_t stream);

/* Multi-sweep aggregation in front of pp_pillarize (data/dataset.py:54-88 with the Lyft SDK's
 * LidarPointCloud.transform and remove_close), IN PLACE on raw lidar rows:
 *   d_points [n_points, point_stride >= 3] float32 (Lyft .bin rows x,y,z,intensity,ring); file f owns rows
 *   d_file_offsets[f] .. d_file_offsets[f+1] (int64, n_files + 1 entries, device);
 *   d_transforms [n_files, 3, 4] float64 row-major = the first three rows of dataset.py:78's transmat.
 * x,y,z <- float32(M . [x,y,z,1]) (float64 product, one rounding); a point with |x| < min_dist and
 * |y| < min_dist afterwards (float32 compare, remove_close) gets the finite out-of-range sentinel 3.0e38 in
 * x,y,z, which pp_pillarize's range filter drops: the surviving points keep their order, so pillars are those
 * of the compacted cloud and sample offsets stay the file boundaries.  d_kept [n_files] int32 (may be NULL)
 * receives the number of points each file keeps. */
int pp_aggregate_sweeps(float* d_points, int64_t n_points, int32_t point_stride, const int64_t* d_file_offsets,
                        int32_t n_files, const double* d_transforms, float min_dist, int32_t* d_kept,
                        pp_stream_t stream);

/* Backward of PPFeatureNet: the gradients torch autograd computes through
 * model/model.py:36-39 (conv1 -> relu -> bn1 -> max over N) for the four parameter tensors, from the same
 * inputs the forward took; nothing saved by the forward is needed (batch statistics are recomputed).
 *   d_grad_out: dL/d(output) [B, C, P] when d_inds is NULL; otherwise the CANVAS gradient [B, C, H, W] and
 *   d_inds [B, P, 3] int64 -- the PPScatter backward (model/model.py:61) is folded in: rows with
 *   inds[b,p,0] != 0 read the gradient at (inds[b,p,2], inds[b,p,1]), the others get zero.
 *   training != 0: BatchNorm batch statistics (the mean/variance terms of the BN backward are included);
 *   training == 0: d_running_mean / d_running_var normalise, as in the forward.
 * Outputs (any may be NULL): d_grad_conv_w [C, D], d_grad_conv_b [C], d_grad_bn_w [C], d_grad_bn_b [C].
 * The gradient with respect to x is not produced (x is input data in train.py).  D == 9, C == 64. */
size_t pp_pfn_backward_workspace_bytes(int32_t B, int32_t P, int32_t C);
int pp_pfn_backward(const float* d_x, int32_t B, int32_t D, int32_t P, int32_t N, int32_t C, const float* d_conv_w,
                    const float* d_conv_b, const float* d_bn_w, const float* d_running_mean,
                    const float* d_running_var, int32_t training, float eps, const float* d_grad_out,
                    const int64_t* d_inds, int32_t canvas_h, int32_t canvas_w, float* d_grad_conv_w,
                    float* d_grad_conv_b, float* d_grad_bn_w, float* d_grad_bn_b, void* d_workspace,
                    size_t workspace_bytes, pp_stream_t stream);

/* Backward of PPScatter alone (model/model.py:61): d_grad_feat [B, C, P] = canvas gradient at each non-empty
 * row's cell (torch's index_put derivative: every indexed row reads its cell), zero for rows with flag 0. */
int pp_scatter_backward(const float* d_grad_canvas, const int64_t* d_inds, int32_t B, int32_t C, int32_t P,
                        int32_t canvas_h, int32_t canvas_w, float* d_grad_feat, pp_stream_t stream);

/* Loss front-end: forward of the reference's PPLoss (model/loss.py:24-63) and the gradient
 * of its total loss b_cls*cls + b_reg*reg + b_ort*ort with respect to both network outputs.
 *   d_cls_out [B, Ad*K, H, W] float NCHW logits; d_reg_out [B, Ad*R, H, W] float, MODIFIED IN PLACE like
 *   the reference does: tanh on network channel 6 (model/loss.py:50 indexes the permuted view's last axis);
 *   d_cls_t [B, A, K], d_reg_t [B, A, 9] float targets (pp_assign_targets layout), A = H*W*Ad.
 * Outputs: d_scores [B, A*K] = sigmoid(logits) in target order (may be NULL), d_grad_cls / d_grad_reg in the
 * layout of the inputs (may be NULL), d_losses float[4] = {cls_loss, reg_loss, ort_loss, total}.
 * focal weight (t == 1 ? alpha_pos : 1) * (1 - pt)^gamma is detached as in the reference (alpha_pos = 25 there).
 * No positive anchor -> reg/ort/total are NaN (torch's mean over an empty tensor).  Ad*K <= 96, R >= 7. */
size_t pp_loss_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t anchors_per_cell);
int pp_loss(const float* d_cls_out, float* d_reg_out, const float* d_cls_t, const float* d_reg_t, int32_t B,
            int32_t H, int32_t W, int32_t anchors_per_cell, int32_t num_classes, int32_t reg_dims, float gamma,
            float alpha_pos, float b_cls, float b_reg, float b_ort, float* d_scores, float* d_grad_cls,
            float* d_grad_reg, float* d_losses, void* d_workspace, size_t workspace_bytes, pp_stream_t stream);

/* pp_loss fed by the positives list of pp_assign_targets_list instead of the dense target tensors: the 78 MB
 * classification-target read and the 78 MB regression-target scan disappear (targets are zero except at the
 * listed anchors).  Same outputs as pp_loss on the densified list.  Needs H*W % 4 == 0 and 16-byte aligned
 * tensors (the TMA kernel); otherwise PP_ERR_UNSUPPORTED -- densify and call pp_loss. */
int pp_loss_list(const float* d_cls_out, float* d_reg_out, const int32_t* d_pos_anchor, const float* d_pos_cls,
                 const float* d_pos_reg, const int32_t* d_pos_offsets, int32_t B, int32_t H, int32_t W,
                 int32_t anchors_per_cell, int32_t num_classes, int32_t reg_dims, float gamma, float alpha_pos,
                 float b_cls, float b_reg, float b_ort, float* d_scores, float* d_grad_cls, float* d_grad_reg,
                 float* d_losses, void* d_workspace, size_t workspace_bytes, pp_stream_t stream);

/* Chain rule for a non-unit upstream gradient of the total loss: both gradient tensors are multiplied in place
 * by *d_scale / *d_applied (d_applied NULL = 1).  The kernel returns at once when the factor is exactly 1, the
 * usual total_loss.backward() case, so no pass over the 147 MB of gradients is spent on it. */
int pp_loss_scale_grads(float* d_grad_cls, size_t n_cls, float* d_grad_reg, size_t n_reg, const float* d_scale,
                        const float* d_applied, pp_stream_t stream);

/* Instrumentation.  pp_launch_count: kernels launched by this library since load (all threads).
 * pp_profile_enable(1): bracket every kernel launch with CUDA events on its stream;
 * pp_profile_report: synchronise the device, write one "name launches total_ms" line per kernel
 * into buf, clear the records, return the bytes needed. */
int64_t pp_launch_count(void);
/* Process-wide options.  "pfn_tensor_cores": 1 (default) runs the dense PFN statistics pass on tcgen05
 * tensor cores (fp16 3-term split, TF32 split as the guarded fallback) whenever D=9, C=64, N%8==0, N<=256;
 * 2 forces the TF32 kernel, 0 the CUDA-core kernel.  "loss_tma": 1 (default) uses the tensor-map TMA kernel for the
 * classification loss where the shape allows, 0 the generic tile kernel.  Unknown keys: PP_ERR_INVALID_ARG.
 * (Development knobs -- role timing, stage ablation -- are not part of this header: include/pp_b200_debug.h,
 * present only in a library built with -DPP_DEBUG.) */
int pp_set_option(const char* key, int value);
int pp_profile_enable(int on);
int64_t pp_profile_report(char* buf, int64_t buf_bytes);

/* ------------------------------------------------------------------------------------------
 * K1  point -> pillar binning, order-exact compaction, decoration, per-slot mean subtraction
 * ----------------------------------------------------------------------------------------
 * Points of all sweeps are concatenated: element (i, c), c in 0..3 = x,y,z,r, lives at
 * d_points[(i*stride_point + c*stride_col)] in units of elements of `point_dtype`
 * (PP_F32 or PP_F64).  Sweep s owns rows h_sweep_offsets[s] .. h_sweep_offsets[s+1]-1.
 * Semantics per sweep are exactly SURVEY.md App. A.3 / data/pillars.cpp:236-398:
 * half-open range filter, floor binning in double, pillars in first-touch order, the first
 * max_pillars kept, the first max_points_per_pillar points of each kept in input order, the
 * pillar mean taken over ALL its in-range points with the reference's sequential update.
 */
size_t pp_pillarize_workspace_bytes(int32_t n_sweeps, int64_t total_points, const pp_grid* grid,
                                    int32_t max_pillars);

/*
 * Dense output (the tensor PPDataset.__getitem__ hands to the network, data/dataset.py:99-106):
 *   d_x        float [n_sweeps, 9, max_pillars, max_points]  = float32(feature) - data_mean
 *              (every slot written; empty slots hold 0 - data_mean, exactly like the reference's
 *              flat subtraction).  d_data_mean is [9*max_pillars*max_points] floats or NULL.
 *   d_indices  int64 [n_sweeps, max_pillars, 3] = [1, canvas_x, canvas_y] for kept pillars, 0 rows after.
 *   d_num_pillars int32 [n_sweeps].
 */
int pp_pillarize(const void* d_points, int32_t point_dtype, int64_t stride_point,
                 int64_t stride_col, const int64_t* h_sweep_offsets, int32_t n_sweeps,
                 const pp_grid* grid, int32_t max_points_per_pillar, int32_t max_pillars,
                 const float* d_data_mean, float* d_x, int64_t* d_indices, int32_t* d_num_pillars,
                 int32_t* d_status, void* d_workspace, size_t workspace_bytes, pp_stream_t stream);

/*
 * Compact double-precision output for the numpy-signature drop-in of create_pillars, whose
 * contract is "only touched slots of the caller's [P,N,9] float64 array are written"
 * (data/pillars.cpp:48-56,390-392).  One sweep per call.
 *   d_rows   double [n_points, 9]  features of the i-th kept point in (pillar, rank) order;
 *   d_slot   int32  [n_points]     pillar*max_points + rank, or -1 for rows past the kept count;
 *   d_pillar_xy int32 [max_pillars, 2] canvas_x, canvas_y of each kept pillar;
 *   d_counts int32 [2]             {number of kept pillars, number of valid rows}.
 */
int pp_pillarize_compact(const void* d_points, int32_t point_dtype, int64_t stride_point,
                         int64_t stride_col, int64_t n_points, const pp_grid* grid,
                         int32_t max_points_per_pillar, int32_t max_pillars, double* d_rows,
                         int32_t* d_slot, int32_t* d_pillar_xy, int32_t* d_counts,
                         int32_t* d_status, void* d_workspace, size_t workspace_bytes,
                         pp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2  PPFeatureNet (1x1 conv D->C, ReLU, BatchNorm, max over N) and PPScatter
 * ----------------------------------------------------------------------------------------
 * d_x [B, D, P, N] float (D == 9).  Parameters keep nn.Conv2d / nn.BatchNorm2d layouts:
 * conv weight [C, D] (= conv1.weight[:, :, 0, 0]), conv bias [C], bn weight/bias/running_mean/
 * running_var [C] (C <= 64, multiple of 32), num_batches_tracked int64[1].
 * training != 0: batch statistics over (B,P,N) normalise the output and the running statistics
 * are updated in place (momentum, unbiased variance) exactly like nn.BatchNorm2d.train();
 * training == 0: running statistics are used and left untouched.
 */
size_t pp_pfn_workspace_bytes(int32_t B, int32_t P, int32_t C, int32_t canvas_h, int32_t canvas_w);

/* out [B, C, P] float : the tensor PPFeatureNet.forward returns (model/model.py:39-40). */
int pp_pfn_forward(const float* d_x, int32_t B, int32_t D, int32_t P, int32_t N, int32_t C,
                   const float* d_conv_w, const float* d_conv_b, const float* d_bn_w,
                   const float* d_bn_b, float* d_running_mean, float* d_running_var,
                   int64_t* d_num_batches_tracked, int32_t training, float momentum, float eps,
                   float* d_out, void* d_workspace, size_t workspace_bytes, pp_stream_t stream);

/* canvas [B, C, H, W] float, fully written: zeros + feat[b,:,p] at (inds[b,p,2], inds[b,p,1]) for
 * rows with inds[b,p,0] != 0 (model/model.py:55-61).  d_feat is [B, C, P]. */
int pp_scatter(const float* d_feat, const int64_t* d_inds, int32_t B, int32_t C, int32_t P,
               int32_t canvas_h, int32_t canvas_w, float* d_canvas, int32_t* d_status,
               void* d_workspace, size_t workspace_bytes, pp_stream_t stream);

/* pp_pfn_forward + pp_scatter in one pass over x; d_out ([B,C,P]) may be NULL. */
int pp_pfn_scatter(const float* d_x, const int64_t* d_inds, int32_t B, int32_t D, int32_t P,
                   int32_t N, int32_t C, const float* d_conv_w, const float* d_conv_b,
                   const float* d_bn_w, const float* d_bn_b, float* d_running_mean,
                   float* d_running_var, int64_t* d_num_batches_tracked, int32_t training,
                   float momentum, float eps, int32_t canvas_h, int32_t canvas_w, float* d_canvas,
                   float* d_out, int32_t* d_status, void* d_workspace, size_t workspace_bytes,
                   pp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3  rotated anchor-vs-GT IoU and target assignment
 * ----------------------------------------------------------------------------------------
 * Ring conventions are the reference's (data/pillars.cpp:15-16,149-157): anchor corners
 * [A,4,2] counter-clockwise, GT corners [G,4,2] clockwise (they are after the y flip of
 * utils/box_utils.py:29-30), both open rings, doubles.  Centres are [.,3] doubles in the same
 * (image) space.  IoU is 0 when |dcx| > 10 or |dcy| > 10 (data/pillars.cpp:418-419).
 */

/* Dense [A,G] double IoU matrix, every entry written: make_ious verbatim. */
int pp_make_ious(const double* d_a_corners, const double* d_g_corners, const double* d_a_centers,
                 const double* d_g_centers, int64_t A, int64_t G, double* d_ious,
                 int32_t* d_status, pp_stream_t stream);

/* Anchors are batch-invariant (built once offline in the reference, train_prep.py:115-120).
 * The index buckets anchor centres on a uniform grid so that each GT visits only the anchors
 * that can pass the centre prefilter.  Built once per anchor set from HOST centres; with
 * h_a_corners ([A,4,2], non-NULL / with_geometry != 0) the index also keeps a bucket-ordered copy of
 * the centres and corners, which turns the kernels' per-candidate gathers into contiguous reads. */
size_t pp_anchor_index_bytes(const double* h_a_centers, int64_t A, int32_t with_geometry);
int pp_anchor_index_build(const double* h_a_centers, const double* h_a_corners, int64_t A, void* d_index,
                          size_t index_bytes, pp_stream_t stream);

size_t pp_assign_targets_workspace_bytes(int32_t n_sweeps, int64_t A, int64_t total_gt,
                                         const void* h_index_header /* first 64 bytes of the index, host copy; may be NULL for a safe upper bound */);

/*
 * Target assignment for a batch of sweeps (SURVEY.md App. A.7 = utils/box_utils.py:178-232).
 * Anchors: corners/centres as above plus d_a_wlh [A,3], d_a_yaw [A] (radians).
 * GT boxes of all sweeps concatenated; sweep s owns rows h_gt_offsets[s]..h_gt_offsets[s+1]-1:
 *   d_g_corners [Gt,4,2], d_g_centers [Gt,3] in IMAGE space (y already flipped),
 *   d_g_wlh [Gt,3], d_g_yaw [Gt] (radians, NOT flipped), d_g_cls int32 [Gt].
 * Outputs (every element written):
 *   d_cls float [n_sweeps, A, num_classes], d_reg float [n_sweeps, A, 9],
 *   d_top_anchor int32 [Gt] (per-GT best anchor, 0 = dropped like np.nonzero does, :204),
 *   d_counts int32 [n_sweeps, 4] = {positives (max IoU > thresh), forced matches kept,
 *                                   anchors with |max IoU - thresh| < 1e-6, reserved}.
 */
int pp_assign_targets(const double* d_a_corners, const double* d_a_centers, const double* d_a_wlh,
                      const double* d_a_yaw, const void* d_anchor_index, int64_t A,
                      const double* d_g_corners, const double* d_g_centers, const double* d_g_wlh,
                      const double* d_g_yaw, const int32_t* d_g_cls, const int64_t* h_gt_offsets,
                      int32_t n_sweeps, int32_t num_classes, double pos_thresh, float* d_cls,
                      float* d_reg, int32_t* d_top_anchor, int32_t* d_counts, int32_t* d_status,
                      void* d_workspace, size_t workspace_bytes, pp_stream_t stream);

/* The same assignment with the targets as a POSITIVES LIST + implicit zeros instead of, or next
 * to, the two dense [B,A,9] tensors (which are > 99.9 % zeros): every anchor whose cls or reg row is non-zero, in
 * ascending (sweep, anchor) order -- d_pos_anchor[i] = sweep * A + anchor (int32), d_pos_cls [capacity, 9] and
 * d_pos_reg [capacity, 9] its two rows exactly as pp_assign_targets writes them, d_pos_offsets [n_sweeps + 1]
 * the list range of each sweep (d_pos_offsets[n_sweeps] = length).  More than `capacity` positives sets
 * PP_STATUS_CAND_OVERFLOW and drops the excess.  d_cls / d_reg may both be NULL (list only).  num_classes == 9. */
int pp_assign_targets_list(const double* d_a_corners, const double* d_a_centers, const double* d_a_wlh,
                           const double* d_a_yaw, const void* d_anchor_index, int64_t A,
                           const double* d_g_corners, const double* d_g_centers, const double* d_g_wlh,
                           const double* d_g_yaw, const int32_t* d_g_cls, const int64_t* h_gt_offsets,
                           int32_t n_sweeps, int32_t num_classes, double pos_thresh, int32_t* d_pos_anchor,
                           float* d_pos_cls, float* d_pos_reg, int32_t* d_pos_offsets, int32_t capacity,
                           float* d_cls, float* d_reg, int32_t* d_top_anchor, int32_t* d_counts, int32_t* d_status,
                           void* d_workspace, size_t workspace_bytes, pp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * One host call per batch: the whole input path with its stream choreography.
 * pp_step issues, without blocking the host:
 *   [copy stream]  (optional, h_blob != NULL) wait ev_slot_free; copy blob_bytes h_blob -> d_blob; with n_files > 0
 *                  pp_aggregate_sweeps on the points inside it; record ev_ready
 *   [main stream]  wait ev_ready; record ev_fork
 *   [side stream]  wait ev_fork; pp_assign_targets (or pp_assign_targets_list when d_pos_anchor != NULL); record ev_join
 *   [main stream]  pp_input_path PP_STAGE_PILLARIZE; wait ev_prev_encode (the previous batch's encode stage: the
 *                  BatchNorm running statistics are read-modify-write); pp_input_path PP_STAGE_ENCODE; record
 *                  ev_encode_done; wait ev_join; record ev_slot_free; copy {num_pillars [B], counts [B,4], status}
 *                  to h_counters (pinned, 5 B + 1 int32); record ev_done
 * i.e. what data/dataset.py:88-118 + model/model.py:170-177 (feature_net + scatter) do for one mini-batch.  All
 * pointer fields are as in pp_input_path / pp_assign_targets[_list] / pp_aggregate_sweeps; events are cudaEvent_t
 * created by the caller (NULL = that dependency does not exist); streams are cudaStream_t.  The struct is plain
 * data: fill it once per lane / output set and update the fields that change from batch to batch. */
typedef struct pp_step_plan {
  /* streams and events */
  pp_stream_t stream_main, stream_side, stream_copy;
  void *ev_fork, *ev_join, *ev_ready, *ev_slot_free, *ev_prev_encode, *ev_encode_done, *ev_done;
  /* upload */
  const void* h_blob; void* d_blob; size_t blob_bytes;
  int32_t n_files; const int64_t* d_file_offsets; const double* d_file_xforms; float min_dist;
  /* points */
  const void* d_points; int64_t total_points; int64_t point_cols; const int64_t* h_sweep_offsets; int32_t n_sweeps;
  /* K1 + K2 */
  pp_grid grid; int32_t max_points_per_pillar, max_pillars; const float* d_data_mean; const void* d_mean_prepared;
  int32_t C; const float *d_conv_w, *d_conv_b, *d_bn_w, *d_bn_b; float *d_running_mean, *d_running_var;
  int64_t* d_num_batches_tracked; int32_t training; float momentum, eps; int32_t canvas_h, canvas_w;
  float* d_canvas; int64_t* d_indices; int32_t* d_num_pillars; void* d_ws_input; size_t ws_input_bytes;
  /* K3 */
  const double *d_a_corners, *d_a_centers, *d_a_wlh, *d_a_yaw; const void* d_anchor_index; int64_t A;
  const double *d_g_corners, *d_g_centers, *d_g_wlh, *d_g_yaw; const int32_t* d_g_cls; const int64_t* h_gt_offsets;
  int32_t num_classes; double pos_thresh; float *d_cls, *d_reg; int32_t *d_top_anchor, *d_counts;
  int32_t* d_pos_anchor; float *d_pos_cls, *d_pos_reg; int32_t* d_pos_offsets; int32_t pos_capacity;
  void* d_ws_targets; size_t ws_targets_bytes;
  /* status + counters */
  int32_t* d_status; int32_t* h_counters;
} pp_step_plan;
int pp_step(const pp_step_plan* plan);
size_t pp_step_plan_bytes(void);   /* sizeof(pp_step_plan): lets a foreign-language binding check its struct layout */

#ifdef __cplusplus
}
#endif
#endif /* PP_B200_H_ */
