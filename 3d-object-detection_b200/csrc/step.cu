// step.cu -- pp_step: one host call per batch for the whole input path (fused pp_input_path + target assignment),
// with the stream choreography the streaming loop needs done in C: the host-to-device copy of the packed batch on a
// copy stream, target assignment on a side stream next to pillarize / encode on the main stream, the ordering of
// consecutive encode stages (BatchNorm running statistics), the counters' device-to-host copy.  The Python step
// (~25 ctypes / torch calls) cost 355 us of host time per batch against 363 us of GPU time: the loop was host-bound.
#include "common.cuh"

namespace pp {
static inline cudaEvent_t ev(void* p) { return (cudaEvent_t)p; }
}

extern "C" {

size_t pp_step_plan_bytes(void) { return sizeof(pp_step_plan); }

int pp_step(const pp_step_plan* p) {
  using namespace pp;
  if (p == nullptr || p->n_sweeps < 1 || p->stream_main == nullptr || p->stream_side == nullptr ||
      p->ev_fork == nullptr || p->ev_join == nullptr)
    return PP_ERR_INVALID_ARG;
  cudaStream_t sm = (cudaStream_t)p->stream_main, ss = (cudaStream_t)p->stream_side;
  // ---- upload (optional): one copy of the packed batch, then the on-device sweep aggregation -----------------
  if (p->h_blob != nullptr) {
    if (p->stream_copy == nullptr || p->ev_ready == nullptr || p->d_blob == nullptr) return PP_ERR_INVALID_ARG;
    cudaStream_t sc = (cudaStream_t)p->stream_copy;
    if (p->ev_slot_free != nullptr) PP_CUDA(cudaStreamWaitEvent(sc, ev(p->ev_slot_free), 0));
    PP_CUDA(cudaMemcpyAsync(p->d_blob, p->h_blob, p->blob_bytes, cudaMemcpyHostToDevice, sc));
    if (p->n_files > 0) {
      const int rc = pp_aggregate_sweeps((float*)p->d_points, p->total_points, p->point_cols, p->d_file_offsets, p->n_files,
                                         p->d_file_xforms, p->min_dist, nullptr, (pp_stream_t)sc);
      if (rc != PP_OK) return rc;
    }
    PP_CUDA(cudaEventRecord(ev(p->ev_ready), sc));
    PP_CUDA(cudaStreamWaitEvent(sm, ev(p->ev_ready), 0));
  }
  // ---- fork: target assignment on the side stream ------------------------------------------------------------
  PP_CUDA(cudaEventRecord(ev(p->ev_fork), sm));
  PP_CUDA(cudaStreamWaitEvent(ss, ev(p->ev_fork), 0));
  int rc;
  if (p->d_pos_anchor != nullptr) {
    rc = pp_assign_targets_list(p->d_a_corners, p->d_a_centers, p->d_a_wlh, p->d_a_yaw, p->d_anchor_index, p->A,
                                p->d_g_corners, p->d_g_centers, p->d_g_wlh, p->d_g_yaw, p->d_g_cls, p->h_gt_offsets,
                                p->n_sweeps, p->num_classes, p->pos_thresh, p->d_pos_anchor, p->d_pos_cls, p->d_pos_reg,
                                p->d_pos_offsets, p->pos_capacity, p->d_cls, p->d_reg, p->d_top_anchor, p->d_counts,
                                p->d_status, p->d_ws_targets, p->ws_targets_bytes, (pp_stream_t)ss);
  } else {
    rc = pp_assign_targets(p->d_a_corners, p->d_a_centers, p->d_a_wlh, p->d_a_yaw, p->d_anchor_index, p->A,
                           p->d_g_corners, p->d_g_centers, p->d_g_wlh, p->d_g_yaw, p->d_g_cls, p->h_gt_offsets,
                           p->n_sweeps, p->num_classes, p->pos_thresh, p->d_cls, p->d_reg, p->d_top_anchor, p->d_counts,
                           p->d_status, p->d_ws_targets, p->ws_targets_bytes, (pp_stream_t)ss);
  }
  if (rc != PP_OK) return rc;
  PP_CUDA(cudaEventRecord(ev(p->ev_join), ss));
  // ---- main: pillarize stage, then the encode stage ordered after the previous batch's ------------------------
  rc = pp_input_path(p->d_points, PP_F32, p->point_cols, 1, p->h_sweep_offsets, p->n_sweeps, &p->grid,
                     p->max_points_per_pillar, p->max_pillars, p->d_data_mean, p->d_mean_prepared, p->C, p->d_conv_w,
                     p->d_conv_b, p->d_bn_w, p->d_bn_b, p->d_running_mean, p->d_running_var, p->d_num_batches_tracked,
                     p->training, p->momentum, p->eps, p->canvas_h, p->canvas_w, p->d_canvas, nullptr, p->d_indices,
                     p->d_num_pillars, p->d_status, p->d_ws_input, p->ws_input_bytes, PP_STAGE_PILLARIZE, (pp_stream_t)sm);
  if (rc != PP_OK) return rc;
  if (p->ev_prev_encode != nullptr) PP_CUDA(cudaStreamWaitEvent(sm, ev(p->ev_prev_encode), 0));
  rc = pp_input_path(p->d_points, PP_F32, p->point_cols, 1, p->h_sweep_offsets, p->n_sweeps, &p->grid,
                     p->max_points_per_pillar, p->max_pillars, p->d_data_mean, p->d_mean_prepared, p->C, p->d_conv_w,
                     p->d_conv_b, p->d_bn_w, p->d_bn_b, p->d_running_mean, p->d_running_var, p->d_num_batches_tracked,
                     p->training, p->momentum, p->eps, p->canvas_h, p->canvas_w, p->d_canvas, nullptr, p->d_indices,
                     p->d_num_pillars, p->d_status, p->d_ws_input, p->ws_input_bytes, PP_STAGE_ENCODE, (pp_stream_t)sm);
  if (rc != PP_OK) return rc;
  if (p->ev_encode_done != nullptr) PP_CUDA(cudaEventRecord(ev(p->ev_encode_done), sm));
  // ---- join, counters home ----------------------------------------------------------------------------------
  PP_CUDA(cudaStreamWaitEvent(sm, ev(p->ev_join), 0));
  if (p->ev_slot_free != nullptr) PP_CUDA(cudaEventRecord(ev(p->ev_slot_free), sm));
  if (p->h_counters != nullptr) {
    const int B = p->n_sweeps;
    PP_CUDA(cudaMemcpyAsync(p->h_counters, p->d_num_pillars, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, sm));
    PP_CUDA(cudaMemcpyAsync(p->h_counters + B, p->d_counts, sizeof(int32_t) * 4 * B, cudaMemcpyDeviceToHost, sm));
    PP_CUDA(cudaMemcpyAsync(p->h_counters + 5 * B, p->d_status, sizeof(int32_t), cudaMemcpyDeviceToHost, sm));
  }
  if (p->ev_done != nullptr) PP_CUDA(cudaEventRecord(ev(p->ev_done), sm));
  return PP_OK;
}

}  // extern "C"
