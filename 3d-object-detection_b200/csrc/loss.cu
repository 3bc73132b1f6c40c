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
#include "common.cuh"

namespace pp {

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

// in-place tanh on network channel 6 of reg_out (model/loss.py:50)
__global__ void __launch_bounds__(256) k_loss_tanh(float* __restrict__ reg, int B, size_t plane, int CR) {
  const size_t n = (size_t)B * plane;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const size_t b = i / plane, c = i - b * plane;
    float* p = reg + (b * CR + 6) * plane + c;
    *p = tanhf(*p);
  }
}

__global__ void __launch_bounds__(256) k_loss_count(const float* __restrict__ reg_t, size_t n_anchors, int* __restrict__ n_pos) {
  int c = 0;
  for (size_t a = (size_t)blockIdx.x * 256 + threadIdx.x; a < n_anchors; a += (size_t)gridDim.x * 256)
    c += __ldg(reg_t + a * 9) == 1.f ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane_id() == 0 && c) atomicAdd(n_pos, c);                  // integer: exact and order-independent
}

// positives: smooth-L1 / orientation terms and the gradient rows (grad_reg is zeroed by the caller)
__global__ void __launch_bounds__(256) k_loss_reg(const float* __restrict__ reg, const float* __restrict__ reg_t,
                                                  int B, int H, int W, int Ad, int R, const int* __restrict__ n_pos,
                                                  float b_reg, float b_ort, float* __restrict__ grad,
                                                  double* __restrict__ partials) {
  __shared__ double s_red[8];
  const size_t plane = (size_t)H * W;
  const size_t A = plane * Ad, n = A * B;
  const int np = *n_pos;
  const float inv_r = np > 0 ? b_reg / (7.f * (float)np) : 0.f;
  const float inv_o = np > 0 ? b_ort / (float)np : 0.f;
  double sr = 0.0, so = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float* tr = reg_t + i * 9;
    if (__ldg(tr) != 1.f) continue;                              // (:54) pos_anchors
    const size_t b = i / A, a = i - b * A;
    const size_t cell = a / Ad;
    const int d = (int)(a - cell * Ad);
    const float* src = reg + (b * (size_t)Ad * R + (size_t)d * R) * plane + cell;
    float* dst = grad != nullptr ? grad + (b * (size_t)Ad * R + (size_t)d * R) * plane + cell : nullptr;
    for (int c = 0; c < 7; ++c) {
      const float v = src[(size_t)c * plane];
      const float df = v - __ldg(tr + 1 + c);
      const float ad = fabsf(df);
      sr += (double)(ad < 1.f ? 0.5f * df * df : ad - 0.5f);     // F.smooth_l1_loss, beta = 1
      float g = (ad < 1.f ? df : (df > 0.f ? 1.f : -1.f)) * inv_r;
      if (d == 0 && c == 6) g *= 1.f - v * v;                    // v is already tanh(.): chain rule of (:50)
      if (dst != nullptr) dst[(size_t)c * plane] = g;
    }
    if (R > 7) {
      const float o = src[(size_t)7 * plane], ot = __ldg(tr + 8);
      so += (double)(fmaxf(o, 0.f) - o * ot + log1pf(expf(-fabsf(o))));
      if (dst != nullptr) dst[(size_t)7 * plane] = (1.f / (1.f + expf(-o)) - ot) * inv_o;
    }
  }
  const double a0 = block_sum(sr, s_red);
  const double a1 = block_sum(so, s_red);
  if (threadIdx.x == 0) {
    partials[(size_t)blockIdx.x * 2] = a0;
    partials[(size_t)blockIdx.x * 2 + 1] = a1;
  }
}

__global__ void __launch_bounds__(256) k_loss_finalize(const double* __restrict__ pc, int nc, const double* __restrict__ pr,
                                                       int nr, const int* __restrict__ n_pos, double n_cls, float b_cls,
                                                       float b_reg, float b_ort, float* __restrict__ losses) {
  __shared__ double s_red[8];
  double c = 0.0, r = 0.0, o = 0.0;
  for (int i = threadIdx.x; i < nc; i += 256) c += pc[i];
  for (int i = threadIdx.x; i < nr; i += 256) { r += pr[2 * i]; o += pr[2 * i + 1]; }
  const double C = block_sum(c, s_red), Rr = block_sum(r, s_red), O = block_sum(o, s_red);
  if (threadIdx.x == 0) {
    const int np = *n_pos;
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

struct LossWs {
  double* pc;
  double* pr;
  int* n_pos;
};

static int loss_reg_blocks() { return sm_count() * 8; }

template <class A>
static void loss_layout(A& a, LossWs* ws, int B, int H, int W) {
  const size_t nc = (size_t)B * H * ((W + kLossCells - 1) / kLossCells);
  auto p0 = a.template take<double>(nc);
  auto p1 = a.template take<double>((size_t)loss_reg_blocks() * 2);
  auto p2 = a.template take<int>(64);
  if (ws) { ws->pc = p0; ws->pr = p1; ws->n_pos = p2; }
}

struct SizeArena4 {
  size_t used = 0;
  template <class T>
  T* take(size_t count) { used += align_up(count * sizeof(T)); return nullptr; }
};

}  // namespace pp

extern "C" {

size_t pp_loss_workspace_bytes(int32_t B, int32_t H, int32_t W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  pp::SizeArena4 a;
  pp::loss_layout(a, (pp::LossWs*)nullptr, B, H, W);
  return a.used + pp::kAlign;
}

int pp_loss(const float* d_cls_out, float* d_reg_out, const float* d_cls_t, const float* d_reg_t, int32_t B,
            int32_t H, int32_t W, int32_t anchors_per_cell, int32_t num_classes, int32_t reg_dims, float gamma,
            float alpha_pos, float b_cls, float b_reg, float b_ort, float* d_scores, float* d_grad_cls,
            float* d_grad_reg, float* d_losses, void* d_workspace, size_t workspace_bytes, pp_stream_t stream) {
  using namespace pp;
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_cls_out || !d_reg_out || !d_cls_t || !d_reg_t || !d_losses || B < 1 || H < 1 || W < 1 ||
      anchors_per_cell < 1 || num_classes < 1 || reg_dims < 7)
    return PP_ERR_INVALID_ARG;
  const int CK = anchors_per_cell * num_classes, CR = anchors_per_cell * reg_dims;
  if (CK > kLossMaxCh || CR <= 6 || B > 65535 || H > 65535) return PP_ERR_UNSUPPORTED;
  Arena arena(d_workspace, workspace_bytes);
  LossWs ws{};
  loss_layout(arena, &ws, B, H, W);
  if (!arena.ok) return PP_ERR_WORKSPACE;
  const size_t plane = (size_t)H * W;
  const size_t n_anchors = plane * anchors_per_cell * B;
  const double n_cls = (double)n_anchors * num_classes;
  PP_CUDA(cudaMemsetAsync(ws.n_pos, 0, sizeof(int), st));
  const dim3 gc((W + kLossCells - 1) / kLossCells, H, B);
  PP_KERNEL("k_loss_cls", st,
            (k_loss_cls<<<gc, 256, 0, st>>>(d_cls_out, d_cls_t, H, W, CK, gamma, alpha_pos, (float)((double)b_cls / n_cls),
                                            d_scores, d_grad_cls, ws.pc)));
  PP_KERNEL("k_loss_tanh", st, (k_loss_tanh<<<(int)((B * plane + 255) / 256), 256, 0, st>>>(d_reg_out, B, plane, CR)));
  const int nb = loss_reg_blocks();
  PP_KERNEL("k_loss_count", st, (k_loss_count<<<nb, 256, 0, st>>>(d_reg_t, n_anchors, ws.n_pos)));
  if (d_grad_reg != nullptr) PP_CUDA(cudaMemsetAsync(d_grad_reg, 0, (size_t)B * CR * plane * sizeof(float), st));
  PP_KERNEL("k_loss_reg", st,
            (k_loss_reg<<<nb, 256, 0, st>>>(d_reg_out, d_reg_t, B, H, W, anchors_per_cell, reg_dims, ws.n_pos, b_reg, b_ort,
                                            d_grad_reg, ws.pr)));
  PP_KERNEL("k_loss_finalize", st,
            (k_loss_finalize<<<1, 256, 0, st>>>(ws.pc, (int)((size_t)gc.x * gc.y * gc.z), ws.pr, nb, ws.n_pos, n_cls, b_cls,
                                                b_reg, b_ort, d_losses)));
  return PP_OK;
}

}  // extern "C"
