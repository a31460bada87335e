// Set-prediction losses of stage 2 (focal labels, L1 + GIoU boxes, Laplace-style w/h uncertainty,
// cardinality / class-error diagnostics) and the stage-1 width/height criterion, forward and analytic
// backward in one small launch each (the reference issues ~40 tiny ATen kernels plus autograd).
// Reference: A2/models/anchor_detr.py:166-289 (SetCriterion losses), A2/models/segmentation.py:198-223
// (sigmoid_focal_loss), A2/util/box_ops.py:46-67 (GIoU), A1/models/anchor_detr.py:317-337
// (BoundingBoxCriterion).
//
// Forward writes the loss values and the UNWEIGHTED per-loss gradients; cdetr_set_loss_bwd combines
// them with the upstream scalars that autograd hands to the criterion's outputs (the weight_dict
// coefficients live in the caller, engine.py:37, exactly as in the reference).
#include "common.cuh"
#include "../../include/cdetr.h"

namespace {

__device__ float block_sum_f(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.0f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

struct GiouGrad { float g; float d[4]; };  // giou and d giou / d (cx, cy, w, h) of the FIRST box

__device__ GiouGrad giou_with_grad(const float* s, const float* t) {
  const float ax0 = s[0] - 0.5f * s[2], ay0 = s[1] - 0.5f * s[3], ax1 = s[0] + 0.5f * s[2], ay1 = s[1] + 0.5f * s[3];
  const float bx0 = t[0] - 0.5f * t[2], by0 = t[1] - 0.5f * t[3], bx1 = t[0] + 0.5f * t[2], by1 = t[1] + 0.5f * t[3];
  const float aw = ax1 - ax0, ah = ay1 - ay0;
  const float area_a = aw * ah, area_b = (bx1 - bx0) * (by1 - by0);
  const float iw_raw = fminf(ax1, bx1) - fmaxf(ax0, bx0), ih_raw = fminf(ay1, by1) - fmaxf(ay0, by0);
  const float iw = fmaxf(iw_raw, 0.0f), ih = fmaxf(ih_raw, 0.0f);
  const float inter = iw * ih;
  const float uni = area_a + area_b - inter;
  const float cw_raw = fmaxf(ax1, bx1) - fminf(ax0, bx0), ch_raw = fmaxf(ay1, by1) - fminf(ay0, by0);
  const float cw = fmaxf(cw_raw, 0.0f), ch = fmaxf(ch_raw, 0.0f);
  const float hull = cw * ch;
  GiouGrad r;
  r.g = inter / uni - (hull - uni) / hull;
  // partials w.r.t. (ax0, ay0, ax1, ay1)
  const float d_area[4] = {-ah, -aw, ah, aw};
  const float iw_on = iw_raw >= 0.0f ? 1.0f : 0.0f, ih_on = ih_raw >= 0.0f ? 1.0f : 0.0f;
  const float d_iw[4] = {ax0 > bx0 ? -iw_on : 0.0f, 0.0f, ax1 < bx1 ? iw_on : 0.0f, 0.0f};
  const float d_ih[4] = {0.0f, ay0 > by0 ? -ih_on : 0.0f, 0.0f, ay1 < by1 ? ih_on : 0.0f};
  const float cw_on = cw_raw >= 0.0f ? 1.0f : 0.0f, ch_on = ch_raw >= 0.0f ? 1.0f : 0.0f;
  const float d_cw[4] = {ax0 < bx0 ? -cw_on : 0.0f, 0.0f, ax1 > bx1 ? cw_on : 0.0f, 0.0f};
  const float d_ch[4] = {0.0f, ay0 < by0 ? -ch_on : 0.0f, 0.0f, ay1 > by1 ? ch_on : 0.0f};
  float dx[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float d_inter = d_iw[k] * ih + iw * d_ih[k];
    const float d_uni = d_area[k] - d_inter;
    const float d_hull = d_cw[k] * ch + cw * d_ch[k];
    const float d_iou = (d_inter * uni - inter * d_uni) / (uni * uni);
    const float d_ratio = (d_uni * hull - uni * d_hull) / (hull * hull);  // d (uni / hull)
    dx[k] = d_iou + d_ratio;
  }
  r.d[0] = dx[0] + dx[2];
  r.d[1] = dx[1] + dx[3];
  r.d[2] = 0.5f * (dx[2] - dx[0]);
  r.d[3] = 0.5f * (dx[3] - dx[1]);
  return r;
}

// single CTA (1024 threads).  out[0..5] = loss_ce, class_error, loss_bbox, loss_giou, cardinality_error,
// loss_variance.  g_* are zero-initialised by the kernel.
__global__ void set_loss_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                                    const float* __restrict__ vars, const float* __restrict__ tgt,
                                    const int* __restrict__ tgt_off, const int64_t* __restrict__ idx_q,
                                    const int64_t* __restrict__ idx_t, const int* __restrict__ idx_n, int B,
                                    int Q, int Kmax, const float* __restrict__ num_boxes_sum, float inv_world,
                                    float alpha, float* __restrict__ out,
                                    float* __restrict__ g_ce, float* __restrict__ g_bbox,
                                    float* __restrict__ g_giou, float* __restrict__ g_var_box,
                                    float* __restrict__ g_var_var, unsigned char* __restrict__ matched,
                                    int* __restrict__ status) {
  __shared__ float red[33];
  const int tid = threadIdx.x, nt = blockDim.x;
  // num_boxes = clamp(sum over ranks / world, min 1)  (A2/models/anchor_detr.py:321-325), read from device memory
  // so the criterion never synchronises with the host
  const float inv_nb = 1.0f / fmaxf(num_boxes_sum[0] * inv_world, 1.0f);
  for (int i = tid; i < B * Q; i += nt) matched[i] = 0;
  for (int i = tid; i < B * Q * 4; i += nt) { g_bbox[i] = 0.0f; g_giou[i] = 0.0f; g_var_box[i] = 0.0f; }
  for (int i = tid; i < B * Q * 2; i += nt) g_var_var[i] = 0.0f;
  __syncthreads();
  // ---- matched pairs: boxes, giou, variance statistics
  float s_l1 = 0.0f, s_giou = 0.0f, s_dw = 0.0f, s_dh = 0.0f, s_iw = 0.0f, s_ih = 0.0f, s_log = 0.0f;
  float n_correct = 0.0f, n_pairs = 0.0f;
  for (int i = tid; i < B * Kmax; i += nt) {
    const int b = i / Kmax, k = i % Kmax;
    if (k >= idx_n[b]) continue;
    const int q = (int)idx_q[(int64_t)b * Kmax + k];
    const int t = (int)idx_t[(int64_t)b * Kmax + k];
    const int64_t row = (int64_t)b * Q + q;
    matched[row] = 1;
    const float* sb = boxes + row * 4;
    const float* tb = tgt + (int64_t)(tgt_off[b] + t) * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float d = sb[c] - tb[c];
      s_l1 += fabsf(d);
      g_bbox[row * 4 + c] = (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) * inv_nb;
    }
    const GiouGrad gg = giou_with_grad(sb, tb);
    s_giou += 1.0f - gg.g;
#pragma unroll
    for (int c = 0; c < 4; ++c) g_giou[row * 4 + c] = -gg.d[c] * inv_nb;
    const float sw = vars[row * 2], sh = vars[row * 2 + 1];
    s_dw += fabsf(sb[2] - tb[2]);
    s_dh += fabsf(sb[3] - tb[3]);
    s_iw += 1.0f / fabsf(sw);
    s_ih += 1.0f / fabsf(sh);
    s_log += fabsf(logf(sw)) + fabsf(logf(sh));
    n_pairs += 1.0f;
    n_correct += (logits[row * 2] >= logits[row * 2 + 1]) ? 1.0f : 0.0f;  // top-1 == class 0
  }
  s_l1 = block_sum_f(s_l1, red);
  s_giou = block_sum_f(s_giou, red);
  s_dw = block_sum_f(s_dw, red);
  s_dh = block_sum_f(s_dh, red);
  s_iw = block_sum_f(s_iw, red);
  s_ih = block_sum_f(s_ih, red);
  s_log = block_sum_f(s_log, red);
  n_pairs = block_sum_f(n_pairs, red);
  n_correct = block_sum_f(n_correct, red);
  const float mw = n_pairs > 0.0f ? s_dw / n_pairs : 0.0f;  // scalar mean L1 (reduction='mean')
  const float mh = n_pairs > 0.0f ? s_dh / n_pairs : 0.0f;
  // ---- variance-loss gradients (second pass over the pairs)
  for (int i = tid; i < B * Kmax; i += nt) {
    const int b = i / Kmax, k = i % Kmax;
    if (k >= idx_n[b]) continue;
    const int q = (int)idx_q[(int64_t)b * Kmax + k];
    const int t = (int)idx_t[(int64_t)b * Kmax + k];
    const int64_t row = (int64_t)b * Q + q;
    const float* sb = boxes + row * 4;
    const float* tb = tgt + (int64_t)(tgt_off[b] + t) * 4;
    const float dw = sb[2] - tb[2], dh = sb[3] - tb[3];
    g_var_box[row * 4 + 2] = (dw > 0.0f ? 1.0f : (dw < 0.0f ? -1.0f : 0.0f)) / n_pairs * s_iw * inv_nb;
    g_var_box[row * 4 + 3] = (dh > 0.0f ? 1.0f : (dh < 0.0f ? -1.0f : 0.0f)) / n_pairs * s_ih * inv_nb;
    const float sw = vars[row * 2], sh = vars[row * 2 + 1];
    const float lsw = logf(sw), lsh = logf(sh);
    const float sgw = sw > 0.0f ? 1.0f : -1.0f, sgh = sh > 0.0f ? 1.0f : -1.0f;
    g_var_var[row * 2] = (-mw * sgw / (sw * sw) + (lsw > 0.0f ? 1.0f : (lsw < 0.0f ? -1.0f : 0.0f)) / sw) * inv_nb;
    g_var_var[row * 2 + 1] = (-mh * sgh / (sh * sh) + (lsh > 0.0f ? 1.0f : (lsh < 0.0f ? -1.0f : 0.0f)) / sh) * inv_nb;
  }
  __syncthreads();
  // ---- focal loss over all (b, q, 2) logits; target one-hot: matched -> (1,0), unmatched -> (0,1)
  float s_ce = 0.0f;
  for (int i = tid; i < B * Q * 2; i += nt) {
    const int c = i & 1;
    const int64_t row = i >> 1;
    const float tcls = (matched[row] != 0) == (c == 0) ? 1.0f : 0.0f;
    const float x = logits[i];
    const float z = tcls > 0.5f ? x : -x;            // p_t = sigmoid(z)
    const float sp = fmaxf(-z, 0.0f) + log1pf(expf(-fabsf(z)));  // softplus(-z) = -log p_t = BCE
    const float pt = 1.0f / (1.0f + expf(-z));
    const float a_t = tcls > 0.5f ? alpha : 1.0f - alpha;
    const float om = 1.0f - pt;
    s_ce += a_t * sp * om * om;
    // d/dx: a_t * s * (1-pt)^2 * (2 pt log pt - (1 - pt)),  s = +1 for t=1, -1 for t=0
    const float sgn = tcls > 0.5f ? 1.0f : -1.0f;
    g_ce[i] = a_t * sgn * om * om * (2.0f * pt * (-sp) - om) * inv_nb;
  }
  s_ce = block_sum_f(s_ce, red);
  // ---- cardinality error
  float s_card = 0.0f;
  for (int b = tid; b < B; b += nt) {
    int cnt = 0;
    for (int q = 0; q < Q; ++q) cnt += logits[((int64_t)b * Q + q) * 2] >= logits[((int64_t)b * Q + q) * 2 + 1];
    s_card += fabsf((float)cnt - (float)(tgt_off[b + 1] - tgt_off[b]));
  }
  s_card = block_sum_f(s_card, red);
  if (tid == 0) {
    out[0] = s_ce * inv_nb;
    out[1] = n_pairs > 0.0f ? 100.0f - n_correct * (100.0f / n_pairs) : 100.0f;
    out[2] = s_l1 * inv_nb;
    out[3] = s_giou * inv_nb;
    out[4] = s_card / (float)B;
    out[5] = (mw * s_iw + mh * s_ih + s_log) * inv_nb;
    // The matcher could not assign (NaN costs: scipy raises ValueError, the reference's GIoU asserts) or an exemplar
    // centre fell outside the feature map (the reference raises IndexError): no exception can cross a stream, so the
    // losses turn NaN, which the reference's training loop treats as fatal (A2/engine.py:47-50).  Flag re-armed.
    if (status != nullptr && *status != 0) {
      const float qnan = __int_as_float(0x7fc00000);
      for (int k = 0; k < 6; ++k) out[k] = qnan;
      *status = 0;
    }
  }
}

// up[0..3]: upstream grads of (loss_ce, loss_bbox, loss_giou, loss_variance)
__global__ void set_loss_bwd_kernel(const float* __restrict__ up, const float* __restrict__ g_ce,
                                    const float* __restrict__ g_bbox, const float* __restrict__ g_giou,
                                    const float* __restrict__ g_var_box, const float* __restrict__ g_var_var,
                                    int64_t rows, float* __restrict__ d_logits, float* __restrict__ d_boxes,
                                    float* __restrict__ d_vars) {
  const float u_ce = up[0], u_bbox = up[1], u_giou = up[2], u_var = up[3];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows * 4;
       i += (int64_t)gridDim.x * blockDim.x) {
    d_boxes[i] = u_bbox * g_bbox[i] + u_giou * g_giou[i] + u_var * g_var_box[i];
    if (i < rows * 2) {
      d_logits[i] = u_ce * g_ce[i];
      d_vars[i] = u_var * g_var_var[i];
    }
  }
}

// Stage 1: boxes are (gt point, predicted wh) vs (gt point, gt wh).  out[0] = loss_wh, out[1] = loss_giou.
__global__ void bbox_loss_fwd_kernel(const float* __restrict__ pred_wh, const float* __restrict__ points,
                                     const float* __restrict__ whs, int64_t n, float* __restrict__ out,
                                     float* __restrict__ g_wh, float* __restrict__ g_giou) {
  __shared__ float red[33];
  float s_l1 = 0.0f, s_g = 0.0f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float sb[4] = {points[i * 2], points[i * 2 + 1], pred_wh[i * 2], pred_wh[i * 2 + 1]};
    const float tb[4] = {points[i * 2], points[i * 2 + 1], whs[i * 2], whs[i * 2 + 1]};
    for (int c = 0; c < 2; ++c) {
      const float d = sb[2 + c] - tb[2 + c];
      s_l1 += fabsf(d);
      g_wh[i * 2 + c] = (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) / (float)(2 * n);
    }
    const GiouGrad gg = giou_with_grad(sb, tb);
    s_g += 1.0f - gg.g;
    g_giou[i * 2] = -gg.d[2] / (float)n;
    g_giou[i * 2 + 1] = -gg.d[3] / (float)n;
  }
  s_l1 = block_sum_f(s_l1, red);
  s_g = block_sum_f(s_g, red);
  if (threadIdx.x == 0) {
    out[0] = s_l1 / (float)(2 * n);
    out[1] = s_g / (float)n;
  }
}
__global__ void bbox_loss_bwd_kernel(const float* __restrict__ up, const float* __restrict__ g_wh,
                                     const float* __restrict__ g_giou, int64_t n2, float* __restrict__ d_wh) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
    d_wh[i] = up[0] * g_wh[i] + up[1] * g_giou[i];
}

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int cdetr_set_loss_fwd(const float* logits, const float* boxes, const float* vars,
                                  const float* tgt_boxes, const int* tgt_off, const int64_t* idx_q,
                                  const int64_t* idx_t, const int* idx_n, int B, int Q, int Kmax,
                                  const float* num_boxes_sum, float inv_world, float focal_alpha, float* out6,
                                  float* g_ce, float* g_bbox,
                                  float* g_giou, float* g_var_box, float* g_var_var, unsigned char* matched,
                                  int* status, cdetr_stream_t s) {
  CDETR_CHECK_ARG(logits && boxes && vars && tgt_boxes && tgt_off && idx_q && idx_t && idx_n && out6 && g_ce &&
                      g_bbox && g_giou && g_var_box && g_var_var && matched && B > 0 && Q > 0 && num_boxes_sum &&
                      inv_world > 0,
                  "set_loss_fwd: bad args");
  set_loss_fwd_kernel<<<1, 1024, 0, STREAM(s)>>>(logits, boxes, vars, tgt_boxes, tgt_off, idx_q, idx_t, idx_n, B,
                                                 Q, Kmax > 0 ? Kmax : 1, num_boxes_sum, inv_world, focal_alpha, out6, g_ce,
                                                 g_bbox,
                                                 g_giou, g_var_box, g_var_var, matched, status);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_set_loss_bwd(const float* upstream4, const float* g_ce, const float* g_bbox,
                                  const float* g_giou, const float* g_var_box, const float* g_var_var,
                                  int64_t rows, float* d_logits, float* d_boxes, float* d_vars,
                                  cdetr_stream_t s) {
  CDETR_CHECK_ARG(upstream4 && g_ce && g_bbox && g_giou && g_var_box && g_var_var && d_logits && d_boxes &&
                      d_vars && rows > 0,
                  "set_loss_bwd: bad args");
  set_loss_bwd_kernel<<<cdiv(rows * 4, 256), 256, 0, STREAM(s)>>>(upstream4, g_ce, g_bbox, g_giou, g_var_box,
                                                                g_var_var, rows, d_logits, d_boxes, d_vars);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_bbox_loss_fwd(const float* pred_wh, const float* points, const float* whs, int64_t n,
                                   float* out2, float* g_wh, float* g_giou, cdetr_stream_t s) {
  CDETR_CHECK_ARG(pred_wh && points && whs && out2 && g_wh && g_giou && n > 0, "bbox_loss_fwd: bad args");
  bbox_loss_fwd_kernel<<<1, 1024, 0, STREAM(s)>>>(pred_wh, points, whs, n, out2, g_wh, g_giou);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_bbox_loss_bwd(const float* upstream2, const float* g_wh, const float* g_giou, int64_t n,
                                   float* d_wh, cdetr_stream_t s) {
  CDETR_CHECK_ARG(upstream2 && g_wh && g_giou && d_wh && n > 0, "bbox_loss_bwd: bad args");
  bbox_loss_bwd_kernel<<<cdiv(n * 2, 256), 256, 0, STREAM(s)>>>(upstream2, g_wh, g_giou, n * 2, d_wh);
  CDETR_CHECK_LAUNCH();
  return 0;
}
