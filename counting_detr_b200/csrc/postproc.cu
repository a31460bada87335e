// Output formats of the two-stage recipe on the device: what the reference does on the host
// with numpy after a device -> host copy of every prediction tensor.
//   cdetr_postprocess_topk   PostProcess.forward, A2/models/anchor_detr.py:370-402: sigmoid, top-k over the flattened
//                            (query, class) scores, label = idx % C, box = cxcywh -> xyxy of query idx / C, scaled to pixels
//   cdetr_infer_select       stage-2 inference, A2/infer.py:74-118: keep queries with sigmoid(logit_0) >= threshold (in
//                            (image, query) order like torch.where), boxes / reference points scaled to the original
//                            image size and truncated to int exactly like the reference's int() on numpy float32 scalars
//   cdetr_pseudo_label_format  stage-1 pseudo-label pass, A1/engine.py:148-166: points and predicted w/h scaled to the
//                            original size, bbox = [int(cx), int(cy), int(w), int(h)], area = int(w * h)
// All arithmetic that feeds an int() truncation is done with round-to-nearest fp32 intrinsics (no FMA contraction), so
// the integers are the reference's integers.
#include "common.cuh"
#include "../../include/cdetr.h"

namespace {

__device__ __forceinline__ float sigmoid_rn(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// ---- top-k: one CTA per image, bitonic sort of (score, index) keys in shared memory (descending score, ascending index)
__global__ void postprocess_topk_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                                        const float* __restrict__ sizes_hw, int Q, int C, int k, int npow2,
                                        float* __restrict__ out_scores, int64_t* __restrict__ out_labels,
                                        float* __restrict__ out_boxes) {
  extern __shared__ unsigned long long keys[];
  const int b = blockIdx.x, n = Q * C;
  const float* lg = logits + (int64_t)b * n;
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    unsigned long long key = 0ull;   // padding sorts last
    if (i < n) {
      const float s = sigmoid_rn(lg[i]);            // in (0, 1): the raw bit pattern orders like the value
      key = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
    }
    keys[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= npow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool desc = (i & size) == 0;
          const unsigned long long a = keys[i], c = keys[j];
          if (desc ? a < c : a > c) { keys[i] = c; keys[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  const float img_h = sizes_hw[2 * b], img_w = sizes_hw[2 * b + 1];
  for (int t = threadIdx.x; t < k; t += blockDim.x) {
    const unsigned long long key = keys[t];
    const int idx = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
    const int q = idx / C;
    out_scores[(int64_t)b * k + t] = __uint_as_float((unsigned)(key >> 32));
    out_labels[(int64_t)b * k + t] = idx % C;
    const float* bx = boxes + ((int64_t)b * Q + q) * 4;
    const float cx = bx[0], cy = bx[1], w = bx[2], h = bx[3];
    float* o = out_boxes + ((int64_t)b * k + t) * 4;
    o[0] = __fmul_rn(__fsub_rn(cx, __fmul_rn(0.5f, w)), img_w);     // box_cxcywh_to_xyxy (A2/util/box_ops.py:17-20) * scale
    o[1] = __fmul_rn(__fsub_rn(cy, __fmul_rn(0.5f, h)), img_h);
    o[2] = __fmul_rn(__fadd_rn(cx, __fmul_rn(0.5f, w)), img_w);
    o[3] = __fmul_rn(__fadd_rn(cy, __fmul_rn(0.5f, h)), img_h);
  }
}

// ---- threshold selection with order-preserving compaction: one CTA (1024 threads) per image
__global__ void infer_select_kernel(const float* __restrict__ logits, int C, const float* __restrict__ boxes,
                                    const float* __restrict__ ref_points, const float* __restrict__ sizes_hw, int Q,
                                    float threshold, int* __restrict__ out_count, int* __restrict__ out_query,
                                    float* __restrict__ out_score, int* __restrict__ out_bbox, int* __restrict__ out_area,
                                    int* __restrict__ out_point) {
  __shared__ int warp_tot[32];
  __shared__ int base;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float ori_h = sizes_hw[2 * b], ori_w = sizes_hw[2 * b + 1];
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int q0 = 0; q0 < Q; q0 += blockDim.x) {
    const int q = q0 + threadIdx.x;
    float p = 0.0f;
    bool keep = false;
    if (q < Q) {
      p = sigmoid_rn(logits[((int64_t)b * Q + q) * C]);
      keep = p >= threshold;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (keep) {
      const int slot = off + __popc(bal & ((1u << lane) - 1u));
      const int64_t o = (int64_t)b * Q + slot;
      const float* bx = boxes + ((int64_t)b * Q + q) * 4;
      const float x = __fmul_rn(bx[0], ori_w), y = __fmul_rn(bx[1], ori_h);      // A2/infer.py:88-92
      const float w_ = __fmul_rn(bx[2], ori_w), h_ = __fmul_rn(bx[3], ori_h);
      out_query[o] = q;
      out_score[o] = p;
      out_bbox[o * 4 + 0] = (int)x; out_bbox[o * 4 + 1] = (int)y; out_bbox[o * 4 + 2] = (int)w_; out_bbox[o * 4 + 3] = (int)h_;
      out_area[o] = (int)__fmul_rn(w_, h_);                                       // int(w * h), A2/infer.py:108
      const float* rp = ref_points + ((int64_t)b * Q + q) * 2;
      out_point[o * 2 + 0] = (int)__fmul_rn(rp[0], ori_w);                         // A2/infer.py:84-86,112
      out_point[o * 2 + 1] = (int)__fmul_rn(rp[1], ori_h);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < nwarp; ++w) t += warp_tot[w];
      base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out_count[b] = base;
}

__global__ void pseudo_label_kernel(const float* __restrict__ points, const float* __restrict__ whs,
                                    const float* __restrict__ size2, int64_t n, int* __restrict__ out_bbox,
                                    int* __restrict__ out_area) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s0 = size2[0], s1 = size2[1];        // orig_size[0] scales x / w, orig_size[1] scales y / h (A1/engine.py:152-157)
  const float x = __fmul_rn(points[2 * i], s0), y = __fmul_rn(points[2 * i + 1], s1);
  const float w = __fmul_rn(whs[2 * i], s0), h = __fmul_rn(whs[2 * i + 1], s1);
  out_bbox[4 * i + 0] = (int)x; out_bbox[4 * i + 1] = (int)y; out_bbox[4 * i + 2] = (int)w; out_bbox[4 * i + 3] = (int)h;
  out_area[i] = (int)__fmul_rn(w, h);
}

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int cdetr_postprocess_topk(const float* logits, const float* boxes, const float* sizes_hw, int B, int Q, int C,
                                      int k, float* out_scores, int64_t* out_labels, float* out_boxes, cdetr_stream_t s) {
  CDETR_CHECK_ARG(logits && boxes && sizes_hw && out_scores && out_labels && out_boxes && B > 0 && Q > 0 && C > 0 && k > 0,
                  "postprocess_topk: bad args");
  CDETR_CHECK_ARG(k <= Q * C, "postprocess_topk: k=%d exceeds the %d scores per image (torch.topk raises)", k, Q * C);
  int npow2 = 1;
  while (npow2 < Q * C) npow2 <<= 1;
  CDETR_CHECK_ARG(npow2 <= 16384, "postprocess_topk: at most 16384 (query, class) scores per image (got %d)", Q * C);
  const int smem = npow2 * 8;
  static DevAttrCache cfg = {};
  CDETR_CHECK_CUDA(cdetr_ensure_smem(postprocess_topk_kernel, 16384 * 8, &cfg));
  postprocess_topk_kernel<<<B, 1024, smem, STREAM(s)>>>(logits, boxes, sizes_hw, Q, C, k, npow2, out_scores, out_labels,
                                                         out_boxes);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_infer_select(const float* logits, int C, const float* boxes, const float* ref_points,
                                  const float* sizes_hw, int B, int Q, float threshold, int* out_count, int* out_query,
                                  float* out_score, int* out_bbox, int* out_area, int* out_point, cdetr_stream_t s) {
  CDETR_CHECK_ARG(logits && boxes && ref_points && sizes_hw && out_count && out_query && out_score && out_bbox &&
                      out_area && out_point && B > 0 && Q > 0 && C > 0,
                  "infer_select: bad args");
  infer_select_kernel<<<B, 1024, 0, STREAM(s)>>>(logits, C, boxes, ref_points, sizes_hw, Q, threshold, out_count,
                                                  out_query, out_score, out_bbox, out_area, out_point);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_pseudo_label_format(const float* points, const float* whs, const float* size2, int64_t n,
                                         int* out_bbox, int* out_area, cdetr_stream_t s) {
  CDETR_CHECK_ARG(points && whs && size2 && out_bbox && out_area && n > 0, "pseudo_label_format: bad args");
  pseudo_label_kernel<<<cdiv(n, 256), 256, 0, STREAM(s)>>>(points, whs, size2, n, out_bbox, out_area);
  CDETR_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------ input side
// transforms.ToTensor() + transforms.Normalize(mean, std) of the reference's datasets (A2/data/fsc147.py:22-24,82,
// A1/datasets/fscd_147.py) on the device: uint8 HWC pixels (what PIL hands over after the resize) -> fp32 NCHW
//   y[b,c,h,w] = ((float(u8) / 255) - mean[c]) / std[c]        same fp32 operation order as torchvision -> bit-identical
// so a batch crosses PCIe as 1 byte per value instead of 4.  One thread per 4 consecutive pixels of a row.
namespace {
__global__ void normalize_u8_kernel(const uint8_t* __restrict__ src, int B, int H, int W, float m0, float m1, float m2,
                                    float s0, float s1, float s2, float* __restrict__ dst) {
  const int64_t npix = (int64_t)B * H * W;
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = p / ((int64_t)H * W), hw = p - b * (int64_t)H * W;
    const uint8_t* s = src + p * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      dst[(b * 3 + c) * (int64_t)H * W + hw] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)s[c], 255.0f), mean[c]), stdv[c]);
  }
}
}  // namespace

extern "C" int cdetr_normalize_u8(const uint8_t* src_hwc, int B, int H, int W, const float* mean3_host,
                                  const float* std3_host, float* dst_nchw, cdetr_stream_t s) {
  CDETR_CHECK_ARG(src_hwc && dst_nchw && mean3_host && std3_host && B > 0 && H > 0 && W > 0, "normalize_u8: bad args");
  const int64_t npix = (int64_t)B * H * W;
  int grid = (int)((npix + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  normalize_u8_kernel<<<grid, 256, 0, STREAM(s)>>>(src_hwc, B, H, W, mean3_host[0], mean3_host[1], mean3_host[2],
                                                   std3_host[0], std3_host[1], std3_host[2], dst_nchw);
  CDETR_CHECK_LAUNCH();
  return 0;
}
