/*
 * cdetr.h — C ABI of libcdetr_sm100a.so, the B200 (sm_100a) hot-path library that sits under the
 * Python classes mirroring Counting-DETR's build_model()/SetCriterion/PostProcess API.
 *
 * The reference (VinAIResearch/Counting-DETR) has no FFI of its own: every FLOP on its hot path is
 * an ATen call made from Python.  Each entry point below therefore cites the
 * reference Python op site (file:line under /root/reference/src/) whose device work it replaces.
 * Path tags: A1 = CountDETR_147_1st_stage, A2 = CountDETR_147_2nd_stage.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes only; every buffer (including workspaces) is owned by the caller;
 *   - all work is enqueued on the passed cudaStream_t, no implicit synchronisation, no host
 *     threads, graph-capturable (no data-dependent host control flow);
 *   - return 0 on success, negative on error; cdetr_last_error() returns a per-thread message;
 *   - "split" tensors are the library's split-bf16 format: two bf16 planes (hi, lo) of a row-major
 *     [rows, ld] matrix, plane 1 starting `plane` elements after plane 0; value = hi + lo.
 */
#ifndef CDETR_H_
#define CDETR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* cdetr_stream_t; /* == cudaStream_t */

int cdetr_version(void);
const char* cdetr_last_error(void);

/* A split-bf16 matrix view: element (r,c) of plane p lives at base[p*plane + r*ld + c] (bf16). */
typedef struct {
  void* base;
  int64_t ld;    /* row pitch in elements, multiple of 8 */
  int64_t plane; /* distance between the hi and lo planes in elements, multiple of 8 */
} cdetr_split_t;

/* ---------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05 + TMA + TMEM), split-bf16 operands, 3-pass error-compensated product,
 * fp32 accumulation.  Replaces every aten::linear / 1x1 conv / im2col-conv / matmul-backward site:
 *   A2/models/resnet.py:143-158 (Bottleneck convs), A2/models/anchor_detr.py:119 (aggr_input_proj),
 *   A2/models/row_column_decoupled_attention.py:172-208,311 (in/out projections),
 *   A2/models/transformer.py:412-426 (FFN), :429-439 (MLP heads), :73-74 (adapt_pos MLPs).
 *   mode 0 (TN): D[M,N] = A[M,K] * B[N,K]^T      (forward, dgrad with pre-transposed weights)
 *   mode 1 (NT): D[M,N] = A[K,M]^T * B[K,N]      (wgrad: contraction over the row index)
 * Epilogue, in this order:  v = acc; v *= row_scale[m]; v += bias[n]; v += add_split[m,n];
 *   v += add_f32[m,n]; if relu v = max(v,0); if mask: v = mask[m,n] > 0 ? v : 0;
 *   out_f32[m,n] = v (or += v atomically when accumulate / split_k > 1); out_split[m,n] = split(v).
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t mode;
  int32_t M, N, K;
  cdetr_split_t a; /* mode 0: [M,K]; mode 1: [K,M] */
  cdetr_split_t b; /* mode 0: [N,K]; mode 1: [K,N] */
  int32_t block_n; /* 0 = auto; else 16..256, multiple of 16 (multiple of 64 in mode 1) */
  int32_t split_k; /* <=1: none; >1 requires out_f32 only (atomic accumulation) */
  const float* row_scale; /* [M] or NULL */
  const float* bias;      /* [N] or NULL */
  cdetr_split_t add_split; /* base NULL = none */
  const float* add_f32;
  int64_t ld_add_f32;
  cdetr_split_t mask; /* only plane 0 (hi) is read; base NULL = none */
  int32_t relu;
  int32_t accumulate;
  float* out_f32;
  int64_t ld_out_f32;
  cdetr_split_t out_split; /* base NULL = none */
  /* Implicit 3x3 convolution (stride 1, padding = dilation), conv_taps = 9 (0 = plain GEMM).  The "conv operand"
   * is an NHWC split activation x [B*H*W, C] read with shifted TMA windows (out-of-image taps zero-filled by the
   * TMA unit) instead of an im2col matrix; replaces F.conv2d at A2/models/resnet.py:147-149 and its autograd:
   *   mode 0: conv operand = a.  K = 9*C ordered (tap, c); row m of D is pixel m:
   *           D[m, n] = sum_{tap,c} x[pixel m + conv_sign * off(tap), c] * B[n, tap*C + c]
   *           (conv_sign = +1: forward with weights [N, 9C];  -1: dgrad with dy as x and weights [Cin, 9*Cout])
   *   mode 1: conv operand = b.  N = 9*C ordered (tap, c); K = B*H*W pixels:
   *           D[m, tap*C + c] = sum_p A[p, m] * x[pixel p + off(tap), c]            (wgrad, staged as [Cout, 9*Cin])
   * off(tap) = ((tap/3 - 1) * dil rows, (tap%3 - 1) * dil columns).  Needs C % 64 == 0, W a divisor of 128 and
   * H*W a multiple of 128 (a 128-row tile = whole image rows); callers lower other shapes through cdetr_im2col3x3. */
  int32_t conv_taps;
  int32_t conv_H, conv_W, conv_C, conv_dil, conv_sign;
  /* Precision policy: which of the three split-bf16 partial products are issued (bit 0: hi_a*hi_b, always on; bit 1:
   * hi_a*lo_b; bit 2: lo_a*hi_b).  0 or 7 = all three (fp32-equivalent, 2^-16 relative per product); 5 = operand b
   * rounded to bf16; 3 = operand a rounded to bf16; 1 = plain bf16 product. */
  int32_t pass_mask;
} cdetr_gemm_t;

int cdetr_gemm(const cdetr_gemm_t* g, cdetr_stream_t stream);
/* Debug hook (no reference counterpart): the following cdetr_gemm launches record 8 device timestamps of CTA 0 each
 * (ns, %globaltimer) into buf[8 * launch]; NULL stops recording. */
int cdetr_gemm_debug_timeline(long long* buf, int capacity_launches);

/* ---------------------------------------------------------------------------------------------
 * Backbone layout kernels (split-bf16 NHWC).  Replace, together with cdetr_gemm:
 *   A2/models/backbone.py:50-60 (FrozenBatchNorm2d -> folded scale/shift),
 *   A2/models/resnet.py:143-158 (Bottleneck convs), :263-271 (stem conv + max-pool).
 * --------------------------------------------------------------------------------------------- */
int cdetr_bn_fold(const float* w, const float* b, const float* rm, const float* rv, float eps, int c,
                  float* scale, float* shift, cdetr_stream_t s);
/* w [cout, cin, taps] fp32 -> dst [cout, taps*cin] and/or dst_t [taps*cin, cout] (row_scale folded) */
int cdetr_pack_weight(const float* w, int cout, int cin, int taps, const float* row_scale,
                      cdetr_split_t dst, cdetr_split_t dst_t, cdetr_stream_t s);
/* w [cout, cin, taps] fp32 -> dst [cin, taps*cout] (row_scale[cout] folded): B operand of the implicit-conv dgrad */
int cdetr_pack_weight_dgrad(const float* w, int cout, int cin, int taps, const float* row_scale,
                            cdetr_split_t dst, cdetr_stream_t s);
/* All weights of the model in ONE launch (the re-pack that follows every optimizer step: 125 cdetr_pack_weight + 53
 * cdetr_bn_fold calls otherwise).  table: device array, one entry per weight; blocks: device int32 [nblocks][2] =
 * (entry, chunk): a block converts elements [chunk*chunk_elems, +chunk_elems) of its entry.  Per entry: the FrozenBN fold
 * (scale = bn_w * rsqrt(bn_rv + eps), shift = bn_b - bn_rm * scale; written by the entry's first block) and any of the
 * three packed layouts cdetr_pack_weight / cdetr_pack_weight_dgrad produce (NULL base = not wanted). */
typedef struct {
  const float* w;                                  /* [cout, cin, taps] fp32 master weight */
  const float *bn_w, *bn_b, *bn_rm, *bn_rv;        /* FrozenBatchNorm buffers or NULL */
  float *scale, *shift;                            /* fold outputs [cout] (with bn_w) */
  cdetr_split_t dst;                               /* [cout, taps*cin]   forward operand */
  cdetr_split_t dst_t;                             /* [taps*cin, cout]   dgrad operand */
  cdetr_split_t dst_d;                             /* [cin, taps*cout]   implicit-conv dgrad operand */
  int32_t cout, cin, taps, pad_;
} cdetr_pack_entry_t;
int cdetr_mt_pack_weights(const cdetr_pack_entry_t* table, const int32_t* blocks, int nblocks, int chunk_elems, float eps,
                          cdetr_stream_t s);
/* grad [cout, cin, taps] += g [cout, taps*cin] */
int cdetr_unpack_conv_grad(const float* g, int cout, int cin, int taps, float* grad, cdetr_stream_t s);
int cdetr_to_split(const float* x, int64_t rows, int cols, int64_t ld_x, cdetr_split_t dst, cdetr_stream_t s);
int cdetr_from_split(cdetr_split_t src, int64_t rows, int cols, float* y, int64_t ld_y, cdetr_stream_t s);
int cdetr_stem_im2col(const float* img_nchw, int B, int H, int W, cdetr_split_t col, cdetr_stream_t s);
/* the whole stem in one kernel, no im2col matrix: out[(b,oy,ox), 0:64] = relu(conv7x7 s2 p3(img) + shift); w = FrozenBN-
   scaled conv1 weights [64, ld >= 152] split bf16 with k = (r*7 + s)*3 + c, shift = FrozenBN shift [64]
   (A2/models/resnet.py:263-271, A2/models/backbone.py:22-60) */
int cdetr_stem_conv(const float* img_nchw, int B, int H, int W, cdetr_split_t w, const float* shift, cdetr_split_t out,
                    cdetr_stream_t s);
int cdetr_im2col3x3(cdetr_split_t x, int B, int H, int W, int C, int stride, int dil, cdetr_split_t col,
                    cdetr_stream_t s);
int cdetr_col2im3x3(cdetr_split_t dcol, int B, int H, int W, int C, int stride, int dil, cdetr_split_t mask,
                    cdetr_split_t dx, cdetr_stream_t s);
int cdetr_maxpool3x3s2(cdetr_split_t x, int B, int H, int W, int C, cdetr_split_t y, cdetr_stream_t s);
int cdetr_subsample2(cdetr_split_t x, int B, int H, int W, int C, cdetr_split_t y, cdetr_stream_t s);
int cdetr_upsample2_zero(cdetr_split_t dy, int B, int H, int W, int C, cdetr_split_t dx, cdetr_stream_t s);

/* Exemplar feature injection, A2/models/backbone.py:116-136: cat = [x | x * p_b], p_b = mean of x at the
 * exemplar centres (yx int32 [n_ex, 2], taken from sample 0's rects by the caller, reference quirk). */
int cdetr_exemplar_concat(cdetr_split_t x, int B, int H, int W, int C, const int* centres_yx, int n_ex,
                          float* p_out, cdetr_split_t cat, cdetr_stream_t s);
int cdetr_exemplar_concat_bwd(cdetr_split_t dcat, cdetr_split_t x, const float* p, int B, int H, int W, int C,
                              const int* centres_yx, int n_ex, float* dp_scratch, cdetr_split_t mask,
                              cdetr_split_t dx, cdetr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * Norms: A2/models/anchor_detr.py:81,119 (GroupNorm 32x8), A2/models/transformer.py:233-426 (LayerNorm).
 * --------------------------------------------------------------------------------------------- */
int cdetr_groupnorm_fwd(const float* x, int B, int N, int C, int G, const float* gamma, const float* beta,
                        float eps, float* y, cdetr_split_t y_split, float* stats, cdetr_stream_t s);
int cdetr_groupnorm_bwd(const float* dy, const float* x, int B, int N, int C, int G, const float* gamma,
                        const float* stats, float* dx, cdetr_split_t dx_split, float* dgamma, float* dbeta,
                        cdetr_stream_t s);
int cdetr_layernorm_fwd(const float* x, const float* res, int64_t M, int C, const float* gamma,
                        const float* beta, float eps, float* z_out, float* y, cdetr_split_t y_split,
                        float* stats, cdetr_stream_t s);
int cdetr_layernorm_bwd(const float* dy, const float* dy2, const float* z, const float* stats, int64_t M, int C,
                        const float* gamma, float* dz, cdetr_split_t dz_split, float* dgamma, float* dbeta,
                        cdetr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * Transformer glue: A2/models/transformer.py:474-503 (sine embeddings, mask2pos), :248-256,:378-392
 * (position broadcasts), A2/models/row_column_decoupled_attention.py:212-213 (key means),
 * A2/models/transformer.py:193-202 + A2/util/misc.py:475-479 (box head).
 * --------------------------------------------------------------------------------------------- */
int cdetr_sine_embed(const float* pos, int64_t n, int pos_stride, int num_feats, int off, int ld, float* out,
                     cdetr_stream_t s);
int cdetr_sine_embed_bwd(const float* pos, int64_t n, int pos_stride, int num_feats, int off, int ld,
                         const float* demb, float* dpos, cdetr_stream_t s);
int cdetr_add_bcast(const float* x, const float* y, int64_t M, int E, int mode, int H, int W, int64_t rows_y,
                    float* out, cdetr_split_t out_split, cdetr_stream_t s);
int cdetr_reduce_axis(const float* x, int B, int H, int W, int E, int axis, float scale, const float* add,
                      int accumulate, float* out, cdetr_split_t out_split, cdetr_stream_t s);
int cdetr_combine_bcast(const float* a, const float* b, const float* c, const float* row, float sr,
                        const float* col, float sc, int64_t M, int E, int H, int W, float* out, cdetr_stream_t s);
int cdetr_colsum(const float* x, cdetr_split_t x_split, int64_t ld, int64_t M, int N, float* out, cdetr_stream_t s);
int cdetr_box_head_fwd(const float* t, const float* ref, int64_t M, float* boxes, cdetr_stream_t s);
int cdetr_box_head_bwd(const float* dboxes, const float* boxes, const float* ref, int64_t M, float* dt,
                       cdetr_split_t dt_split, float* dref, cdetr_stream_t s);
int cdetr_scale(float* x, int64_t n, float a, cdetr_stream_t s);
/* Padding mask [B,S1,S2] (1 = padded pixel) -> what the transformer reads of it: the nearest-neighbour downsample to
 * the H x W feature map (A2/models/backbone.py:112), of which only the first row / first column are ever used:
 * mask_row [B,W] / mask_col [B,H] (RCDA key padding, A2/models/row_column_decoupled_attention.py:238-249) and the
 * normalised positions pos_row [B,W] / pos_col [B,H] of mask2pos (A2/models/transformer.py:497-503). */
int cdetr_mask_prepare(const uint8_t* mask, int B, int S1, int S2, int H, int W, uint8_t* mask_row, uint8_t* mask_col,
                       float* pos_row, float* pos_col, cdetr_stream_t s);
/* Exemplar centres on the device (no host read of the rects): rects0 = fp32 [n_ex,4] xyxy in [0,1] of SAMPLE 0,
 * centres_yx int32 [n_ex,2] = (int((y1*H + y2*H)/2), int((x1*W + x2*W)/2)) in fp32 with truncation
 * (A2/models/backbone.py:122-128).  Out-of-range centres are clamped and bit 1 of *status is set. */
int cdetr_exemplar_centres(const float* rects0, int n_ex, int H, int W, int* centres_yx, int* status, cdetr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * Attention cores.  RCDA: A2/models/row_column_decoupled_attention.py:210-291 (q scaling, row/col logits,
 * padding masks, two softmaxes, decoupled contraction); attention maps stored transposed
 * ([B,heads,W,L], [B,heads,H,L]).  MHA: decoder self-attention, A2/models/transformer.py:366-372.
 * --------------------------------------------------------------------------------------------- */
int cdetr_rcda_fwd(int B, int L, int H, int W, int E, int nh, const float* qr, const float* qc, const float* kr,
                   const float* kc, const float* v, const uint8_t* mask_row, const uint8_t* mask_col, float* ar,
                   float* ac, cdetr_split_t o, cdetr_stream_t s);
/* Same contract as cdetr_rcda_fwd with the contraction on tcgen05 tensor cores (V given as a split tensor
 * [B*H*W, E]); requires H, W <= 64 (V resident in shared memory up to 32 x 32, streamed through a TMA ring above). */
int cdetr_rcda_fwd_tc(int B, int L, int H, int W, int E, int nh, const float* qr, const float* qc, const float* kr,
                      const float* kc, cdetr_split_t v, const uint8_t* mask_row, const uint8_t* mask_col, float* ar,
                      float* ac, cdetr_split_t o, cdetr_stream_t s);
/* Query-side backward on tensor cores (dS maps + dq); pair it with cdetr_rcda_bwd_k + cdetr_rcda_bwd_v_tc.  H, W <= 64. */
int cdetr_rcda_bwd_q_tc(int B, int L, int H, int W, int E, int nh, const float* kr, const float* kc, cdetr_split_t v,
                        const float* ar, const float* ac, const float* d_o, float* dsr, float* dsc,
                        cdetr_split_t dqr, cdetr_split_t dqc, cdetr_stream_t s);
/* Value-side backward on tensor cores (d_o as a split tensor [B*L, E]); H, W <= 64. */
int cdetr_rcda_bwd_v_tc(int B, int L, int H, int W, int E, int nh, const float* ar, const float* ac,
                        cdetr_split_t d_o, cdetr_split_t dv, cdetr_stream_t s);
/* Key-side backward only: dK_r, dK_c (split) from the dS maps. */
int cdetr_rcda_bwd_k(int B, int L, int H, int W, int E, int nh, const float* qr, const float* qc, const float* dsr,
                     const float* dsc, cdetr_split_t dkr, cdetr_split_t dkc, cdetr_stream_t s);
/* Key/value-side backward given dsr/dsc: dK_r, dK_c (split) and dV (split). */
int cdetr_rcda_bwd_kv(int B, int L, int H, int W, int E, int nh, const float* qr, const float* qc, const float* ar,
                      const float* ac, const float* d_o, const float* dsr, const float* dsc, cdetr_split_t dkr,
                      cdetr_split_t dkc, cdetr_split_t dv, cdetr_stream_t s);
int cdetr_rcda_bwd(int B, int L, int H, int W, int E, int nh, const float* qr, const float* qc, const float* kr,
                   const float* kc, const float* v, const float* ar, const float* ac, const float* d_o,
                   float* dsr, float* dsc, cdetr_split_t dqr, cdetr_split_t dqc, cdetr_split_t dkr,
                   cdetr_split_t dkc, cdetr_split_t dv, cdetr_stream_t s);
int cdetr_mha_fwd(int B, int L, int E, int nh, const float* q, const float* k, const float* v, int64_t ldq,
                  cdetr_split_t o, float* lse, cdetr_stream_t s);
int cdetr_mha_bwd(int B, int L, int E, int nh, const float* q, const float* k, const float* v, int64_t ldq,
                  cdetr_split_t o, const float* lse, const float* d_o, float* dsum, cdetr_split_t dq,
                  cdetr_split_t dk, cdetr_split_t dv, cdetr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * Matcher + criterion.  A2/models/matcher.py:221-247 (cost + scipy linear_sum_assignment per image),
 * A2/models/anchor_detr.py:166-289 (SetCriterion losses), A1/models/anchor_detr.py:317-337.
 * tgt_off: int32 [B+1] prefix offsets into the concatenated target boxes [sum T, 4] (cxcywh).
 * cost: B slabs of Q*Tmax floats, each [T_b, Q] when T_b < Q (LSAP rows = targets) else [Q, T_b].
 * out_q/out_t: int64 [B, min(Q,Tmax)], first out_n[b] valid, query index ascending (scipy order).
 * --------------------------------------------------------------------------------------------- */
int cdetr_match_cost(const float* logits, int num_logits, const float* boxes, const float* tgt_boxes,
                     const int* tgt_off, int B, int Q, int Tmax, float w_class, float w_bbox, float w_giou,
                     float* cost, cdetr_stream_t s);
int cdetr_lsap(const float* cost, const int* tgt_off, int B, int Q, int Tmax, int64_t* out_q, int64_t* out_t,
               int* out_n, int* status, cdetr_stream_t s);
/* out6 = loss_ce, class_error, loss_bbox, loss_giou, cardinality_error, loss_variance.
 * num_boxes_sum: device float = sum of target counts over all ranks; num_boxes = max(sum * inv_world, 1). */
int cdetr_set_loss_fwd(const float* logits, const float* boxes, const float* vars, const float* tgt_boxes,
                       const int* tgt_off, const int64_t* idx_q, const int64_t* idx_t, const int* idx_n, int B,
                       int Q, int Kmax, const float* num_boxes_sum, float inv_world, float focal_alpha,
                       float* out6, float* g_ce,
                       float* g_bbox, float* g_giou, float* g_var_box, float* g_var_var, unsigned char* matched,
                       int* status /* NULL or the flag cdetr_lsap / cdetr_exemplar_centres set: non-zero turns the six
                                      losses NaN (the reference raises there) and re-arms the flag */,
                       cdetr_stream_t s);
int cdetr_set_loss_bwd(const float* upstream4, const float* g_ce, const float* g_bbox, const float* g_giou,
                       const float* g_var_box, const float* g_var_var, int64_t rows, float* d_logits,
                       float* d_boxes, float* d_vars, cdetr_stream_t s);
int cdetr_bbox_loss_fwd(const float* pred_wh, const float* points, const float* whs, int64_t n, float* out2,
                        float* g_wh, float* g_giou, cdetr_stream_t s);
int cdetr_bbox_loss_bwd(const float* upstream2, const float* g_wh, const float* g_giou, int64_t n, float* d_wh,
                        cdetr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * Output formats of the two-stage recipe on the device; the reference copies every prediction tensor
 * to the host and formats with numpy.  sizes_hw: fp32 [B,2] = (height, width) per image, like `target_sizes`.
 *   cdetr_postprocess_topk: PostProcess.forward, A2/models/anchor_detr.py:370-402 (sigmoid, top-k over the flattened
 *     [Q*C] scores in descending order, ties by ascending index; labels = idx % C; xyxy boxes in pixels); k <= Q*C <= 16384.
 *   cdetr_infer_select: A2/infer.py:74-118: queries with sigmoid(logit[..., 0]) >= threshold, compacted in query order
 *     (torch.where); out_* rows [B, Q]: query index, score, bbox = int([cx*w, cy*h, bw*w, bh*h]), area = int(bw*w * bh*h),
 *     point = int(ref * [w, h]) -- the integers of the reference's int() on numpy float32 scalars; out_count[b] rows valid.
 *   cdetr_pseudo_label_format: A1/engine.py:148-166: bbox = int([x*s0, y*s1, w*s0, h*s1]), area = int(w*s0 * h*s1) for n
 *     (point, predicted wh) pairs; size2 = device fp32 [2] = orig_size as the reference indexes it.
 * --------------------------------------------------------------------------------------------- */
int cdetr_postprocess_topk(const float* logits, const float* boxes, const float* sizes_hw, int B, int Q, int C, int k,
                           float* out_scores, int64_t* out_labels, float* out_boxes, cdetr_stream_t s);
int cdetr_infer_select(const float* logits, int C, const float* boxes, const float* ref_points, const float* sizes_hw,
                       int B, int Q, float threshold, int* out_count, int* out_query, float* out_score, int* out_bbox,
                       int* out_area, int* out_point, cdetr_stream_t s);
int cdetr_pseudo_label_format(const float* points, const float* whs, const float* size2, int64_t n, int* out_bbox,
                              int* out_area, cdetr_stream_t s);
/* Input side: transforms.ToTensor() + Normalize(mean, std) of the reference's datasets
 * (A2/data/fsc147.py:22-24,82) on the device: uint8 [B,H,W,3] pixels -> fp32 [B,3,H,W], y = ((u8 / 255) - mean[c]) / std[c]
 * in torchvision's fp32 operation order (bit-identical); mean3 / std3 are HOST arrays of 3 floats. */
int cdetr_normalize_u8(const uint8_t* src_hwc, int B, int H, int W, const float* mean3_host, const float* std3_host,
                       float* dst_nchw, cdetr_stream_t s);

/* ---------------------------------------------------------------------------------------------
 * Optimizer tail: multi-tensor gradient-norm clipping + AdamW.  Replace
 *   A2/engine.py:53-56  torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
 *   A2/engine.py:57 / A2/main.py:188  torch.optim.AdamW(param_dicts, lr, weight_decay).step()
 * table: device array of tensors; blocks: device int32 [nblocks][2] = (tensor index, chunk index), every block
 * handles elements [chunk_index*chunk, +chunk) of its tensor.  hyper: device float [ngroups][4] = {lr, weight_decay,
 * 0, 0}.  norm_out: device float[3] = {sum of squares, total L2 norm, clip coefficient min(1, max_norm/(norm+1e-6))}.
 * cdetr_mt_adamw applies norm_out[2] to the gradients on the fly when norm_out != NULL (fused clip) and increments
 * the device step counter.  Update order follows torch.optim.adamw._single_tensor_adamw.
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  float* p;       /* parameter */
  float* g;       /* gradient */
  float* m;       /* exp_avg */
  float* v;       /* exp_avg_sq */
  int64_t n;      /* elements */
  int32_t group;  /* row of `hyper` */
  int32_t pad_;
} cdetr_mt_tensor_t;
int cdetr_mt_grad_norm(const cdetr_mt_tensor_t* table, const int32_t* blocks, int nblocks, int chunk, float max_norm,
                       float* partial /* [nblocks] scratch */, float* norm_out, cdetr_stream_t s);
int cdetr_mt_clip_scale(const cdetr_mt_tensor_t* table, const int32_t* blocks, int nblocks, int chunk,
                        const float* norm_out, cdetr_stream_t s);
int cdetr_mt_adamw(const cdetr_mt_tensor_t* table, const int32_t* blocks, int nblocks, int chunk, const float* hyper,
                   float beta1, float beta2, float eps, int* step, const float* norm_out, cdetr_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* CDETR_H_ */
