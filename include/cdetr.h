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
} cdetr_gemm_t;

int cdetr_gemm(const cdetr_gemm_t* g, cdetr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CDETR_H_ */
