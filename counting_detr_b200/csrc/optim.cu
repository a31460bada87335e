// Optimizer tail of the train step: gradient-norm clipping + AdamW over every trainable tensor
// as multi-tensor kernels.  Replaces, in the reference's train loop (A2/engine.py:53-57, A2/main.py:157-189):
//   torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)   (2 x 265 small launches + a host sync-free norm)
//   torch.optim.AdamW(param_dicts, lr, weight_decay).step()          (foreach path: ~10 passes over p/g/m/v)
// with: one sum-of-squares pass over the gradients (per-block partials, then a fixed-order final reduction: the
// norm is deterministic), and ONE pass that applies the clip coefficient and the AdamW update (reads g, p, m, v
// once, writes p, m, v once: 28 bytes per parameter, the HBM floor for fp32 AdamW).
//
// Work distribution is a block table (tensor, chunk) built once by the host (apex-style multi_tensor_apply), so
// tensors need not be contiguous with each other; hyper-parameters live in a small device array (graph friendly:
// an LR scheduler only rewrites that array).  The step counter is a device int incremented by the kernel.
#include "common.cuh"
#include "../../include/cdetr.h"

namespace {

constexpr int MT_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.0f;
  if (warp == 0) {
    t = lane < (MT_THREADS / 32) ? red[lane] : 0.0f;
    t = warp_sum(t);
  }
  return t;  // valid in warp 0
}

__global__ void __launch_bounds__(MT_THREADS)
mt_sumsq_kernel(const cdetr_mt_tensor_t* __restrict__ table, const int32_t* __restrict__ blocks, int chunk,
                float* __restrict__ partial) {
  __shared__ float red[MT_THREADS / 32];
  const int t = blocks[2 * blockIdx.x], c = blocks[2 * blockIdx.x + 1];
  const cdetr_mt_tensor_t e = table[t];
  const int64_t begin = (int64_t)c * chunk;
  const int64_t end = min(e.n, begin + chunk);
  const float* g = e.g + begin;
  const int n = (int)(end - begin);
  float s = 0.0f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int i = threadIdx.x; i < n4; i += MT_THREADS) {
      const float4 v = __ldg(g4 + i);
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += MT_THREADS) s += g[i] * g[i];
  } else {
    for (int i = threadIdx.x; i < n; i += MT_THREADS) s += g[i] * g[i];
  }
  const float tot = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// out[0] = sum of squares, out[1] = total norm, out[2] = clip coefficient min(1, max_norm / (norm + 1e-6))
__global__ void __launch_bounds__(MT_THREADS)
mt_norm_finish_kernel(const float* __restrict__ partial, int n, float max_norm, float* __restrict__ out) {
  __shared__ double red[MT_THREADS];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += MT_THREADS) s += (double)partial[i];   // fixed order per thread
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = MT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float ss = (float)red[0];
    const float norm = sqrtf(ss);
    out[0] = ss;
    out[1] = norm;
    float coef = 1.0f;
    if (max_norm > 0.0f) coef = fminf(max_norm / (norm + 1e-6f), 1.0f);   // torch.nn.utils.clip_grad_norm_
    out[2] = coef;
  }
}

__global__ void __launch_bounds__(MT_THREADS)
mt_scale_kernel(const cdetr_mt_tensor_t* __restrict__ table, const int32_t* __restrict__ blocks, int chunk,
                const float* __restrict__ norm_out) {
  const float coef = norm_out[2];
  if (coef >= 1.0f) return;        // torch multiplies by the clamped coefficient (== 1): a no-op
  const int t = blocks[2 * blockIdx.x], c = blocks[2 * blockIdx.x + 1];
  const cdetr_mt_tensor_t e = table[t];
  const int64_t begin = (int64_t)c * chunk;
  const int n = (int)(min(e.n, begin + chunk) - begin);
  float* g = e.g + begin;
  for (int i = threadIdx.x; i < n; i += MT_THREADS) g[i] *= coef;
}

// hyper: [ngroups][4] = lr, weight_decay, (unused), (unused);  step: device int, value BEFORE this step.
__global__ void __launch_bounds__(MT_THREADS)
mt_adamw_kernel(const cdetr_mt_tensor_t* __restrict__ table, const int32_t* __restrict__ blocks, int chunk,
                const float* __restrict__ hyper, float beta1, float beta2, float eps,
                const int* __restrict__ step, const float* __restrict__ norm_out) {
  const int t = blocks[2 * blockIdx.x], c = blocks[2 * blockIdx.x + 1];
  const cdetr_mt_tensor_t e = table[t];
  const float lr = hyper[4 * e.group], wd = hyper[4 * e.group + 1];
  const float coef = norm_out != nullptr ? norm_out[2] : 1.0f;
  const int st = *step + 1;
  // bias corrections as torch computes them on the host in double and rounds the derived scalars to fp32
  const double bc1 = 1.0 - pow((double)beta1, (double)st);
  const double bc2 = 1.0 - pow((double)beta2, (double)st);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const float decay = (float)(1.0 - (double)lr * (double)wd);
  const int64_t begin = (int64_t)c * chunk;
  const int n = (int)(min(e.n, begin + chunk) - begin);
  float* p = e.p + begin;
  const float* g = e.g + begin;
  float* m = e.m + begin;
  float* v = e.v + begin;
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  auto upd = [&](float& pv, float gv, float& mv, float& vv) {
    gv *= coef;
    pv *= decay;                                       // param.mul_(1 - lr * weight_decay)
    mv = mv + omb1 * (gv - mv);                        // exp_avg.lerp_(grad, 1 - beta1)
    vv = vv * beta2 + omb2 * gv * gv;                  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv = pv - step_size * (mv / denom);                // param.addcdiv_(exp_avg, denom, value=-step_size)
  };
  if (vec) {
    const int n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (int i = threadIdx.x; i < n4; i += MT_THREADS) {
      float4 pv = p4[i], mv = m4[i], vv = v4[i];
      const float4 gv = __ldg(g4 + i);
      upd(pv.x, gv.x, mv.x, vv.x);
      upd(pv.y, gv.y, mv.y, vv.y);
      upd(pv.z, gv.z, mv.z, vv.z);
      upd(pv.w, gv.w, mv.w, vv.w);
      p4[i] = pv; m4[i] = mv; v4[i] = vv;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += MT_THREADS) upd(p[i], g[i], m[i], v[i]);
  } else {
    for (int i = threadIdx.x; i < n; i += MT_THREADS) upd(p[i], g[i], m[i], v[i]);
  }
}

__global__ void mt_step_inc_kernel(int* step) { *step += 1; }

}  // namespace

extern "C" int cdetr_mt_grad_norm(const cdetr_mt_tensor_t* table, const int32_t* blocks, int nblocks, int chunk,
                                  float max_norm, float* partial, float* norm_out, cdetr_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  CDETR_CHECK_ARG(table && blocks && partial && norm_out && nblocks > 0 && chunk > 0, "mt_grad_norm: bad args");
  mt_sumsq_kernel<<<nblocks, MT_THREADS, 0, s>>>(table, blocks, chunk, partial);
  CDETR_CHECK_LAUNCH();
  mt_norm_finish_kernel<<<1, MT_THREADS, 0, s>>>(partial, nblocks, max_norm, norm_out);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_mt_clip_scale(const cdetr_mt_tensor_t* table, const int32_t* blocks, int nblocks, int chunk,
                                   const float* norm_out, cdetr_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  CDETR_CHECK_ARG(table && blocks && norm_out && nblocks > 0 && chunk > 0, "mt_clip_scale: bad args");
  mt_scale_kernel<<<nblocks, MT_THREADS, 0, s>>>(table, blocks, chunk, norm_out);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_mt_adamw(const cdetr_mt_tensor_t* table, const int32_t* blocks, int nblocks, int chunk,
                              const float* hyper, float beta1, float beta2, float eps, int* step,
                              const float* norm_out, cdetr_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  CDETR_CHECK_ARG(table && blocks && hyper && step && nblocks > 0 && chunk > 0, "mt_adamw: bad args");
  mt_adamw_kernel<<<nblocks, MT_THREADS, 0, s>>>(table, blocks, chunk, hyper, beta1, beta2, eps, step, norm_out);
  CDETR_CHECK_LAUNCH();
  mt_step_inc_kernel<<<1, 1, 0, s>>>(step);
  CDETR_CHECK_LAUNCH();
  return 0;
}
