// GroupNorm(32,256) and LayerNorm(256) forward/backward on channels-last fp32 rows, with the split-bf16
// copy of the output written in the same pass (it is the A operand of the next tensor-core GEMM).
// Reference op sites: A2/models/anchor_detr.py:81,119 (GroupNorm after the 1x1 projection),
// A2/models/transformer.py:233,274,334,339,372,404,420,425 (LayerNorm, post-norm residual blocks).
#include "common.cuh"
#include "../../include/cdetr.h"

namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
  // red: >= 33 floats of shared memory
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

// ---------------------------------------------------------------- GroupNorm: one CTA per (sample, group)
__global__ void groupnorm_fwd_kernel(const float* __restrict__ x, int N, int C, int G,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float eps, float* __restrict__ y, __nv_bfloat16* y_hi,
                                     __nv_bfloat16* y_lo, int64_t ld_split, float* __restrict__ stats) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[33];
  const int b = blockIdx.x / G, g = blockIdx.x % G;
  const int cg = C / G;
  const int64_t base = (int64_t)b * N * C + (int64_t)g * cg;
  const int cnt = N * cg;
  float s = 0.0f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) s += x[base + (int64_t)(i / cg) * C + (i % cg)];
  const float mean = block_sum(s, red) / cnt;
  float q = 0.0f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const float d = x[base + (int64_t)(i / cg) * C + (i % cg)] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(block_sum(q, red) / cnt + eps);
  if (threadIdx.x == 0) {
    stats[2 * blockIdx.x] = mean;
    stats[2 * blockIdx.x + 1] = rstd;
  }
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int n = i / cg, c = g * cg + (i % cg);
    const int64_t off = (int64_t)b * N * C + (int64_t)n * C + c;
    const float v = (x[off] - mean) * rstd * gamma[c] + beta[c];
    if (y) y[off] = v;
    if (y_hi) {
      const int64_t so = ((int64_t)b * N + n) * ld_split + c;
      split_bf16(v, y_hi[so], y_lo[so]);
    }
  }
}

// dx = rstd * (dxh - mean(dxh) - xh * mean(dxh*xh)),  dxh = dy*gamma;  dgamma += sum dy*xh, dbeta += sum dy
__global__ void groupnorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, int N,
                                     int C, int G, const float* __restrict__ gamma,
                                     const float* __restrict__ stats, float* __restrict__ dx,
                                     __nv_bfloat16* dx_hi, __nv_bfloat16* dx_lo, int64_t ld_split,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[33];
  __shared__ float cs[2][64];  // per-channel partial sums (cg <= 64)
  const int b = blockIdx.x / G, g = blockIdx.x % G;
  const int cg = C / G;
  const int64_t base = (int64_t)b * N * C + (int64_t)g * cg;
  const int cnt = N * cg;
  const float mean = stats[2 * blockIdx.x], rstd = stats[2 * blockIdx.x + 1];
  if (threadIdx.x < 2 * 64) (&cs[0][0])[threadIdx.x] = 0.0f;
  __syncthreads();
  float s1 = 0.0f, s2 = 0.0f;
  // threads stride by blockDim (multiple of cg) so each thread always sees the same channel
  const int cl = threadIdx.x % cg;
  float g_acc = 0.0f, b_acc = 0.0f;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int64_t off = base + (int64_t)(i / cg) * C + (i % cg);
    const float xh = (x[off] - mean) * rstd;
    const float d = dy[off];
    const float dxh = d * gamma[g * cg + cl];
    s1 += dxh;
    s2 += dxh * xh;
    g_acc += d * xh;
    b_acc += d;
  }
  atomicAdd(&cs[0][cl], g_acc);
  atomicAdd(&cs[1][cl], b_acc);
  const float m1 = block_sum(s1, red) / cnt;
  const float m2 = block_sum(s2, red) / cnt;
  if (threadIdx.x < cg) {
    atomicAdd(dgamma + g * cg + threadIdx.x, cs[0][threadIdx.x]);
    atomicAdd(dbeta + g * cg + threadIdx.x, cs[1][threadIdx.x]);
  }
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int n = i / cg, c = g * cg + (i % cg);
    const int64_t off = (int64_t)b * N * C + (int64_t)n * C + c;
    const float xh = (x[off] - mean) * rstd;
    const float v = rstd * (dy[off] * gamma[c] - m1 - xh * m2);
    if (dx) dx[off] = v;
    if (dx_hi) {
      const int64_t so = ((int64_t)b * N + n) * ld_split + c;
      split_bf16(v, dx_hi[so], dx_lo[so]);
    }
  }
}

// ---------------------------------------------------------------- LayerNorm over C = 256: one warp per row
constexpr int LN_C = 256;

__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                     int64_t M, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, float* __restrict__ z_out,
                                     float* __restrict__ y, __nv_bfloat16* y_hi, __nv_bfloat16* y_lo,
                                     int64_t ld_split, float* __restrict__ stats) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* xr = x + row * LN_C + lane * 8;
  float v[8];
  {
    const float4 a = *reinterpret_cast<const float4*>(xr);
    const float4 b = *reinterpret_cast<const float4*>(xr + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  if (res) {
    const float* rr = res + row * LN_C + lane * 8;
    const float4 a = *reinterpret_cast<const float4*>(rr);
    const float4 b = *reinterpret_cast<const float4*>(rr + 4);
    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
  }
  if (z_out) {
    float* zr = z_out + row * LN_C + lane * 8;
    *reinterpret_cast<float4*>(zr) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(zr + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  const float mean = warp_sum(s) * (1.0f / LN_C);
  float q = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float d = v[j] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / LN_C) + eps);
  if (stats && lane == 0) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * gamma[lane * 8 + j] + beta[lane * 8 + j];
  if (y) {
    float* yr = y + row * LN_C + lane * 8;
    *reinterpret_cast<float4*>(yr) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(yr + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
  if (y_hi) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      split_bf16_pair(o[2 * j], o[2 * j + 1], hw[j], lw[j]);
    }
    *reinterpret_cast<uint4*>(y_hi + row * ld_split + lane * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(y_lo + row * ld_split + lane * 8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

// dz = rstd * (dxh - mean(dxh) - xh*mean(dxh*xh)); dy may be the sum of two gradient streams (dy, dy2).
// dgamma/dbeta are accumulated: each CTA reduces its rows in shared memory, then 256 atomics.
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dy2,
                                     const float* __restrict__ z, const float* __restrict__ stats,
                                     int64_t M, const float* __restrict__ gamma, float* __restrict__ dz,
                                     __nv_bfloat16* dz_hi, __nv_bfloat16* dz_lo, int64_t ld_split,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     int rows_per_cta) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sg[LN_C], sb[LN_C];
  for (int i = threadIdx.x; i < LN_C; i += blockDim.x) { sg[i] = 0.0f; sb[i] = 0.0f; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float ga[8], ba[8], gm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ga[j] = 0.0f; ba[j] = 0.0f; gm[j] = gamma[lane * 8 + j]; }
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  for (int r = warp; r < rows_per_cta; r += nw) {
    const int64_t row = row0 + r;
    if (row >= M) break;
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float d[8], xh[8];
    {
      const float* p = dy + row * LN_C + lane * 8;
      const float4 a = *reinterpret_cast<const float4*>(p);
      const float4 b = *reinterpret_cast<const float4*>(p + 4);
      d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
    }
    if (dy2) {
      const float* p = dy2 + row * LN_C + lane * 8;
      const float4 a = *reinterpret_cast<const float4*>(p);
      const float4 b = *reinterpret_cast<const float4*>(p + 4);
      d[0] += a.x; d[1] += a.y; d[2] += a.z; d[3] += a.w; d[4] += b.x; d[5] += b.y; d[6] += b.z; d[7] += b.w;
    }
    {
      const float* p = z + row * LN_C + lane * 8;
      const float4 a = *reinterpret_cast<const float4*>(p);
      const float4 b = *reinterpret_cast<const float4*>(p + 4);
      xh[0] = a.x; xh[1] = a.y; xh[2] = a.z; xh[3] = a.w; xh[4] = b.x; xh[5] = b.y; xh[6] = b.z; xh[7] = b.w;
    }
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xh[j] = (xh[j] - mean) * rstd;
      ga[j] += d[j] * xh[j];
      ba[j] += d[j];
      d[j] *= gm[j];
      s1 += d[j];
      s2 += d[j] * xh[j];
    }
    s1 = warp_sum(s1) * (1.0f / LN_C);
    s2 = warp_sum(s2) * (1.0f / LN_C);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = rstd * (d[j] - s1 - xh[j] * s2);
    if (dz) {
      float* p = dz + row * LN_C + lane * 8;
      *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
    if (dz_hi) {
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        split_bf16_pair(o[2 * j], o[2 * j + 1], hw[j], lw[j]);
      }
      *reinterpret_cast<uint4*>(dz_hi + row * ld_split + lane * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(dz_lo + row * ld_split + lane * 8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&sg[lane * 8 + j], ga[j]);
    atomicAdd(&sb[lane * 8 + j], ba[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < LN_C; i += blockDim.x) {
    atomicAdd(dgamma + i, sg[i]);
    atomicAdd(dbeta + i, sb[i]);
  }
}

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int cdetr_groupnorm_fwd(const float* x, int B, int N, int C, int G, const float* gamma,
                                   const float* beta, float eps, float* y, cdetr_split_t y_split,
                                   float* stats, cdetr_stream_t s) {
  CDETR_CHECK_ARG(x && gamma && beta && stats && C % G == 0 && C / G <= 64, "groupnorm_fwd: bad args");
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(y_split.base);
  launch_light(groupnorm_fwd_kernel, dim3(B * G), dim3(512), 0, STREAM(s), x, N, C, G, gamma, beta, eps, y, hi,
                                                     hi ? hi + y_split.plane : nullptr, y_split.ld, stats);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_groupnorm_bwd(const float* dy, const float* x, int B, int N, int C, int G,
                                   const float* gamma, const float* stats, float* dx,
                                   cdetr_split_t dx_split, float* dgamma, float* dbeta,
                                   cdetr_stream_t s) {
  CDETR_CHECK_ARG(dy && x && gamma && stats && dgamma && dbeta && C % G == 0 && C / G <= 64 &&
                      512 % (C / G) == 0,
                  "groupnorm_bwd: bad args");
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(dx_split.base);
  launch_light(groupnorm_bwd_kernel, dim3(B * G), dim3(512), 0, STREAM(s), dy, x, N, C, G, gamma, stats, dx, hi,
                                                     hi ? hi + dx_split.plane : nullptr, dx_split.ld,
                                                     dgamma, dbeta);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_layernorm_fwd(const float* x, const float* res, int64_t M, int C,
                                   const float* gamma, const float* beta, float eps, float* z_out,
                                   float* y, cdetr_split_t y_split, float* stats, cdetr_stream_t s) {
  CDETR_CHECK_ARG(x && gamma && beta && C == LN_C && M > 0, "layernorm_fwd: needs C == 256");
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(y_split.base);
  const int warps = 8;
  launch_light(layernorm_fwd_kernel, dim3(cdiv(M, warps)), dim3(warps * 32), 0, STREAM(s), 
      x, res, M, gamma, beta, eps, z_out, y, hi, hi ? hi + y_split.plane : nullptr, y_split.ld, stats);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_layernorm_bwd(const float* dy, const float* dy2, const float* z, const float* stats,
                                   int64_t M, int C, const float* gamma, float* dz, cdetr_split_t dz_split,
                                   float* dgamma, float* dbeta, cdetr_stream_t s) {
  CDETR_CHECK_ARG(dy && z && stats && gamma && dgamma && dbeta && C == LN_C && M > 0,
                  "layernorm_bwd: needs C == 256");
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(dz_split.base);
  const int rows_per_cta = 64;
  launch_light(layernorm_bwd_kernel, dim3(cdiv(M, rows_per_cta)), dim3(256), 0, STREAM(s), 
      dy, dy2, z, stats, M, gamma, dz, hi, hi ? hi + dz_split.plane : nullptr, dz_split.ld, dgamma,
      dbeta, rows_per_cta);
  CDETR_CHECK_LAUNCH();
  return 0;
}
