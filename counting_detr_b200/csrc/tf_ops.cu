// Small HBM-bound kernels of the RCDA encoder/decoder around the GEMMs (E = 256 channels-last rows):
// sine position embeddings (+ their derivative w.r.t. learned anchor points), broadcast adds of the
// row/column embeddings, the H/W means that commute with the key projections, bias-gradient column
// sums and the box head's sigmoid / inverse-sigmoid epilogue.
// Reference op sites: A2/models/transformer.py:474-503 (pos2posemb1d/2d, mask2pos), :248-256,:378-392
// (with_pos_embed broadcasts), A2/models/row_column_decoupled_attention.py:212-213 (k means),
// A2/models/transformer.py:193-202 + A2/util/misc.py:475-479 (box head, inverse_sigmoid).
#include "common.cuh"
#include "../../include/cdetr.h"

namespace {

constexpr float TWO_PI = 6.283185307179586f;

__device__ __forceinline__ float dim_t_of(int i, int num_feats) {
  // temperature ** (2 * (i // 2) / num_feats), temperature = 10000
  return powf(10000.0f, (float)(2 * (i / 2)) / (float)num_feats);
}

// emb[n, off + i] = sin/cos(pos[n*pos_stride] * 2pi / dim_t(i)),  i < num_feats (even: sin, odd: cos)
__global__ void sine_embed_kernel(const float* __restrict__ pos, int64_t n, int pos_stride, int num_feats,
                                  int off, int ld, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = n * num_feats;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t % num_feats);
    const int64_t r = t / num_feats;
    const float a = pos[r * pos_stride] * TWO_PI / dim_t_of(i, num_feats);
    out[r * ld + off + i] = (i & 1) ? cosf(a) : sinf(a);
  }
}

// dpos[n*pos_stride] += sum_i demb[n, off+i] * d/dpos
__global__ void sine_embed_bwd_kernel(const float* __restrict__ pos, int64_t n, int pos_stride,
                                      int num_feats, int off, int ld, const float* __restrict__ demb,
                                      float* __restrict__ dpos) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float p = pos[row * pos_stride];
  float acc = 0.0f;
  for (int i = lane; i < num_feats; i += 32) {
    const float w = TWO_PI / dim_t_of(i, num_feats);
    const float a = p * w;
    acc += demb[row * ld + off + i] * ((i & 1) ? -sinf(a) : cosf(a)) * w;
  }
  acc = warp_sum(acc);
  if (lane == 0) atomicAdd(dpos + row * pos_stride, acc);
}

// out[m,:] = x[m,:] + y[idx(m),:] as split (and optionally fp32);  E % 8 == 0
//   mode 0: idx = m            mode 1: idx = (m / (H*W)) * W + m % W   (row embedding, bcast over h)
//   mode 2: idx = m / W        (column embedding [B,H,E], bcast over w)   mode 3: idx = m % rows_y
__global__ void add_bcast_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t M,
                                 int E, int mode, int H, int W, int64_t rows_y, float* __restrict__ out,
                                 __nv_bfloat16* o_hi, __nv_bfloat16* o_lo, int64_t ld_split) {
  pdl_trigger();
  pdl_wait();
  const int e8 = E / 8;
  const int64_t total = M * e8;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(t % e8) * 8;
    const int64_t m = t / e8;
    int64_t yi;
    if (mode == 0) yi = m;
    else if (mode == 1) yi = (m / ((int64_t)H * W)) * W + (m % W);
    else if (mode == 2) yi = m / W;
    else yi = m % rows_y;
    float v[8];
    const float4 a = *reinterpret_cast<const float4*>(x + m * E + c0);
    const float4 b = *reinterpret_cast<const float4*>(x + m * E + c0 + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    if (y) {
      const float4 c = *reinterpret_cast<const float4*>(y + yi * E + c0);
      const float4 d = *reinterpret_cast<const float4*>(y + yi * E + c0 + 4);
      v[0] += c.x; v[1] += c.y; v[2] += c.z; v[3] += c.w; v[4] += d.x; v[5] += d.y; v[6] += d.z; v[7] += d.w;
    }
    if (out) {
      *reinterpret_cast<float4*>(out + m * E + c0) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(out + m * E + c0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (o_hi) {
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        split_bf16_pair(v[2 * j], v[2 * j + 1], hw[j], lw[j]);
      }
      *reinterpret_cast<uint4*>(o_hi + m * ld_split + c0) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(o_lo + m * ld_split + c0) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
  }
}

// x [B,H,W,E] fp32.  axis 1: out[b,w,:] = scale * sum_h x[b,h,w,:] (+ add[b,w,:]);
//                    axis 2: out[b,h,:] = scale * sum_w x[b,h,w,:] (+ add[b,h,:]).
// accumulate: out += result (fp32 only).  One thread per output element, coalesced over E.
__global__ void reduce_axis_kernel(const float* __restrict__ x, int B, int H, int W, int E, int axis,
                                   float scale, const float* __restrict__ add, int accumulate,
                                   float* __restrict__ out, __nv_bfloat16* o_hi, __nv_bfloat16* o_lo,
                                   int64_t ld_split) {
  pdl_trigger();
  pdl_wait();
  const int R = axis == 1 ? W : H;
  const int64_t total = (int64_t)B * R * E;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(t % E);
    const int r = (int)((t / E) % R);
    const int b = (int)(t / ((int64_t)E * R));
    float s = 0.0f;
    if (axis == 1) {
      for (int h = 0; h < H; ++h) s += x[(((int64_t)b * H + h) * W + r) * E + e];
    } else {
      for (int w = 0; w < W; ++w) s += x[(((int64_t)b * H + r) * W + w) * E + e];
    }
    s *= scale;
    if (add) s += add[t];
    if (out) {
      if (accumulate) out[t] += s; else out[t] = s;
    }
    if (o_hi) {
      const int64_t so = ((int64_t)b * R + r) * ld_split + e;
      split_bf16(s, o_hi[so], o_lo[so]);
    }
  }
}

// reduce_axis, four channels per thread (E % 4 == 0, 16-byte aligned fp32 operands, no split output): the reduced axis is
// walked with independent 16-byte loads, 8 in flight
__global__ void reduce_axis4_kernel(const float4* __restrict__ x, int B, int H, int W, int E4, int axis, float scale,
                                    const float4* __restrict__ add, int accumulate, float4* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int R = axis == 1 ? W : H, n = axis == 1 ? H : W;
  const int64_t total = (int64_t)B * R * E4;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(t % E4);
    const int r = (int)((t / E4) % R);
    const int b = (int)(t / ((int64_t)E4 * R));
    // element i of the walk: axis 1 -> (h = i, w = r), axis 2 -> (h = r, w = i)
    const float4* p = x + (axis == 1 ? ((int64_t)b * H * W + r) : ((int64_t)b * H + r) * W) * E4 + e;
    const int64_t step = (int64_t)(axis == 1 ? W : 1) * E4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = 0;
    for (; i + 8 <= n; i += 8) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldg(p + (i + j) * step);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s.x += v[j].x; s.y += v[j].y; s.z += v[j].z; s.w += v[j].w; }
    }
    for (; i < n; ++i) {
      const float4 v = __ldg(p + i * step);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
    if (add) { const float4 v = __ldg(add + t); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
    if (accumulate) { const float4 o = out[t]; s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w; }
    out[t] = s;
  }
}

// out[m,:] = a[m,:] + b[m,:] + c[m,:] + sr * row[(b, w),:] + sc * col[(b, h),:]   ([B,H,W,E] fp32; any may be null)
__global__ void combine_bcast_kernel(const float* __restrict__ a, const float* __restrict__ b2,
                                     const float* __restrict__ c, const float* __restrict__ row, float sr,
                                     const float* __restrict__ col, float sc, int64_t M, int E, int H, int W,
                                     float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = M * E;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(t % E);
    const int64_t m = t / E;
    float v = 0.0f;
    if (a) v += a[t];
    if (b2) v += b2[t];
    if (c) v += c[t];
    if (row) v += sr * row[((m / ((int64_t)H * W)) * W + (m % W)) * E + e];
    if (col) v += sc * col[(m / W) * E + e];
    out[t] = v;
  }
}
// same, four channels per thread (E % 4 == 0, 16-byte aligned operands): one index decode per 16 bytes of every stream
__global__ void combine_bcast4_kernel(const float4* __restrict__ a, const float4* __restrict__ b2,
                                      const float4* __restrict__ c, const float4* __restrict__ row, float sr,
                                      const float4* __restrict__ col, float sc, int64_t M, int E4, int H, int W,
                                      float4* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = M * E4;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(t % E4);
    const int64_t m = t / E4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a) { const float4 x = __ldg(a + t); v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w; }
    if (b2) { const float4 x = __ldg(b2 + t); v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w; }
    if (c) { const float4 x = __ldg(c + t); v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w; }
    if (row) {
      const float4 x = __ldg(row + ((m / ((int64_t)H * W)) * W + (m % W)) * E4 + e);
      v.x += sr * x.x; v.y += sr * x.y; v.z += sr * x.z; v.w += sr * x.w;
    }
    if (col) {
      const float4 x = __ldg(col + (m / W) * E4 + e);
      v.x += sc * x.x; v.y += sc * x.y; v.z += sc * x.z; v.w += sc * x.w;
    }
    out[t] = v;
  }
}

// out[n] += sum_m x[m,n]  (bias gradients); x fp32 [M, N] or split.  8 warps stride the rows of a row block,
// each lane owns 8 consecutive columns (16-byte loads per plane); warps are reduced through shared memory and
// the CTA issues one atomic per column.
__global__ void colsum_kernel(const float* __restrict__ x, const __nv_bfloat16* x_hi,
                              const __nv_bfloat16* x_lo, int64_t ld, int64_t M, int N, int rows_per_cta,
                              float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][256 + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(M, r0 + rows_per_cta);
  const int c0 = blockIdx.y * 256 + lane * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool vec = x == nullptr && (ld % 8 == 0) && c0 + 8 <= N &&
                   ((reinterpret_cast<uintptr_t>(x_hi) | reinterpret_cast<uintptr_t>(x_lo)) & 15) == 0;
  if (vec) {
    for (int64_t r = r0 + warp; r < r1; r += 8) {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(x_hi + r * ld + c0));
      const uint4 l = __ldg(reinterpret_cast<const uint4*>(x_lo + r * ld + c0));
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] += bf16_bits_to_float(hw[j] & 0xffffu) + bf16_bits_to_float(lw[j] & 0xffffu);
        acc[2 * j + 1] += bf16_bits_to_float(hw[j] >> 16) + bf16_bits_to_float(lw[j] >> 16);
      }
    }
  } else {
    for (int64_t r = r0 + warp; r < r1; r += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c0 + j < N)
          acc[j] += x ? x[r * ld + c0 + j] : join_bf16(x_hi[r * ld + c0 + j], x_lo[r * ld + c0 + j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c < N) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(out + c, s);
  }
}

// Box head epilogue: t [M,4] raw MLP output, ref [M,2] anchor points -> boxes = sigmoid(t + [inv_sig(ref),0,0])
__device__ __forceinline__ float inv_sigmoid(float x) {
  x = fminf(fmaxf(x, 0.0f), 1.0f);
  const float x1 = fmaxf(x, 1e-5f), x2 = fmaxf(1.0f - x, 1e-5f);
  return logf(x1 / x2);
}
__global__ void box_head_fwd_kernel(const float* __restrict__ t, const float* __restrict__ ref, int64_t M,
                                    float* __restrict__ boxes) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M * 4) return;
  const int64_t m = i / 4;
  const int c = (int)(i % 4);
  float v = t[i];
  if (c < 2) v += inv_sigmoid(ref[m * 2 + c]);
  boxes[i] = 1.0f / (1.0f + expf(-v));
}
// dt = dboxes * s * (1 - s);  dref[m,c] += dt * d inv_sigmoid/dx  (zero outside the clamps)
__global__ void box_head_bwd_kernel(const float* __restrict__ dboxes, const float* __restrict__ boxes,
                                    const float* __restrict__ ref, int64_t M, float* __restrict__ dt,
                                    __nv_bfloat16* dt_hi, __nv_bfloat16* dt_lo, int64_t ld_split,
                                    float* __restrict__ dref) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M * 4) return;
  const int64_t m = i / 4;
  const int c = (int)(i % 4);
  const float s = boxes[i];
  const float g = dboxes[i] * s * (1.0f - s);
  if (dt) dt[i] = g;
  if (dt_hi) split_bf16(g, dt_hi[m * ld_split + c], dt_lo[m * ld_split + c]);
  if (dref && c < 2) {
    const float x = ref[m * 2 + c];
    float d = 0.0f;
    if (x >= 0.0f && x <= 1.0f) {
      // log(max(x,eps)) - log(max(1-x,eps))
      if (x > 1e-5f) d += 1.0f / x;
      if (1.0f - x > 1e-5f) d += 1.0f / (1.0f - x);
    }
    atomicAdd(dref + m * 2 + c, g * d);
  }
}

// y = relu'(mask) * x as split (mask fp32 or via >0 of saved activations) -- used for MLP/FFN hidden grads
__global__ void scale_kernel(float* __restrict__ x, int64_t n, float s) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    x[i] *= s;
}

inline int grid_for(int64_t n, int threads = 256) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = 148LL * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define SPLIT_HI(t) reinterpret_cast<__nv_bfloat16*>((t).base)
#define SPLIT_LO(t) ((t).base ? reinterpret_cast<__nv_bfloat16*>((t).base) + (t).plane : nullptr)

extern "C" int cdetr_sine_embed(const float* pos, int64_t n, int pos_stride, int num_feats, int off,
                                int ld, float* out, cdetr_stream_t s) {
  CDETR_CHECK_ARG(pos && out && n > 0 && num_feats > 0 && off + num_feats <= ld, "sine_embed: bad args");
  launch_light(sine_embed_kernel, dim3(grid_for(n * num_feats)), dim3(256), 0, STREAM(s), pos, n, pos_stride, num_feats, off, ld, out);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_sine_embed_bwd(const float* pos, int64_t n, int pos_stride, int num_feats, int off,
                                    int ld, const float* demb, float* dpos, cdetr_stream_t s) {
  CDETR_CHECK_ARG(pos && demb && dpos && n > 0, "sine_embed_bwd: bad args");
  launch_light(sine_embed_bwd_kernel, dim3(cdiv(n, 8)), dim3(256), 0, STREAM(s), pos, n, pos_stride, num_feats, off, ld, demb, dpos);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_add_bcast(const float* x, const float* y, int64_t M, int E, int mode, int H, int W,
                               int64_t rows_y, float* out, cdetr_split_t out_split, cdetr_stream_t s) {
  CDETR_CHECK_ARG(x && M > 0 && E % 8 == 0 && mode >= 0 && mode <= 3, "add_bcast: bad args");
  launch_light(add_bcast_kernel, dim3(grid_for(M * (E / 8))), dim3(256), 0, STREAM(s), x, y, M, E, mode, H, W, rows_y, out,
                                                                SPLIT_HI(out_split), SPLIT_LO(out_split),
                                                                out_split.ld);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_reduce_axis(const float* x, int B, int H, int W, int E, int axis, float scale,
                                 const float* add, int accumulate, float* out, cdetr_split_t out_split,
                                 cdetr_stream_t s) {
  CDETR_CHECK_ARG(x && (axis == 1 || axis == 2) && (out || out_split.base), "reduce_axis: bad args");
  const int R = axis == 1 ? W : H;
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(add) | reinterpret_cast<uintptr_t>(out);
  static const int scalar_light = getenv("CDETR_SCALAR_LIGHT") ? atoi(getenv("CDETR_SCALAR_LIGHT")) : 0;   // A/B: 2 = scalar reduce_axis
  if (!(scalar_light & 2) && out && !out_split.base && E % 4 == 0 && (al & 15) == 0) {
    launch_light(reduce_axis4_kernel, dim3(grid_for((int64_t)B * R * (E / 4), 128)), dim3(128), 0, STREAM(s),
                 reinterpret_cast<const float4*>(x), B, H, W, E / 4, axis, scale, reinterpret_cast<const float4*>(add), accumulate,
                 reinterpret_cast<float4*>(out));
    CDETR_CHECK_LAUNCH();
    return 0;
  }
  launch_light(reduce_axis_kernel, dim3(grid_for((int64_t)B * R * E)), dim3(256), 0, STREAM(s), 
      x, B, H, W, E, axis, scale, add, accumulate, out, SPLIT_HI(out_split), SPLIT_LO(out_split), out_split.ld);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_combine_bcast(const float* a, const float* b, const float* c, const float* row,
                                   float sr, const float* col, float sc, int64_t M, int E, int H, int W,
                                   float* out, cdetr_stream_t s) {
  CDETR_CHECK_ARG(out && M > 0, "combine_bcast: bad args");
  const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                       reinterpret_cast<uintptr_t>(row) | reinterpret_cast<uintptr_t>(col) | reinterpret_cast<uintptr_t>(out);
  static const int scalar_light = getenv("CDETR_SCALAR_LIGHT") ? atoi(getenv("CDETR_SCALAR_LIGHT")) : 0;   // A/B: 1 = scalar combine_bcast
  if (!(scalar_light & 1) && E % 4 == 0 && (al & 15) == 0) {
    launch_light(combine_bcast4_kernel, dim3(grid_for(M * (E / 4))), dim3(256), 0, STREAM(s), reinterpret_cast<const float4*>(a),
                 reinterpret_cast<const float4*>(b), reinterpret_cast<const float4*>(c), reinterpret_cast<const float4*>(row), sr,
                 reinterpret_cast<const float4*>(col), sc, M, E / 4, H, W, reinterpret_cast<float4*>(out));
  } else {
    launch_light(combine_bcast_kernel, dim3(grid_for(M * E)), dim3(256), 0, STREAM(s), a, b, c, row, sr, col, sc, M, E, H, W, out);
  }
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_colsum(const float* x, cdetr_split_t x_split, int64_t ld, int64_t M, int N, float* out,
                            cdetr_stream_t s) {
  CDETR_CHECK_ARG((x || x_split.base) && out && M > 0 && N > 0, "colsum: bad args");
  int rows_per_cta = 512;
  while (rows_per_cta > 64 && cdiv(M, rows_per_cta) * cdiv(N, 256) < 2 * 148) rows_per_cta >>= 1;
  dim3 grid(cdiv(M, rows_per_cta), cdiv(N, 256));
  launch_light(colsum_kernel, dim3(grid), dim3(256), 0, STREAM(s), x, SPLIT_HI(x_split), SPLIT_LO(x_split),
                                             x ? ld : x_split.ld, M, N, rows_per_cta, out);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_box_head_fwd(const float* t, const float* ref, int64_t M, float* boxes,
                                  cdetr_stream_t s) {
  CDETR_CHECK_ARG(t && ref && boxes && M > 0, "box_head_fwd: bad args");
  launch_light(box_head_fwd_kernel, dim3(cdiv(M * 4, 256)), dim3(256), 0, STREAM(s), t, ref, M, boxes);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_box_head_bwd(const float* dboxes, const float* boxes, const float* ref, int64_t M,
                                  float* dt, cdetr_split_t dt_split, float* dref, cdetr_stream_t s) {
  CDETR_CHECK_ARG(dboxes && boxes && ref && M > 0, "box_head_bwd: bad args");
  launch_light(box_head_bwd_kernel, dim3(cdiv(M * 4, 256)), dim3(256), 0, STREAM(s), dboxes, boxes, ref, M, dt, SPLIT_HI(dt_split),
                                                             SPLIT_LO(dt_split), dt_split.ld, dref);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_scale(float* x, int64_t n, float a, cdetr_stream_t s) {
  CDETR_CHECK_ARG(x && n > 0, "scale: bad args");
  launch_light(scale_kernel, dim3(grid_for(n)), dim3(256), 0, STREAM(s), x, n, a);
  CDETR_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------ padding mask -> feature-level masks + positions
// A2/models/backbone.py:112 (F.interpolate(m[None].float(), size=x.shape[-2:]).to(bool)[0]: nearest, source index
// floor(dst * in/out) in fp32, clamped), A2/models/transformer.py:497-503 (mask2pos: positions from the FIRST column /
// FIRST row of the feature mask: (cumsum(~m) - 0.5) / count), A2/models/row_column_decoupled_attention.py:238-249
// (key padding taken from mask[:, 0, :] / mask[:, :, 0]).  One CTA per sample; H, W <= 1024.
namespace {
__global__ void mask_prepare_kernel(const uint8_t* __restrict__ mask, int S1, int S2, int H, int W,
                                    uint8_t* __restrict__ mrow, uint8_t* __restrict__ mcol, float* __restrict__ pos_row,
                                    float* __restrict__ pos_col) {
  __shared__ float cnt[2];
  const int b = blockIdx.x;
  const uint8_t* m = mask + (int64_t)b * S1 * S2;
  const float sy = (float)S1 / (float)H, sx = (float)S2 / (float)W;
  // first feature row -> row mask [W]; first feature column -> column mask [H]  (source row / column 0)
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const int xs = min((int)floorf((float)x * sx), S2 - 1);
    mrow[b * W + x] = m[xs] != 0;
  }
  for (int y = threadIdx.x; y < H; y += blockDim.x) {
    const int ys = min((int)floorf((float)y * sy), S1 - 1);
    mcol[b * H + y] = m[(int64_t)ys * S2] != 0;
  }
  __syncthreads();
  if (threadIdx.x < 2) {   // serial fp32 cumsum of <= 1024 ones: exact integers, same values as torch.cumsum
    const int n = threadIdx.x == 0 ? W : H;
    const uint8_t* mk = threadIdx.x == 0 ? mrow + b * W : mcol + b * H;
    float* out = threadIdx.x == 0 ? pos_row + b * W : pos_col + b * H;
    float c = 0.0f;
    for (int i = 0; i < n; ++i) {
      c += mk[i] ? 0.0f : 1.0f;
      out[i] = c;
    }
    cnt[threadIdx.x] = c;
  }
  __syncthreads();
  for (int x = threadIdx.x; x < W; x += blockDim.x) pos_row[b * W + x] = __fdiv_rn(__fsub_rn(pos_row[b * W + x], 0.5f), cnt[0]);
  for (int y = threadIdx.x; y < H; y += blockDim.x) pos_col[b * H + y] = __fdiv_rn(__fsub_rn(pos_col[b * H + y], 0.5f), cnt[1]);
}

// A2/models/backbone.py:122-128: centre = int((x1*W + x2*W) / 2) in fp32 (truncation), rects of SAMPLE 0 only.
// Out-of-range centres (the reference raises IndexError for >= 1 and wraps negatives) are clamped and flagged.
__global__ void exemplar_centres_kernel(const float* __restrict__ rects0, int n_ex, int H, int W, int* __restrict__ yx,
                                        int* __restrict__ status) {
  const int k = threadIdx.x;
  if (k >= n_ex) return;
  const float x1 = rects0[4 * k], y1 = rects0[4 * k + 1], x2 = rects0[4 * k + 2], y2 = rects0[4 * k + 3];
  const float fx = __fdiv_rn(__fadd_rn(__fmul_rn(x1, (float)W), __fmul_rn(x2, (float)W)), 2.0f);
  const float fy = __fdiv_rn(__fadd_rn(__fmul_rn(y1, (float)H), __fmul_rn(y2, (float)H)), 2.0f);
  int xc = (int)fx, yc = (int)fy;   // cvt.rzi: truncation toward zero like Python int()
  if (xc < 0 || xc >= W || yc < 0 || yc >= H || !(fx == fx) || !(fy == fy)) {
    if (status) atomicOr(status, 2);
    xc = min(max(xc, 0), W - 1);
    yc = min(max(yc, 0), H - 1);
  }
  yx[2 * k] = yc;
  yx[2 * k + 1] = xc;
}
}  // namespace

extern "C" int cdetr_mask_prepare(const uint8_t* mask, int B, int S1, int S2, int H, int W, uint8_t* mask_row,
                                  uint8_t* mask_col, float* pos_row, float* pos_col, cdetr_stream_t s) {
  CDETR_CHECK_ARG(mask && mask_row && mask_col && pos_row && pos_col && B > 0 && H > 0 && W > 0 && S1 >= H && S2 >= W,
                  "mask_prepare: bad args");
  mask_prepare_kernel<<<B, 128, 0, STREAM(s)>>>(mask, S1, S2, H, W, mask_row, mask_col, pos_row, pos_col);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_exemplar_centres(const float* rects0, int n_ex, int H, int W, int* centres_yx, int* status,
                                      cdetr_stream_t s) {
  CDETR_CHECK_ARG(rects0 && centres_yx && n_ex > 0 && n_ex <= 32 && H > 0 && W > 0, "exemplar_centres: bad args");
  exemplar_centres_kernel<<<1, 32, 0, STREAM(s)>>>(rects0, n_ex, H, W, centres_yx, status);
  CDETR_CHECK_LAUNCH();
  return 0;
}
