// HBM-bound layout kernels around the GEMM for the ResNet-50 (DC5) backbone, all on split-bf16 NHWC
// activations ([2 planes][B*H*W][C], 16-byte vectors of 8 channels per thread, coalesced over C):
//   weight packing (FrozenBN scale folded, tap-major K order, optional transpose for dgrad),
//   im2col / col2im for the 3x3 convs (any stride / dilation), the 7x7 stem gather, 3x3 s2 max-pool,
//   stride-2 subsample / zero-upsample for the strided 1x1 downsample convs.
// Reference op sites: A2/models/resnet.py:143-158,263-271; A2/models/backbone.py:50-60.
#include "common.cuh"
#include "../../include/cdetr.h"

namespace {

struct V8 {  // 8 fp32 values <-> one 16-byte vector per plane
  float v[8];
};

__device__ __forceinline__ V8 load_join8(const __nv_bfloat16* hi, const __nv_bfloat16* lo) {
  const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi));
  const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo));
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
  const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
  V8 r;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    r.v[2 * j] = bf16_bits_to_float(hw[j] & 0xffffu) + bf16_bits_to_float(lw[j] & 0xffffu);
    r.v[2 * j + 1] = bf16_bits_to_float(hw[j] >> 16) + bf16_bits_to_float(lw[j] >> 16);
  }
  return r;
}

__device__ __forceinline__ void store_split8(__nv_bfloat16* hi, __nv_bfloat16* lo, const V8& x) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    split_bf16_pair(x.v[2 * j], x.v[2 * j + 1], hw[j], lw[j]);
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  *reinterpret_cast<uint4*>(lo) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

struct SplitPtr {
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  int64_t ld;
};
inline SplitPtr sp(const cdetr_split_t& t) {
  SplitPtr p;
  p.hi = reinterpret_cast<__nv_bfloat16*>(t.base);
  p.lo = p.hi ? p.hi + t.plane : nullptr;
  p.ld = t.ld;
  return p;
}

// ------------------------------------------------------------------ weights
__global__ void bn_fold_kernel(const float* w, const float* b, const float* rm, const float* rv,
                               float eps, int c, float* scale, float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float s = w[i] * rsqrtf(rv[i] + eps);
  scale[i] = s;
  shift[i] = b[i] - rm[i] * s;
}

// w [cout, cin, taps] fp32  ->  dst [cout, taps*cin] and/or dst_t [taps*cin, cout], scaled per cout.
__global__ void pack_weight_kernel(const float* __restrict__ w, int cout, int cin, int taps,
                                   const float* __restrict__ row_scale, SplitPtr dst, SplitPtr dst_t) {
  const int64_t n = (int64_t)cout * cin * taps;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const int c = (int)((i / taps) % cin);
    const int o = (int)(i / ((int64_t)taps * cin));
    float x = w[i];
    if (row_scale) x *= row_scale[o];
    __nv_bfloat16 h, l;
    split_bf16(x, h, l);
    const int64_t k = (int64_t)t * cin + c;
    if (dst.hi) {
      dst.hi[(int64_t)o * dst.ld + k] = h;
      dst.lo[(int64_t)o * dst.ld + k] = l;
    }
    if (dst_t.hi) {
      dst_t.hi[k * dst_t.ld + o] = h;
      dst_t.lo[k * dst_t.ld + o] = l;
    }
  }
}

// dst[cin, taps*cout]: element (c, t*cout + o) = w[o, c, t] * row_scale[o]   (implicit-conv dgrad weights)
__global__ void pack_weight_dgrad_kernel(const float* __restrict__ w, int cout, int cin, int taps,
                                         const float* __restrict__ row_scale, SplitPtr dst) {
  const int64_t n = (int64_t)cout * cin * taps;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const int c = (int)((i / taps) % cin);
    const int o = (int)(i / ((int64_t)taps * cin));
    float x = w[i];
    if (row_scale) x *= row_scale[o];
    __nv_bfloat16 h, l;
    split_bf16(x, h, l);
    const int64_t k = (int64_t)t * cout + o;
    dst.hi[(int64_t)c * dst.ld + k] = h;
    dst.lo[(int64_t)c * dst.ld + k] = l;
  }
}

// every weight of the model in one launch: block table (entry, chunk), see cdetr_mt_pack_weights in cdetr.h
__global__ void __launch_bounds__(256)
mt_pack_weights_kernel(const cdetr_pack_entry_t* __restrict__ table, const int32_t* __restrict__ blocks, int chunk,
                       float eps) {
  const int t = blocks[2 * blockIdx.x], ck = blocks[2 * blockIdx.x + 1];
  const cdetr_pack_entry_t e = table[t];
  const int64_t n = (int64_t)e.cout * e.cin * e.taps;
  const int64_t begin = (int64_t)ck * chunk, end = min(n, begin + chunk);
  const bool bn = e.bn_w != nullptr;
  if (bn && ck == 0) {
    for (int o = threadIdx.x; o < e.cout; o += blockDim.x) {
      const float sc = e.bn_w[o] * rsqrtf(e.bn_rv[o] + eps);      // same expression as bn_fold_kernel
      e.scale[o] = sc;
      e.shift[o] = e.bn_b[o] - e.bn_rm[o] * sc;
    }
  }
  __nv_bfloat16* d_hi = reinterpret_cast<__nv_bfloat16*>(e.dst.base);
  __nv_bfloat16* t_hi = reinterpret_cast<__nv_bfloat16*>(e.dst_t.base);
  __nv_bfloat16* g_hi = reinterpret_cast<__nv_bfloat16*>(e.dst_d.base);
  for (int64_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const int tp = (int)(i % e.taps);
    const int c = (int)((i / e.taps) % e.cin);
    const int o = (int)(i / ((int64_t)e.taps * e.cin));
    float x = e.w[i];
    if (bn) x *= e.bn_w[o] * rsqrtf(e.bn_rv[o] + eps);
    __nv_bfloat16 h, l;
    split_bf16(x, h, l);
    const int64_t k = (int64_t)tp * e.cin + c;
    if (d_hi) {
      d_hi[(int64_t)o * e.dst.ld + k] = h;
      d_hi[e.dst.plane + (int64_t)o * e.dst.ld + k] = l;
    }
    if (t_hi) {
      t_hi[k * e.dst_t.ld + o] = h;
      t_hi[e.dst_t.plane + k * e.dst_t.ld + o] = l;
    }
    if (g_hi) {
      const int64_t kd = (int64_t)tp * e.cout + o;
      g_hi[(int64_t)c * e.dst_d.ld + kd] = h;
      g_hi[e.dst_d.plane + (int64_t)c * e.dst_d.ld + kd] = l;
    }
  }
}

// grad[cout, cin, taps] += g[cout, taps*cin]
__global__ void unpack_conv_grad_kernel(const float* __restrict__ g, int cout, int cin, int taps,
                                        float* __restrict__ grad) {
  const int64_t n = (int64_t)cout * cin * taps;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const int c = (int)((i / taps) % cin);
    const int o = (int)(i / ((int64_t)taps * cin));
    grad[i] += g[(int64_t)o * taps * cin + (int64_t)t * cin + c];
  }
}

// ------------------------------------------------------------------ fp32 <-> split
__global__ void to_split_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t ld_x,
                                SplitPtr dst) {
  const int c8n = (cols + 7) / 8;
  const int64_t n = rows * c8n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c8n;
    const int c0 = (int)(i % c8n) * 8;
    V8 v;
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] = (c0 + j < cols) ? x[r * ld_x + c0 + j] : 0.0f;
    store_split8(dst.hi + r * dst.ld + c0, dst.lo + r * dst.ld + c0, v);
  }
}

__global__ void from_split_kernel(SplitPtr src, int64_t rows, int cols, float* __restrict__ y,
                                  int64_t ld_y) {
  const int64_t n = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i % cols);
    y[r * ld_y + c] = join_bf16(src.hi[r * src.ld + c], src.lo[r * src.ld + c]);
  }
}

// ------------------------------------------------------------------ stem: 7x7 s2 p3 gather from NCHW fp32
// col[m, (r*7+s)*3 + c], m = (b, oy, ox); K = 147 (ld >= 152, tail columns zeroed).
// One CTA per (image, output row, 64 output columns): the 7 x 3 x 133 input window is staged in shared memory with
// coalesced row reads, then every thread assembles 8-column vectors of the im2col rows from it (the direct gather
// version spent 460 us on 8 scattered loads per thread: 1.5 TB/s; this one is bound by the 637 MB it writes).
constexpr int STEM_PX = 64;
constexpr int STEM_TW = 2 * STEM_PX + 5;   // input columns touched by 64 stride-2 outputs of a 7-wide filter
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ img, int B, int H, int W, int Ho,
                                                          int Wo, SplitPtr col) {
  __shared__ float tile[3][7][STEM_TW + 3];
  const int segs = (Wo + STEM_PX - 1) / STEM_PX;
  const int seg = blockIdx.x % segs;
  const int oy = (blockIdx.x / segs) % Ho;
  const int b = blockIdx.x / (segs * Ho);
  const int ox0 = seg * STEM_PX;
  for (int i = threadIdx.x; i < 3 * 7 * STEM_TW; i += 256) {
    const int x = i % STEM_TW, r = (i / STEM_TW) % 7, c = i / (7 * STEM_TW);
    const int iy = oy * 2 - 3 + r, ix = ox0 * 2 - 3 + x;
    float v = 0.0f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(img + (((int64_t)b * 3 + c) * H + iy) * W + ix);
    tile[c][r][x] = v;
  }
  __syncthreads();
  const int npx = min(STEM_PX, Wo - ox0);
  for (int i = threadIdx.x; i < npx * 19; i += 256) {
    const int v8 = i % 19, p = i / 19;
    V8 v;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = v8 * 8 + j;
      float x = 0.0f;
      if (k < 147) {
        const int c = k % 3, t = k / 3, s_ = t % 7, r = t / 7;
        x = tile[c][r][2 * p + s_];
      }
      v.v[j] = x;
    }
    const int64_t m = ((int64_t)b * Ho + oy) * Wo + ox0 + p;
    store_split8(col.hi + m * col.ld + v8 * 8, col.lo + m * col.ld + v8 * 8, v);
  }
}

// ------------------------------------------------------------------ 3x3 im2col / col2im (pad = dilation)
__global__ void im2col3x3_kernel(SplitPtr x, int B, int H, int W, int C, int Ho, int Wo, int stride,
                                 int dil, SplitPtr col) {
  const int c8n = C / 8;
  const int64_t n = (int64_t)B * Ho * Wo * 9 * c8n;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % c8n) * 8;
    const int t = (int)((i / c8n) % 9);
    const int64_t m = i / ((int64_t)c8n * 9);
    const int ox = (int)(m % Wo);
    const int oy = (int)((m / Wo) % Ho);
    const int b = (int)(m / ((int64_t)Wo * Ho));
    const int iy = oy * stride - dil + (t / 3) * dil;
    const int ix = ox * stride - dil + (t % 3) * dil;
    uint4 h = zero, l = zero;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const int64_t src = (((int64_t)b * H + iy) * W + ix) * x.ld + c0;
      h = __ldg(reinterpret_cast<const uint4*>(x.hi + src));
      l = __ldg(reinterpret_cast<const uint4*>(x.lo + src));
    }
    const int64_t dst = m * col.ld + (int64_t)t * C + c0;
    *reinterpret_cast<uint4*>(col.hi + dst) = h;
    *reinterpret_cast<uint4*>(col.lo + dst) = l;
  }
}

// dx[b,iy,ix,:] = sum over taps of dcol[(b,oy,ox), tap, :] with iy = oy*stride - dil + r*dil (gather form),
// then optional ReLU mask (mask.hi > 0).
__global__ void col2im3x3_kernel(SplitPtr dcol, int B, int H, int W, int C, int Ho, int Wo, int stride,
                                 int dil, SplitPtr mask, SplitPtr dx) {
  const int c8n = C / 8;
  const int64_t n = (int64_t)B * H * W * c8n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % c8n) * 8;
    const int64_t m = i / c8n;
    const int ix = (int)(m % W);
    const int iy = (int)((m / W) % H);
    const int b = (int)(m / ((int64_t)W * H));
    V8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = 0.0f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int ny = iy + dil - (t / 3) * dil;
      const int nx = ix + dil - (t % 3) * dil;
      if (ny < 0 || nx < 0 || (ny % stride) != 0 || (nx % stride) != 0) continue;
      const int oy = ny / stride, ox = nx / stride;
      if (oy >= Ho || ox >= Wo) continue;
      const int64_t src = (((int64_t)b * Ho + oy) * Wo + ox) * dcol.ld + (int64_t)t * C + c0;
      const V8 v = load_join8(dcol.hi + src, dcol.lo + src);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc.v[j] += v.v[j];
    }
    if (mask.hi) {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(mask.hi + m * mask.ld + c0));
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!(bf16_bits_to_float(hw[j] & 0xffffu) > 0.0f)) acc.v[2 * j] = 0.0f;
        if (!(bf16_bits_to_float(hw[j] >> 16) > 0.0f)) acc.v[2 * j + 1] = 0.0f;
      }
    }
    store_split8(dx.hi + m * dx.ld + c0, dx.lo + m * dx.ld + c0, acc);
  }
}

// ------------------------------------------------------------------ max-pool 3x3 s2 p1
__global__ void maxpool3x3s2_kernel(SplitPtr x, int B, int H, int W, int C, int Ho, int Wo, SplitPtr y) {
  const int c8n = C / 8;
  const int64_t n = (int64_t)B * Ho * Wo * c8n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % c8n) * 8;
    const int64_t m = i / c8n;
    const int ox = (int)(m % Wo);
    const int oy = (int)((m / Wo) % Ho);
    const int b = (int)(m / ((int64_t)Wo * Ho));
    V8 best;
#pragma unroll
    for (int j = 0; j < 8; ++j) best.v[j] = -INFINITY;
    for (int r = 0; r < 3; ++r) {
      const int iy = oy * 2 - 1 + r;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int ix = ox * 2 - 1 + s;
        if (ix < 0 || ix >= W) continue;
        const int64_t src = (((int64_t)b * H + iy) * W + ix) * x.ld + c0;
        const V8 v = load_join8(x.hi + src, x.lo + src);
#pragma unroll
        for (int j = 0; j < 8; ++j) best.v[j] = fmaxf(best.v[j], v.v[j]);
      }
    }
    store_split8(y.hi + m * y.ld + c0, y.lo + m * y.ld + c0, best);
  }
}

// ------------------------------------------------------------------ stride-2 subsample and its adjoint
__global__ void subsample2_kernel(SplitPtr x, int B, int H, int W, int C, int Ho, int Wo, SplitPtr y) {
  const int c8n = C / 8;
  const int64_t n = (int64_t)B * Ho * Wo * c8n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % c8n) * 8;
    const int64_t m = i / c8n;
    const int ox = (int)(m % Wo);
    const int oy = (int)((m / Wo) % Ho);
    const int b = (int)(m / ((int64_t)Wo * Ho));
    const int64_t src = (((int64_t)b * H + oy * 2) * W + ox * 2) * x.ld + c0;
    *reinterpret_cast<uint4*>(y.hi + m * y.ld + c0) = __ldg(reinterpret_cast<const uint4*>(x.hi + src));
    *reinterpret_cast<uint4*>(y.lo + m * y.ld + c0) = __ldg(reinterpret_cast<const uint4*>(x.lo + src));
  }
}

__global__ void upsample2_zero_kernel(SplitPtr dy, int B, int H, int W, int C, int Ho, int Wo, SplitPtr dx) {
  const int c8n = C / 8;
  const int64_t n = (int64_t)B * H * W * c8n;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % c8n) * 8;
    const int64_t m = i / c8n;
    const int ix = (int)(m % W);
    const int iy = (int)((m / W) % H);
    const int b = (int)(m / ((int64_t)W * H));
    uint4 h = zero, l = zero;
    if ((iy & 1) == 0 && (ix & 1) == 0 && iy / 2 < Ho && ix / 2 < Wo) {
      const int64_t src = (((int64_t)b * Ho + iy / 2) * Wo + ix / 2) * dy.ld + c0;
      h = __ldg(reinterpret_cast<const uint4*>(dy.hi + src));
      l = __ldg(reinterpret_cast<const uint4*>(dy.lo + src));
    }
    *reinterpret_cast<uint4*>(dx.hi + m * dx.ld + c0) = h;
    *reinterpret_cast<uint4*>(dx.lo + m * dx.ld + c0) = l;
  }
}

inline int grid_for(int64_t n, int threads = 256) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = 148LL * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int cdetr_bn_fold(const float* w, const float* b, const float* rm, const float* rv,
                             float eps, int c, float* scale, float* shift, cdetr_stream_t s) {
  CDETR_CHECK_ARG(w && b && rm && rv && scale && shift && c > 0, "bn_fold: bad args");
  bn_fold_kernel<<<cdiv(c, 256), 256, 0, STREAM(s)>>>(w, b, rm, rv, eps, c, scale, shift);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_pack_weight(const float* w, int cout, int cin, int taps, const float* row_scale,
                                 cdetr_split_t dst, cdetr_split_t dst_t, cdetr_stream_t s) {
  CDETR_CHECK_ARG(w && cout > 0 && cin > 0 && taps > 0, "pack_weight: bad args");
  CDETR_CHECK_ARG(dst.base || dst_t.base, "pack_weight: no destination");
  pack_weight_kernel<<<grid_for((int64_t)cout * cin * taps), 256, 0, STREAM(s)>>>(
      w, cout, cin, taps, row_scale, sp(dst), sp(dst_t));
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_pack_weight_dgrad(const float* w, int cout, int cin, int taps, const float* row_scale,
                                       cdetr_split_t dst, cdetr_stream_t s) {
  CDETR_CHECK_ARG(w && cout > 0 && cin > 0 && taps > 0 && dst.base, "pack_weight_dgrad: bad args");
  CDETR_CHECK_ARG(dst.ld >= (int64_t)taps * cout, "pack_weight_dgrad: dst ld too small");
  pack_weight_dgrad_kernel<<<grid_for((int64_t)cout * cin * taps), 256, 0, STREAM(s)>>>(
      w, cout, cin, taps, row_scale, sp(dst));
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_mt_pack_weights(const cdetr_pack_entry_t* table, const int32_t* blocks, int nblocks, int chunk_elems,
                                     float eps, cdetr_stream_t s) {
  CDETR_CHECK_ARG(table && blocks && nblocks > 0 && chunk_elems > 0, "mt_pack_weights: bad args");
  mt_pack_weights_kernel<<<nblocks, 256, 0, STREAM(s)>>>(table, blocks, chunk_elems, eps);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_unpack_conv_grad(const float* g, int cout, int cin, int taps, float* grad,
                                      cdetr_stream_t s) {
  CDETR_CHECK_ARG(g && grad, "unpack_conv_grad: bad args");
  unpack_conv_grad_kernel<<<grid_for((int64_t)cout * cin * taps), 256, 0, STREAM(s)>>>(g, cout, cin,
                                                                                      taps, grad);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_to_split(const float* x, int64_t rows, int cols, int64_t ld_x, cdetr_split_t dst,
                              cdetr_stream_t s) {
  CDETR_CHECK_ARG(x && dst.base && rows > 0 && cols > 0, "to_split: bad args");
  CDETR_CHECK_ARG(dst.ld % 8 == 0 && dst.ld >= ((cols + 7) / 8) * 8, "to_split: dst ld too small");
  to_split_kernel<<<grid_for(rows * ((cols + 7) / 8)), 256, 0, STREAM(s)>>>(x, rows, cols, ld_x, sp(dst));
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_from_split(cdetr_split_t src, int64_t rows, int cols, float* y, int64_t ld_y,
                                cdetr_stream_t s) {
  CDETR_CHECK_ARG(y && src.base && rows > 0 && cols > 0, "from_split: bad args");
  from_split_kernel<<<grid_for(rows * cols), 256, 0, STREAM(s)>>>(sp(src), rows, cols, y, ld_y);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_stem_im2col(const float* img, int B, int H, int W, cdetr_split_t col,
                                 cdetr_stream_t s) {
  CDETR_CHECK_ARG(img && col.base && col.ld >= 152 && col.ld % 8 == 0, "stem_im2col: bad args");
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  stem_im2col_kernel<<<B * Ho * ((Wo + STEM_PX - 1) / STEM_PX), 256, 0, STREAM(s)>>>(img, B, H, W, Ho, Wo, sp(col));
  CDETR_CHECK_LAUNCH();
  return 0;
}

static inline int conv_out(int n, int stride) { return (n - 1) / stride + 1; }  // 3x3, pad = dilation

extern "C" int cdetr_im2col3x3(cdetr_split_t x, int B, int H, int W, int C, int stride, int dil,
                               cdetr_split_t col, cdetr_stream_t s) {
  CDETR_CHECK_ARG(x.base && col.base && C % 8 == 0 && col.ld >= 9 * C, "im2col3x3: bad args");
  const int Ho = conv_out(H, stride), Wo = conv_out(W, stride);
  im2col3x3_kernel<<<grid_for((int64_t)B * Ho * Wo * 9 * (C / 8)), 256, 0, STREAM(s)>>>(
      sp(x), B, H, W, C, Ho, Wo, stride, dil, sp(col));
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_col2im3x3(cdetr_split_t dcol, int B, int H, int W, int C, int stride, int dil,
                               cdetr_split_t mask, cdetr_split_t dx, cdetr_stream_t s) {
  CDETR_CHECK_ARG(dcol.base && dx.base && C % 8 == 0, "col2im3x3: bad args");
  const int Ho = conv_out(H, stride), Wo = conv_out(W, stride);
  col2im3x3_kernel<<<grid_for((int64_t)B * H * W * (C / 8)), 256, 0, STREAM(s)>>>(
      sp(dcol), B, H, W, C, Ho, Wo, stride, dil, sp(mask), sp(dx));
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_maxpool3x3s2(cdetr_split_t x, int B, int H, int W, int C, cdetr_split_t y,
                                  cdetr_stream_t s) {
  CDETR_CHECK_ARG(x.base && y.base && C % 8 == 0, "maxpool: bad args");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  maxpool3x3s2_kernel<<<grid_for((int64_t)B * Ho * Wo * (C / 8)), 256, 0, STREAM(s)>>>(
      sp(x), B, H, W, C, Ho, Wo, sp(y));
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_subsample2(cdetr_split_t x, int B, int H, int W, int C, cdetr_split_t y,
                                cdetr_stream_t s) {
  CDETR_CHECK_ARG(x.base && y.base && C % 8 == 0, "subsample2: bad args");
  const int Ho = conv_out(H, 2), Wo = conv_out(W, 2);
  subsample2_kernel<<<grid_for((int64_t)B * Ho * Wo * (C / 8)), 256, 0, STREAM(s)>>>(sp(x), B, H, W, C,
                                                                                    Ho, Wo, sp(y));
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_upsample2_zero(cdetr_split_t dy, int B, int H, int W, int C, cdetr_split_t dx,
                                    cdetr_stream_t s) {
  CDETR_CHECK_ARG(dy.base && dx.base && C % 8 == 0, "upsample2_zero: bad args");
  const int Ho = conv_out(H, 2), Wo = conv_out(W, 2);
  upsample2_zero_kernel<<<grid_for((int64_t)B * H * W * (C / 8)), 256, 0, STREAM(s)>>>(
      sp(dy), B, H, W, C, Ho, Wo, sp(dx));
  CDETR_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------ exemplar feature injection
// A2/models/backbone.py:116-136.  p[b,c] = mean_k x[b, yc_k, xc_k, c];  cat[m, :C] = x[m,:], cat[m, C:] = x[m,:]*p[b,:]
namespace {
__global__ void exemplar_p_kernel(SplitPtr x, int B, int H, int W, int C, const int* __restrict__ yx, int n_ex,
                                  float* __restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  float s = 0.0f;
  for (int k = 0; k < n_ex; ++k) {
    const int64_t m = ((int64_t)b * H + yx[2 * k]) * W + yx[2 * k + 1];
    s += join_bf16(x.hi[m * x.ld + c], x.lo[m * x.ld + c]);
  }
  p[i] = s / (float)n_ex;
}
__global__ void exemplar_concat_kernel(SplitPtr x, int B, int HW, int C, const float* __restrict__ p, SplitPtr cat) {
  const int c8n = C / 8;
  const int64_t n = (int64_t)B * HW * c8n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % c8n) * 8;
    const int64_t m = i / c8n;
    const int b = (int)(m / HW);
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(x.hi + m * x.ld + c0));
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(x.lo + m * x.ld + c0));
    *reinterpret_cast<uint4*>(cat.hi + m * cat.ld + c0) = h;
    *reinterpret_cast<uint4*>(cat.lo + m * cat.ld + c0) = l;
    V8 v = load_join8(x.hi + m * x.ld + c0, x.lo + m * x.ld + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] *= p[b * C + c0 + j];
    store_split8(cat.hi + m * cat.ld + C + c0, cat.lo + m * cat.ld + C + c0, v);
  }
}
// dp[b,c] = sum_m dcat[m, C+c] * x[m,c]   (one warp per (b, 8-channel group) -> 8 sums)
__global__ void exemplar_dp_kernel(SplitPtr dcat, SplitPtr x, int B, int HW, int C, float* __restrict__ dp) {
  const int c8n = C / 8;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= B * c8n) return;
  const int b = wid / c8n, c0 = (wid % c8n) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int n = lane; n < HW; n += 32) {
    const int64_t m = (int64_t)b * HW + n;
    const V8 d = load_join8(dcat.hi + m * dcat.ld + C + c0, dcat.lo + m * dcat.ld + C + c0);
    const V8 xv = load_join8(x.hi + m * x.ld + c0, x.lo + m * x.ld + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += d.v[j] * xv.v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float t = warp_sum(acc[j]);
    if (lane == 0) dp[b * C + c0 + j] = t;
  }
}
// dx[m,:] = (dcat[m,:C] + dcat[m,C:]*p[b,:] + [m is a centre] * count * dp[b,:]/n_ex) * mask
__global__ void exemplar_dx_kernel(SplitPtr dcat, const float* __restrict__ p, const float* __restrict__ dp, int B,
                                   int H, int W, int C, const int* __restrict__ yx, int n_ex, SplitPtr mask,
                                   SplitPtr dx) {
  const int c8n = C / 8;
  const int HW = H * W;
  const int64_t n = (int64_t)B * HW * c8n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % c8n) * 8;
    const int64_t m = i / c8n;
    const int b = (int)(m / HW);
    const int pos = (int)(m % HW);
    int cnt = 0;
    for (int k = 0; k < n_ex; ++k) cnt += (yx[2 * k] * W + yx[2 * k + 1]) == pos;
    V8 d1 = load_join8(dcat.hi + m * dcat.ld + c0, dcat.lo + m * dcat.ld + c0);
    const V8 d2 = load_join8(dcat.hi + m * dcat.ld + C + c0, dcat.lo + m * dcat.ld + C + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      d1.v[j] += d2.v[j] * p[b * C + c0 + j];
      if (cnt) d1.v[j] += (float)cnt * dp[b * C + c0 + j] / (float)n_ex;
    }
    if (mask.hi) {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(mask.hi + m * mask.ld + c0));
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!(bf16_bits_to_float(hw[j] & 0xffffu) > 0.0f)) d1.v[2 * j] = 0.0f;
        if (!(bf16_bits_to_float(hw[j] >> 16) > 0.0f)) d1.v[2 * j + 1] = 0.0f;
      }
    }
    store_split8(dx.hi + m * dx.ld + c0, dx.lo + m * dx.ld + c0, d1);
  }
}
}  // namespace

extern "C" int cdetr_exemplar_concat(cdetr_split_t x, int B, int H, int W, int C, const int* centres_yx, int n_ex,
                                     float* p_out, cdetr_split_t cat, cdetr_stream_t s) {
  CDETR_CHECK_ARG(x.base && cat.base && centres_yx && p_out && n_ex > 0 && C % 8 == 0 && cat.ld >= 2 * C,
                  "exemplar_concat: bad args");
  exemplar_p_kernel<<<cdiv((int64_t)B * C, 256), 256, 0, STREAM(s)>>>(sp(x), B, H, W, C, centres_yx, n_ex, p_out);
  CDETR_CHECK_LAUNCH();
  exemplar_concat_kernel<<<grid_for((int64_t)B * H * W * (C / 8)), 256, 0, STREAM(s)>>>(sp(x), B, H * W, C, p_out,
                                                                                     sp(cat));
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_exemplar_concat_bwd(cdetr_split_t dcat, cdetr_split_t x, const float* p, int B, int H, int W,
                                         int C, const int* centres_yx, int n_ex, float* dp_scratch,
                                         cdetr_split_t mask, cdetr_split_t dx, cdetr_stream_t s) {
  CDETR_CHECK_ARG(dcat.base && x.base && p && centres_yx && dp_scratch && dx.base && C % 8 == 0,
                  "exemplar_concat_bwd: bad args");
  exemplar_dp_kernel<<<cdiv((int64_t)B * (C / 8) * 32, 256), 256, 0, STREAM(s)>>>(sp(dcat), sp(x), B, H * W, C,
                                                                               dp_scratch);
  CDETR_CHECK_LAUNCH();
  exemplar_dx_kernel<<<grid_for((int64_t)B * H * W * (C / 8)), 256, 0, STREAM(s)>>>(
      sp(dcat), p, dp_scratch, B, H, W, C, centres_yx, n_ex, sp(mask), sp(dx));
  CDETR_CHECK_LAUNCH();
  return 0;
}
