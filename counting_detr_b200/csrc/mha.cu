// Decoder self-attention core (nn.MultiheadAttention over the Q object queries), forward + backward,
// head dim 32, four threads per query (or key), each walking every 4th key of the K/V (or Q/dO) tiles staged
// in shared memory with an online softmax, partials merged with warp shuffles; the Q x Q logits never touch HBM.  Q <= ~1000 so the whole problem is a few MFLOP per
// (sample, head): latency, not throughput, is what matters here.
// Reference: A2/models/transformer.py:366-372 (decoder self-attention), torch F.multi_head_attention_forward.
#include "common.cuh"
#include "../../include/cdetr.h"
#include "mha_args.cuh"
#include <stdlib.h>

namespace {
constexpr int HD = 32;
constexpr int TK = 128;  // keys / queries per shared-memory tile


__device__ __forceinline__ void load32(const float* p, float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + j);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
__device__ __forceinline__ float dot32(const float* a, const float* smem_row) {
  const float4* p = reinterpret_cast<const float4*>(smem_row);
  float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    const float4 t = p[j], u = p[j + 1];
    s0 += a[4 * j] * t.x + a[4 * j + 1] * t.y + a[4 * j + 2] * t.z + a[4 * j + 3] * t.w;
    s1 += a[4 * j + 4] * u.x + a[4 * j + 5] * u.y + a[4 * j + 6] * u.z + a[4 * j + 7] * u.w;
  }
  return s0 + s1;
}
__device__ __forceinline__ void axpy32(float* acc, float c, const float* smem_row) {
  const float4* p = reinterpret_cast<const float4*>(smem_row);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = p[j];
    acc[4 * j] += c * t.x; acc[4 * j + 1] += c * t.y; acc[4 * j + 2] += c * t.z; acc[4 * j + 3] += c * t.w;
  }
}
__device__ __forceinline__ void store_split32(__nv_bfloat16* hi, __nv_bfloat16* lo, const float* v) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      split_bf16_pair(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1], hw[j], lw[j]);
    }
    reinterpret_cast<uint4*>(hi)[g] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    reinterpret_cast<uint4*>(lo)[g] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}
// K/V (or Q/dO) tiles are staged with a padded pitch so that the 4 key-split lanes of a query read 4
// different rows without bank conflicts.
constexpr int PITCH = HD + 4;
constexpr int QPB = 32;    // queries (or keys) per CTA
constexpr int SPLIT = 4;   // threads per query: each walks every 4th key; partials merged with shuffles

__device__ __forceinline__ void stage_rows(const float* src, int64_t ld, int64_t row0, int nrows, int col0,
                                           float* dst) {
  for (int i = threadIdx.x; i < TK * 8; i += blockDim.x) {
    const int r = i >> 3, c4 = i & 7;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrows) t = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * ld + col0 + c4 * 4));
    *reinterpret_cast<float4*>(dst + r * PITCH + c4 * 4) = t;
  }
}
__device__ __forceinline__ void shfl_add32(float* v) {
#pragma unroll
  for (int o = 1; o < SPLIT; o <<= 1)
#pragma unroll
    for (int c = 0; c < HD; ++c) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
}

// grid (ceil(L/QPB), nh, B), block QPB*SPLIT: thread = (query tid/4, key split tid%4)
__global__ void __launch_bounds__(QPB * SPLIT) mha_fwd_kernel(const MhaArgs a) {
  __shared__ __align__(16) float Ks[TK * PITCH];
  __shared__ __align__(16) float Vs[TK * PITCH];
  const int head = blockIdx.y, b = blockIdx.z;
  const int i = blockIdx.x * QPB + (threadIdx.x >> 2);
  const int ks = threadIdx.x & 3;
  const bool ok = i < a.L;
  const float scale = rsqrtf((float)HD);
  float q[HD], acc[HD];
#pragma unroll
  for (int j = 0; j < HD; ++j) { q[j] = 0.0f; acc[j] = 0.0f; }
  if (ok) {
    load32(a.q + ((int64_t)b * a.L + i) * a.ldq + head * HD, q);
#pragma unroll
    for (int j = 0; j < HD; ++j) q[j] *= scale;
  }
  float m = -INFINITY, l = 0.0f;
  for (int j0 = 0; j0 < a.L; j0 += TK) {
    const int nk = min(TK, a.L - j0);
    __syncthreads();
    stage_rows(a.k, a.ldq, (int64_t)b * a.L + j0, nk, head * HD, Ks);
    stage_rows(a.v, a.ldq, (int64_t)b * a.L + j0, nk, head * HD, Vs);
    __syncthreads();
    for (int j = ks; j < nk; j += 2 * SPLIT) {   // two keys per iteration: independent dot products
      const int j2 = j + SPLIT;
      const float s0 = dot32(q, Ks + j * PITCH);
      const float s1 = j2 < nk ? dot32(q, Ks + j2 * PITCH) : -INFINITY;
      const float mn = fmaxf(m, fmaxf(s0, s1));
      const float corr = expf(m - mn);
      const float p0 = expf(s0 - mn), p1 = expf(s1 - mn);
      l = l * corr + p0 + p1;
#pragma unroll
      for (int c = 0; c < HD; ++c) acc[c] *= corr;
      axpy32(acc, p0, Vs + j * PITCH);
      if (j2 < nk) axpy32(acc, p1, Vs + j2 * PITCH);
      m = mn;
    }
  }
  // merge the 4 key-split partials of each query (adjacent lanes)
#pragma unroll
  for (int o = 1; o < SPLIT; o <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, o);
    const float mn = fmaxf(m, m2);
    const float c1 = mn == -INFINITY ? 0.0f : expf(m - mn), c2 = mn == -INFINITY ? 0.0f : expf(m2 - mn);
    l = l * c1 + l2 * c2;
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = acc[c] * c1 + __shfl_xor_sync(0xffffffffu, acc[c], o) * c2;
    m = mn;
  }
  if (ok && ks == 0) {
    const float inv = 1.0f / l;
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] *= inv;
    const int64_t off = ((int64_t)b * a.L + i) * a.ld_o + head * HD;
    store_split32(a.o_hi + off, a.o_lo + off, acc);
    a.lse[((int64_t)b * a.nh + head) * a.L + i] = m + logf(l);
  }
}

// backward, per query: D_i, dq_i.   O is re-read from its split copy.
__global__ void __launch_bounds__(QPB * SPLIT) mha_bwd_q_kernel(const MhaArgs a) {
  __shared__ __align__(16) float Ks[TK * PITCH];
  __shared__ __align__(16) float Vs[TK * PITCH];
  const int head = blockIdx.y, b = blockIdx.z;
  const int i = blockIdx.x * QPB + (threadIdx.x >> 2);
  const int ks = threadIdx.x & 3;
  const bool ok = i < a.L;
  const float scale = rsqrtf((float)HD);
  float q[HD], dov[HD], dq[HD];
  float D = 0.0f, lse = 0.0f;
#pragma unroll
  for (int j = 0; j < HD; ++j) { q[j] = 0.0f; dov[j] = 0.0f; dq[j] = 0.0f; }
  if (ok) {
    load32(a.q + ((int64_t)b * a.L + i) * a.ldq + head * HD, q);
#pragma unroll
    for (int j = 0; j < HD; ++j) q[j] *= scale;
    load32(a.d_o + ((int64_t)b * a.L + i) * a.E + head * HD, dov);
    const int64_t off = ((int64_t)b * a.L + i) * a.ld_o + head * HD;
#pragma unroll
    for (int c = 0; c < HD; ++c) D += dov[c] * join_bf16(a.o_hi[off + c], a.o_lo[off + c]);
    lse = a.lse[((int64_t)b * a.nh + head) * a.L + i];
    if (ks == 0) a.dsum[((int64_t)b * a.nh + head) * a.L + i] = D;
  }
  for (int j0 = 0; j0 < a.L; j0 += TK) {
    const int nk = min(TK, a.L - j0);
    __syncthreads();
    stage_rows(a.k, a.ldq, (int64_t)b * a.L + j0, nk, head * HD, Ks);
    stage_rows(a.v, a.ldq, (int64_t)b * a.L + j0, nk, head * HD, Vs);
    __syncthreads();
    if (ok) {
      for (int j = ks; j < nk; j += SPLIT) {
        const float p = expf(dot32(q, Ks + j * PITCH) - lse);
        const float ds = p * (dot32(dov, Vs + j * PITCH) - D);
        axpy32(dq, ds, Ks + j * PITCH);
      }
    }
  }
  shfl_add32(dq);
  if (ok && ks == 0) {
#pragma unroll
    for (int c = 0; c < HD; ++c) dq[c] *= scale;
    const int64_t off = ((int64_t)b * a.L + i) * a.ld_g + head * HD;
    store_split32(a.dq_hi + off, a.dq_lo + off, dq);
  }
}

// backward, per key: dk_j, dv_j (queries staged in tiles; the 4 lanes of a key walk every 4th query)
__global__ void __launch_bounds__(QPB * SPLIT) mha_bwd_kv_kernel(const MhaArgs a) {
  __shared__ __align__(16) float Qs[TK * PITCH];
  __shared__ __align__(16) float Ds[TK * PITCH];   // dO tile
  __shared__ float lses[TK], dsums[TK];
  const int head = blockIdx.y, b = blockIdx.z;
  const int j = blockIdx.x * QPB + (threadIdx.x >> 2);
  const int qs = threadIdx.x & 3;
  const bool ok = j < a.L;
  const float scale = rsqrtf((float)HD);
  float kk[HD], vv[HD], dk[HD], dv[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) { kk[c] = 0.0f; vv[c] = 0.0f; dk[c] = 0.0f; dv[c] = 0.0f; }
  if (ok) {
    load32(a.k + ((int64_t)b * a.L + j) * a.ldq + head * HD, kk);
    load32(a.v + ((int64_t)b * a.L + j) * a.ldq + head * HD, vv);
#pragma unroll
    for (int c = 0; c < HD; ++c) kk[c] *= scale;   // s_ij = (scale q_i) . k_j = q_i . (scale k_j)
  }
  const int64_t bh = (int64_t)b * a.nh + head;
  for (int i0 = 0; i0 < a.L; i0 += TK) {
    const int nq = min(TK, a.L - i0);
    __syncthreads();
    stage_rows(a.q, a.ldq, (int64_t)b * a.L + i0, nq, head * HD, Qs);
    stage_rows(a.d_o, a.E, (int64_t)b * a.L + i0, nq, head * HD, Ds);
    if (threadIdx.x < nq) {
      lses[threadIdx.x] = a.lse[bh * a.L + i0 + threadIdx.x];
      dsums[threadIdx.x] = a.dsum[bh * a.L + i0 + threadIdx.x];
    }
    __syncthreads();
    if (ok) {
      for (int i = qs; i < nq; i += SPLIT) {
        const float p = expf(dot32(kk, Qs + i * PITCH) - lses[i]);
        axpy32(dv, p, Ds + i * PITCH);
        const float ds = p * (dot32(vv, Ds + i * PITCH) - dsums[i]);
        axpy32(dk, ds * scale, Qs + i * PITCH);
      }
    }
  }
  shfl_add32(dk);
  shfl_add32(dv);
  if (ok && qs == 0) {
    const int64_t off = ((int64_t)b * a.L + j) * a.ld_g + head * HD;
    store_split32(a.dk_hi + off, a.dk_lo + off, dk);
    store_split32(a.dv_hi + off, a.dv_lo + off, dv);
  }
}

}  // namespace

// tensor-core path (mha_tc.cu) whenever one head fits in shared memory; CDETR_MHA_LEGACY=1 forces the CUDA-core kernels
static bool use_tc(int L) { return cdetr_tuning().mha_legacy == 0 && mha_tc_fits(L) != 0; }

extern "C" int cdetr_mha_fwd(int B, int L, int E, int nh, const float* q, const float* k, const float* v,
                             int64_t ldq, cdetr_split_t o, float* lse, cdetr_stream_t s) {
  CDETR_CHECK_ARG(E == nh * HD && q && k && v && o.base && lse && ldq % 4 == 0, "mha_fwd: bad args");
  MhaArgs a = {};
  a.B = B; a.L = L; a.E = E; a.nh = nh; a.q = q; a.k = k; a.v = v; a.ldq = ldq;
  a.o_hi = reinterpret_cast<__nv_bfloat16*>(o.base); a.o_lo = a.o_hi + o.plane; a.ld_o = o.ld;
  a.lse = lse;
  if (use_tc(L)) return mha_fwd_tc_launch(a, reinterpret_cast<cudaStream_t>(s));
  mha_fwd_kernel<<<dim3(cdiv(L, QPB), nh, B), QPB * SPLIT, 0, reinterpret_cast<cudaStream_t>(s)>>>(a);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_mha_bwd(int B, int L, int E, int nh, const float* q, const float* k, const float* v,
                             int64_t ldq, cdetr_split_t o, const float* lse, const float* d_o, float* dsum,
                             cdetr_split_t dq, cdetr_split_t dk, cdetr_split_t dv, cdetr_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  CDETR_CHECK_ARG(E == nh * HD && q && k && v && o.base && lse && d_o && dsum && dq.base && dk.base && dv.base,
                  "mha_bwd: bad args");
  CDETR_CHECK_ARG(dq.ld == dk.ld && dq.ld == dv.ld, "mha_bwd: gradient tensors must share ld");
  MhaArgs a = {};
  a.B = B; a.L = L; a.E = E; a.nh = nh; a.q = q; a.k = k; a.v = v; a.ldq = ldq;
  a.o_hi = reinterpret_cast<__nv_bfloat16*>(o.base); a.o_lo = a.o_hi + o.plane; a.ld_o = o.ld;
  a.lse = const_cast<float*>(lse); a.d_o = d_o; a.dsum = dsum;
  auto hi = [](cdetr_split_t t) { return reinterpret_cast<__nv_bfloat16*>(t.base); };
  a.dq_hi = hi(dq); a.dq_lo = hi(dq) + dq.plane;
  a.dk_hi = hi(dk); a.dk_lo = hi(dk) + dk.plane;
  a.dv_hi = hi(dv); a.dv_lo = hi(dv) + dv.plane;
  a.ld_g = dq.ld;
  if (use_tc(L)) return mha_bwd_tc_launch(a, s);
  mha_bwd_q_kernel<<<dim3(cdiv(L, QPB), nh, B), QPB * SPLIT, 0, s>>>(a);
  CDETR_CHECK_LAUNCH();
  mha_bwd_kv_kernel<<<dim3(cdiv(L, QPB), nh, B), QPB * SPLIT, 0, s>>>(a);
  CDETR_CHECK_LAUNCH();
  return 0;
}
