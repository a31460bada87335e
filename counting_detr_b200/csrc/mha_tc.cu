// Decoder self-attention core on the tensor cores: warp-level mma.sync.m16n8k16 (bf16 x bf16 -> fp32) with the
// library's split-bf16 operands (three products hi*lo + lo*hi + hi*hi per MMA position, fp32 accumulation), so the
// results keep the fp32-level accuracy of the CUDA-core kernels in mha.cu while issuing ~15x fewer instructions.
// tcgen05 is not used here on purpose: per (sample, head) the problem is [L x 32] x [32 x L] with L = 300 - the
// operands of one head fit in shared memory once and 16-row strips per warp need no TMEM round trip.
//
// One CTA per (sample, head); K / V (forward, query-side backward) or Q / dO (key-side backward) of the head are
// converted once to split-bf16 in shared memory in the two layouts the MMA B operand needs:
//   X[row][KP] read directly        : B[k = channel][n = row]  (S = Q K^T, dP = dO V^T, S^T = K Q^T, dP^T = V dO^T)
//   X[row][KP] through ldmatrix.trans: B[k = row][n = channel]  (O = P V, dQ = dS K, dV = P^T dO, dK = dS^T Q)
// Each warp owns strips of 16 queries (or keys) and walks the other dimension in chunks of 64 with the accumulator
// fragments of S / P re-used directly as the A fragments of the second product (no shared-memory round trip).
// Reference: A2/models/transformer.py:366-372 (nn.MultiheadAttention in the decoder layer) and its autograd.
#include "common.cuh"
#include "../../include/cdetr.h"
#include "mha_args.cuh"

namespace {
constexpr int HD = 32;
constexpr int KP = 40;        // bf16 pitch of row-major staged rows: fragment loads hit 32 distinct banks
constexpr int NWARPS = 10;    // 19 strips of 16 rows at L = 300 -> two rounds
constexpr int NTHREADS = NWARPS * 32;

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// error-compensated product: a = ah + al, b = bh + bl; the al*bl term (2^-16 relative) is dropped
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                     uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_bf16(c, ah, bl0, bl1);
  mma_bf16(c, al, bh0, bh1);
  mma_bf16(c, ah, bh0, bh1);
}

// rows of a [*, ld] fp32 matrix (columns col0..col0+31) -> split-bf16 row-major smem [Lp][KP]; rows >= L are zero
__device__ __forceinline__ void stage_rowmajor(const float* __restrict__ src, int64_t ld, int64_t row0, int L, int Lp,
                                               int col0, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  for (int i = threadIdx.x; i < Lp * 8; i += NTHREADS) {
    const int r = i >> 3, c4 = i & 7;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < L) t = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * ld + col0 + c4 * 4));
    uint32_t h0, l0, h1, l1;
    split_bf16_pair(t.x, t.y, h0, l0);
    split_bf16_pair(t.z, t.w, h1, l1);
    *reinterpret_cast<uint2*>(hi + r * KP + c4 * 4) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(lo + r * KP + c4 * 4) = make_uint2(l0, l1);
  }
}
// A fragments (both k-steps of the 32 channels) of two rows of a global fp32 matrix, scaled, as split-bf16
__device__ __forceinline__ void load_a_frags(const float* __restrict__ p0, const float* __restrict__ p1, bool ok0,
                                             bool ok1, float scale, int t, uint32_t (&ah)[2][4], uint32_t (&al)[2][4]) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int c = 16 * ks + 2 * t;
    const float2 z = make_float2(0.f, 0.f);
    const float2 x00 = ok0 ? __ldg(reinterpret_cast<const float2*>(p0 + c)) : z;
    const float2 x01 = ok0 ? __ldg(reinterpret_cast<const float2*>(p0 + c + 8)) : z;
    const float2 x10 = ok1 ? __ldg(reinterpret_cast<const float2*>(p1 + c)) : z;
    const float2 x11 = ok1 ? __ldg(reinterpret_cast<const float2*>(p1 + c + 8)) : z;
    split_bf16_pair(x00.x * scale, x00.y * scale, ah[ks][0], al[ks][0]);
    split_bf16_pair(x10.x * scale, x10.y * scale, ah[ks][1], al[ks][1]);
    split_bf16_pair(x01.x * scale, x01.y * scale, ah[ks][2], al[ks][2]);
    split_bf16_pair(x11.x * scale, x11.y * scale, ah[ks][3], al[ks][3]);
  }
}

// c[8][4] (16 rows x 64 columns starting at column n0 of the row-major staged matrix X) += A[16 x 32] X[n0.., :]^T
__device__ __forceinline__ void gemm_rows_x_rowmajor(float (&c)[8][4], const uint32_t (&ah)[2][4],
                                                     const uint32_t (&al)[2][4], const __nv_bfloat16* Xh,
                                                     const __nv_bfloat16* Xl, int n0, int ntn, int g, int t) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (nt < ntn) {
      const int off = (n0 + nt * 8 + g) * KP + 2 * t;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(Xh + off + 16 * ks);
        const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(Xh + off + 16 * ks + 8);
        const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(Xl + off + 16 * ks);
        const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(Xl + off + 16 * ks + 8);
        mma3(c[nt], ah[ks], al[ks], bh0, bh1, bl0, bl1);
      }
    }
  }
}
// B fragments of a [16 k x 8 n] block of a ROW-MAJOR staged matrix X[k][n] (k = row, pitch KP): ldmatrix with .trans
// hands thread (g, t) the pairs X[k0 + 2t .. 2t+1][n0 + g] (b0) and X[k0 + 8 + 2t ..][n0 + g] (b1), i.e. the "col"
// operand layout of mma.m16n8k16 without a transposed copy in shared memory.
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& b0, uint32_t& b1, const __nv_bfloat16* X, int k0, int n0,
                                                  int lane) {
  const __nv_bfloat16* p = X + (k0 + (lane & 15)) * KP + n0;   // lanes 0-7: rows k0..k0+7, lanes 8-15: rows k0+8..k0+15
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];"
               : "=r"(b0), "=r"(b1)
               : "r"(smem_u32(p)));
}
// o[4][4] (16 rows x 32 channels) += P[16 x 64] X[k0..k0+63, :], P given as accumulator fragments p[8][4], X row-major
__device__ __forceinline__ void gemm_p_x_rows(float (&o)[4][4], const float (&p)[8][4], const __nv_bfloat16* Xh,
                                              const __nv_bfloat16* Xl, int k0, int ntn, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    if (2 * kk < ntn) {
      uint32_t ah[4], al[4];
      split_bf16_pair(p[2 * kk][0], p[2 * kk][1], ah[0], al[0]);
      split_bf16_pair(p[2 * kk][2], p[2 * kk][3], ah[1], al[1]);
      split_bf16_pair(p[2 * kk + 1][0], p[2 * kk + 1][1], ah[2], al[2]);
      split_bf16_pair(p[2 * kk + 1][2], p[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        uint32_t bh0, bh1, bl0, bl1;
        ldmatrix_x2_trans(bh0, bh1, Xh, k0 + 16 * kk, dt * 8, lane);
        ldmatrix_x2_trans(bl0, bl1, Xl, k0 + 16 * kk, dt * 8, lane);
        mma3(o[dt], ah, al, bh0, bh1, bl0, bl1);
      }
    }
  }
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
// 16 x 32 accumulator -> split-bf16 global rows (row g: o[.][0..1], row g+8: o[.][2..3])
__device__ __forceinline__ void store_rows_split(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t off0, int64_t off1,
                                                 bool ok0, bool ok1, const float (&o)[4][4], float s0, float s1, int t) {
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    uint32_t h, l;
    if (ok0) {
      split_bf16_pair(o[dt][0] * s0, o[dt][1] * s0, h, l);
      *reinterpret_cast<uint32_t*>(hi + off0 + dt * 8 + 2 * t) = h;
      *reinterpret_cast<uint32_t*>(lo + off0 + dt * 8 + 2 * t) = l;
    }
    if (ok1) {
      split_bf16_pair(o[dt][2] * s1, o[dt][3] * s1, h, l);
      *reinterpret_cast<uint32_t*>(hi + off1 + dt * 8 + 2 * t) = h;
      *reinterpret_cast<uint32_t*>(lo + off1 + dt * 8 + 2 * t) = l;
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward
// grid (nh, B, nsplit); strips of 16 queries round-robin over (warp, blockIdx.z)
__global__ void __launch_bounds__(NTHREADS, 1) mha_fwd_tc_kernel(const MhaArgs a, const int Lp) {
  pdl_trigger();   // light successors (launch_light) may pre-launch; they wait for this grid to finish
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* Kh = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Kl = Kh + Lp * KP;
  __nv_bfloat16* Vh = Kl + Lp * KP;
  __nv_bfloat16* Vl = Vh + Lp * KP;
  const int head = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t row0 = (int64_t)b * a.L;
  stage_rowmajor(a.k, a.ldq, row0, a.L, Lp, head * HD, Kh, Kl);
  stage_rowmajor(a.v, a.ldq, row0, a.L, Lp, head * HD, Vh, Vl);
  __syncthreads();
  const float scale = rsqrtf((float)HD);
  for (int strip = warp + NWARPS * blockIdx.z; strip * 16 < a.L; strip += NWARPS * gridDim.z) {
    const int i0 = strip * 16 + g, i1 = i0 + 8;
    const bool ok0 = i0 < a.L, ok1 = i1 < a.L;
    uint32_t qh[2][4], ql[2][4];
    load_a_frags(a.q + (row0 + i0) * a.ldq + head * HD, a.q + (row0 + i1) * a.ldq + head * HD, ok0, ok1, scale, t, qh, ql);
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[4][4];
#pragma unroll
    for (int dt = 0; dt < 4; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;
    for (int j0 = 0; j0 < Lp; j0 += 64) {
      const int ntn = min(8, (Lp - j0) >> 3);
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
      gemm_rows_x_rowmajor(s, qh, ql, Kh, Kl, j0, ntn, g, t);
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = j0 + nt * 8 + 2 * t;
        if (nt >= ntn || col >= a.L) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (nt >= ntn || col + 1 >= a.L) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));
      const float c0 = expf(m0 - mn0), c1 = expf(m1 - mn1);
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = expf(s[nt][0] - mn0); s[nt][1] = expf(s[nt][1] - mn0);
        s[nt][2] = expf(s[nt][2] - mn1); s[nt][3] = expf(s[nt][3] - mn1);
        sum0 += s[nt][0] + s[nt][1];
        sum1 += s[nt][2] + s[nt][3];
      }
      l0 = l0 * c0 + sum0; l1 = l1 * c1 + sum1;
      m0 = mn0; m1 = mn1;
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) { o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1; }
      gemm_p_x_rows(o, s, Vh, Vl, j0, ntn, lane);
    }
    l0 = quad_sum(l0); l1 = quad_sum(l1);
    store_rows_split(a.o_hi, a.o_lo, (row0 + i0) * a.ld_o + head * HD, (row0 + i1) * a.ld_o + head * HD, ok0, ok1, o,
                     1.0f / l0, 1.0f / l1, t);
    if (t == 0) {
      const int64_t bh = ((int64_t)b * a.nh + head) * a.L;
      if (ok0) a.lse[bh + i0] = m0 + logf(l0);
      if (ok1) a.lse[bh + i1] = m1 + logf(l1);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward, queries
// D_i = dO_i . O_i (written to dsum for the key-side kernel), dq_i = scale * sum_j dS_ij k_j
__global__ void __launch_bounds__(NTHREADS, 1) mha_bwd_q_tc_kernel(const MhaArgs a, const int Lp) {
  pdl_trigger();   // light successors (launch_light) may pre-launch; they wait for this grid to finish
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* Kh = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Kl = Kh + Lp * KP;
  __nv_bfloat16* Vh = Kl + Lp * KP;
  __nv_bfloat16* Vl = Vh + Lp * KP;
  const int head = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t row0 = (int64_t)b * a.L;
  stage_rowmajor(a.k, a.ldq, row0, a.L, Lp, head * HD, Kh, Kl);
  stage_rowmajor(a.v, a.ldq, row0, a.L, Lp, head * HD, Vh, Vl);
  __syncthreads();
  const float scale = rsqrtf((float)HD);
  const int64_t bh = ((int64_t)b * a.nh + head) * a.L;
  for (int strip = warp + NWARPS * blockIdx.z; strip * 16 < a.L; strip += NWARPS * gridDim.z) {
    const int i0 = strip * 16 + g, i1 = i0 + 8;
    const bool ok0 = i0 < a.L, ok1 = i1 < a.L;
    uint32_t qh[2][4], ql[2][4], dh[2][4], dl[2][4];
    load_a_frags(a.q + (row0 + i0) * a.ldq + head * HD, a.q + (row0 + i1) * a.ldq + head * HD, ok0, ok1, scale, t, qh, ql);
    const float* d0p = a.d_o + (row0 + i0) * a.E + head * HD;
    const float* d1p = a.d_o + (row0 + i1) * a.E + head * HD;
    load_a_frags(d0p, d1p, ok0, ok1, 1.0f, t, dh, dl);
    // D = dO . O over this thread's 8 channels of each row, then across the quad
    float D0 = 0.f, D1 = 0.f;
    {
      const int64_t o0 = (row0 + i0) * a.ld_o + head * HD, o1 = (row0 + i1) * a.ld_o + head * HD;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int hseg = 0; hseg < 2; ++hseg) {
          const int c = 16 * ks + 8 * hseg + 2 * t;
          if (ok0) {
            const float2 dv = __ldg(reinterpret_cast<const float2*>(d0p + c));
            D0 += dv.x * join_bf16(a.o_hi[o0 + c], a.o_lo[o0 + c]) + dv.y * join_bf16(a.o_hi[o0 + c + 1], a.o_lo[o0 + c + 1]);
          }
          if (ok1) {
            const float2 dv = __ldg(reinterpret_cast<const float2*>(d1p + c));
            D1 += dv.x * join_bf16(a.o_hi[o1 + c], a.o_lo[o1 + c]) + dv.y * join_bf16(a.o_hi[o1 + c + 1], a.o_lo[o1 + c + 1]);
          }
        }
      D0 = quad_sum(D0); D1 = quad_sum(D1);
    }
    const float lse0 = ok0 ? a.lse[bh + i0] : 0.f, lse1 = ok1 ? a.lse[bh + i1] : 0.f;
    if (t == 0) {
      if (ok0) a.dsum[bh + i0] = D0;
      if (ok1) a.dsum[bh + i1] = D1;
    }
    float dq[4][4];
#pragma unroll
    for (int dt = 0; dt < 4; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[dt][e] = 0.f;
    for (int j0 = 0; j0 < Lp; j0 += 64) {
      const int ntn = min(8, (Lp - j0) >> 3);
      float s[8][4], dp[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { s[nt][e] = 0.f; dp[nt][e] = 0.f; }
      gemm_rows_x_rowmajor(s, qh, ql, Kh, Kl, j0, ntn, g, t);
      gemm_rows_x_rowmajor(dp, dh, dl, Vh, Vl, j0, ntn, g, t);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = j0 + nt * 8 + 2 * t;
        const bool v0 = nt < ntn && col < a.L, v1 = nt < ntn && col + 1 < a.L;
        const float p0 = v0 ? expf(s[nt][0] - lse0) : 0.f, p1 = v1 ? expf(s[nt][1] - lse0) : 0.f;
        const float p2 = v0 ? expf(s[nt][2] - lse1) : 0.f, p3 = v1 ? expf(s[nt][3] - lse1) : 0.f;
        s[nt][0] = p0 * (dp[nt][0] - D0); s[nt][1] = p1 * (dp[nt][1] - D0);
        s[nt][2] = p2 * (dp[nt][2] - D1); s[nt][3] = p3 * (dp[nt][3] - D1);
      }
      gemm_p_x_rows(dq, s, Kh, Kl, j0, ntn, lane);
    }
    store_rows_split(a.dq_hi, a.dq_lo, (row0 + i0) * a.ld_g + head * HD, (row0 + i1) * a.ld_g + head * HD, ok0, ok1, dq,
                     scale, scale, t);
  }
}

// ------------------------------------------------------------------------------------------------ backward, keys
// dv_j = sum_i P_ij dO_i ;  dk_j = scale * sum_i dS_ij q_i     (strips of 16 keys, chunks of 64 queries)
__global__ void __launch_bounds__(NTHREADS, 1) mha_bwd_kv_tc_kernel(const MhaArgs a, const int Lp) {
  pdl_trigger();   // light successors (launch_light) may pre-launch; they wait for this grid to finish
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* Qh = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Ql = Qh + Lp * KP;
  __nv_bfloat16* Dh = Ql + Lp * KP;
  __nv_bfloat16* Dl = Dh + Lp * KP;
  float* lses = reinterpret_cast<float*>(Dl + Lp * KP);
  float* dsums = lses + Lp;
  const int head = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t row0 = (int64_t)b * a.L;
  const int64_t bh = ((int64_t)b * a.nh + head) * a.L;
  stage_rowmajor(a.q, a.ldq, row0, a.L, Lp, head * HD, Qh, Ql);
  stage_rowmajor(a.d_o, a.E, row0, a.L, Lp, head * HD, Dh, Dl);
  for (int i = threadIdx.x; i < Lp; i += NTHREADS) {
    lses[i] = i < a.L ? a.lse[bh + i] : INFINITY;   // padded queries: P = exp(s - inf) = 0
    dsums[i] = i < a.L ? a.dsum[bh + i] : 0.f;
  }
  __syncthreads();
  const float scale = rsqrtf((float)HD);
  for (int strip = warp + NWARPS * blockIdx.z; strip * 16 < a.L; strip += NWARPS * gridDim.z) {
    const int j0r = strip * 16 + g, j1r = j0r + 8;
    const bool ok0 = j0r < a.L, ok1 = j1r < a.L;
    uint32_t kh[2][4], kl[2][4], vh[2][4], vl[2][4];
    load_a_frags(a.k + (row0 + j0r) * a.ldq + head * HD, a.k + (row0 + j1r) * a.ldq + head * HD, ok0, ok1, scale, t, kh, kl);
    load_a_frags(a.v + (row0 + j0r) * a.ldq + head * HD, a.v + (row0 + j1r) * a.ldq + head * HD, ok0, ok1, 1.0f, t, vh, vl);
    float dk[4][4], dv[4][4];
#pragma unroll
    for (int dt = 0; dt < 4; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) { dk[dt][e] = 0.f; dv[dt][e] = 0.f; }
    for (int i0 = 0; i0 < Lp; i0 += 64) {
      const int ntn = min(8, (Lp - i0) >> 3);
      float s[8][4], dp[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { s[nt][e] = 0.f; dp[nt][e] = 0.f; }
      gemm_rows_x_rowmajor(s, kh, kl, Qh, Ql, i0, ntn, g, t);     // S^T[key, query]
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < ntn) {
          const float2 ls = *reinterpret_cast<const float2*>(lses + i0 + nt * 8 + 2 * t);
          s[nt][0] = expf(s[nt][0] - ls.x); s[nt][1] = expf(s[nt][1] - ls.y);
          s[nt][2] = expf(s[nt][2] - ls.x); s[nt][3] = expf(s[nt][3] - ls.y);
        }
      }
      gemm_p_x_rows(dv, s, Dh, Dl, i0, ntn, lane);                 // dV += P^T dO
      gemm_rows_x_rowmajor(dp, vh, vl, Dh, Dl, i0, ntn, g, t);     // dP^T[key, query] = V dO^T
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < ntn) {
          const float2 ds = *reinterpret_cast<const float2*>(dsums + i0 + nt * 8 + 2 * t);
          s[nt][0] *= dp[nt][0] - ds.x; s[nt][1] *= dp[nt][1] - ds.y;
          s[nt][2] *= dp[nt][2] - ds.x; s[nt][3] *= dp[nt][3] - ds.y;
        }
      }
      gemm_p_x_rows(dk, s, Qh, Ql, i0, ntn, lane);                 // dK += dS^T Q
    }
    const int64_t off0 = (row0 + j0r) * a.ld_g + head * HD, off1 = (row0 + j1r) * a.ld_g + head * HD;
    store_rows_split(a.dk_hi, a.dk_lo, off0, off1, ok0, ok1, dk, scale, scale, t);
    store_rows_split(a.dv_hi, a.dv_lo, off0, off1, ok0, ok1, dv, 1.0f, 1.0f, t);
  }
}

size_t smem_fwd(int Lp) { return (size_t)(4 * Lp * KP) * 2; }
size_t smem_bwd_q(int Lp) { return (size_t)(4 * Lp * KP) * 2; }
size_t smem_bwd_kv(int Lp) { return (size_t)(4 * Lp * KP) * 2 + (size_t)2 * Lp * 4; }
constexpr size_t SMEM_MAX = 227 * 1024;

}  // namespace

// Host launchers used by cdetr_mha_fwd / cdetr_mha_bwd (mha.cu).  Return 1 when the head does not fit in shared
// memory (L > ~700): the caller then falls back to the CUDA-core kernels.
// One CTA per (sample, head) leaves SMs idle when B * heads < #SMs (C4: 64 CTAs on 148 SMs; inference at B = 1: 8):
// the strips of 16 rows are dealt round-robin over gridDim.z CTAs, each staging the head's K / V (or Q / dO) itself.
static int mha_query_split(const MhaArgs& a) {
  int num_sms = 148;
  if (cdetr_num_sms(&num_sms) != cudaSuccess) num_sms = 148;
  const int strips = (a.L + 15) / 16;
  int nsplit = num_sms / (a.nh * a.B > 0 ? a.nh * a.B : 1);
  const int max_split = (strips + NWARPS - 1) / NWARPS;      // below one round of strips per CTA there is nothing to gain
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit > 4) nsplit = 4;
  return nsplit < 1 ? 1 : nsplit;
}
int mha_tc_fits(int L) {
  const int Lp = (L + 15) / 16 * 16;
  return smem_bwd_kv(Lp) <= SMEM_MAX && smem_bwd_q(Lp) <= SMEM_MAX ? 1 : 0;
}
int mha_fwd_tc_launch(const MhaArgs& a, cudaStream_t s) {
  const int Lp = (a.L + 15) / 16 * 16;
  const size_t smem = smem_fwd(Lp);
  static DevAttrCache cfg = {};
  CDETR_CHECK_CUDA(cdetr_ensure_smem(mha_fwd_tc_kernel, (int)SMEM_MAX, &cfg));
  const int nsplit = mha_query_split(a);
  mha_fwd_tc_kernel<<<dim3(a.nh, a.B, nsplit), NTHREADS, smem, s>>>(a, Lp);   // ~160 registers x 320 threads: one CTA per SM
  CDETR_CHECK_LAUNCH();
  return 0;
}
int mha_bwd_tc_launch(const MhaArgs& a, cudaStream_t s) {
  const int Lp = (a.L + 15) / 16 * 16;
  static DevAttrCache cfg_q = {}, cfg_kv = {};
  CDETR_CHECK_CUDA(cdetr_ensure_smem(mha_bwd_q_tc_kernel, (int)SMEM_MAX, &cfg_q));
  CDETR_CHECK_CUDA(cdetr_ensure_smem(mha_bwd_kv_tc_kernel, (int)SMEM_MAX, &cfg_kv));
  const int nsplit = mha_query_split(a);
  mha_bwd_q_tc_kernel<<<dim3(a.nh, a.B, nsplit), NTHREADS, smem_bwd_q(Lp), s>>>(a, Lp);
  CDETR_CHECK_LAUNCH();
  mha_bwd_kv_tc_kernel<<<dim3(a.nh, a.B, nsplit), NTHREADS, smem_bwd_kv(Lp), s>>>(a, Lp);
  CDETR_CHECK_LAUNCH();
  return 0;
}
