// Row/column-decoupled attention core (AnchorDETR RCDA), forward and hand-written backward.
//   A_r = softmax_w(s * q_r K_r^T)  [L,W]     A_c = softmax_h(s * q_c K_c^T)  [L,H]      (s = d^-1/2)
//   O[q,:] = sum_h sum_w A_c[q,h] A_r[q,w] V[h,w,:]
// per (sample, head), head dim d = 32.  The reference materialises a [B*heads, L, W, d] intermediate
// (537 MB per encoder layer at B=16, 512x512) and keeps it for autograd; here the contraction is fused
// with the two softmaxes, the intermediate never exists, and backward recomputes from A_r/A_c.
// Reference: A2/models/row_column_decoupled_attention.py:210-291 (core), backward derived in
// SURVEY.md §8a-4.  Attention maps are stored transposed ([B,heads,W,L] / [B,heads,H,L]) so that
// per-query threads read/write them coalesced.
//
// Round-1 kernels run the contraction on the fp32 FMA pipe (V tile staged in shared memory and
// broadcast to one-query-per-thread accumulators); the tcgen05 formulation is the next step.
#include "common.cuh"
#include "../../include/cdetr.h"

namespace {

constexpr int HD = 32;  // head dim

struct RcdaArgs {
  int B, L, H, W, E, nh;
  int Hc;  // V rows per shared-memory chunk
  const float* qr;  // [B,L,E] projected (bias added, unscaled)
  const float* qc;
  const float* kr;  // [B,W,E]
  const float* kc;  // [B,H,E]
  const float* v;   // [B,H,W,E]
  const uint8_t* mask_row;  // [B,W] or null (1 = padded)
  const uint8_t* mask_col;  // [B,H] or null
  float* ar;  // [B,nh,W,L]
  float* ac;  // [B,nh,H,L]
  __nv_bfloat16* o_hi;  // [B,L,E] split
  __nv_bfloat16* o_lo;
  int64_t ld_o;
  // backward
  const float* d_o;  // [B,L,E]
  float* dsr;        // [B,nh,W,L]
  float* dsc;        // [B,nh,H,L]
  __nv_bfloat16 *dqr_hi, *dqr_lo, *dqc_hi, *dqc_lo;  // [B,L,E] split
  __nv_bfloat16 *dkr_hi, *dkr_lo, *dkc_hi, *dkc_lo;  // [B,W,E], [B,H,E] split
  __nv_bfloat16 *dv_hi, *dv_lo;                      // [B,H,W,E] split
  int64_t ld_g;  // ld of all gradient split tensors (E)
};

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

__device__ __forceinline__ void load32(const float* p, float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + j);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}

// one softmax over `n` keys for this thread's query; logits written to s_col[k*T + tid] then normalised
__device__ __forceinline__ void thread_softmax(const float* q, const float* Ks, int n,
                                               const uint8_t* mask, float* s_col, int T, int tid) {
  float mx = -INFINITY;
  for (int k = 0; k < n; ++k) {
    float s = 0.0f;
    const float4* kp = reinterpret_cast<const float4*>(Ks + k * HD);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = kp[j];
      s += q[4 * j] * t.x + q[4 * j + 1] * t.y + q[4 * j + 2] * t.z + q[4 * j + 3] * t.w;
    }
    if (mask && mask[k]) s = -INFINITY;
    s_col[k * T + tid] = s;
    mx = fmaxf(mx, s);
  }
  float sum = 0.0f;
  for (int k = 0; k < n; ++k) {
    const float e = expf(s_col[k * T + tid] - mx);
    s_col[k * T + tid] = e;
    sum += e;
  }
  const float inv = 1.0f / sum;
  for (int k = 0; k < n; ++k) s_col[k * T + tid] *= inv;
}

__device__ __forceinline__ void load_v_chunk(const RcdaArgs& a, int b, int head, int h0, int hc, float* Vs) {
  const int n4 = hc * a.W * 8;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const int pos = i >> 3, c4 = i & 7;
    const int h = h0 + pos / a.W, w = pos % a.W;
    const float4 t = __ldg(reinterpret_cast<const float4*>(
        a.v + (((int64_t)b * a.H + h) * a.W + w) * a.E + head * HD + c4 * 4));
    reinterpret_cast<float4*>(Vs)[i] = t;
  }
}

// ------------------------------------------------------------------ forward: grid (ceil(L/T), nh, B)
__global__ void rcda_fwd_kernel(const RcdaArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int T = blockDim.x, tid = threadIdx.x;
  const int head = blockIdx.y, b = blockIdx.z;
  float* Ksr = sm;                       // [W][32]
  float* Ksc = Ksr + a.W * HD;           // [H][32]
  float* ars = Ksc + a.H * HD;           // [W][T]
  float* acs = ars + a.W * T;            // [H][T]
  float* Vs = acs + a.H * T;             // [Hc*W][32]
  for (int i = tid; i < a.W * HD; i += T)
    Ksr[i] = a.kr[((int64_t)b * a.W + i / HD) * a.E + head * HD + (i % HD)];
  for (int i = tid; i < a.H * HD; i += T)
    Ksc[i] = a.kc[((int64_t)b * a.H + i / HD) * a.E + head * HD + (i % HD)];
  __syncthreads();
  const int q = blockIdx.x * T + tid;
  const bool ok = q < a.L;
  const float scale = rsqrtf((float)HD);
  {
    float qv[HD];
    if (ok) {
      load32(a.qr + ((int64_t)b * a.L + q) * a.E + head * HD, qv);
#pragma unroll
      for (int j = 0; j < HD; ++j) qv[j] *= scale;
      thread_softmax(qv, Ksr, a.W, a.mask_row ? a.mask_row + (int64_t)b * a.W : nullptr, ars, T, tid);
      load32(a.qc + ((int64_t)b * a.L + q) * a.E + head * HD, qv);
#pragma unroll
      for (int j = 0; j < HD; ++j) qv[j] *= scale;
      thread_softmax(qv, Ksc, a.H, a.mask_col ? a.mask_col + (int64_t)b * a.H : nullptr, acs, T, tid);
      const int64_t bh = (int64_t)b * a.nh + head;
      for (int w = 0; w < a.W; ++w) a.ar[(bh * a.W + w) * a.L + q] = ars[w * T + tid];
      for (int h = 0; h < a.H; ++h) a.ac[(bh * a.H + h) * a.L + q] = acs[h * T + tid];
    } else {
      for (int w = 0; w < a.W; ++w) ars[w * T + tid] = 0.0f;
      for (int h = 0; h < a.H; ++h) acs[h * T + tid] = 0.0f;
    }
  }
  float acc[HD];
#pragma unroll
  for (int j = 0; j < HD; ++j) acc[j] = 0.0f;
  for (int h0 = 0; h0 < a.H; h0 += a.Hc) {
    const int hc = min(a.Hc, a.H - h0);
    __syncthreads();
    load_v_chunk(a, b, head, h0, hc, Vs);
    __syncthreads();
    for (int hh = 0; hh < hc; ++hh) {
      const float cc = acs[(h0 + hh) * T + tid];
      for (int w = 0; w < a.W; ++w) {
        const float coef = cc * ars[w * T + tid];
        const float4* vp = reinterpret_cast<const float4*>(Vs + (hh * a.W + w) * HD);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = vp[j];
          acc[4 * j] += coef * t.x; acc[4 * j + 1] += coef * t.y;
          acc[4 * j + 2] += coef * t.z; acc[4 * j + 3] += coef * t.w;
        }
      }
    }
  }
  if (ok) {
    const int64_t off = ((int64_t)b * a.L + q) * a.ld_o + head * HD;
    store_split32(a.o_hi + off, a.o_lo + off, acc);
  }
}

// ------------------------------------------------------------------ backward 1 (per query): dA -> dS -> dq
__global__ void rcda_bwd_q_kernel(const RcdaArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int T = blockDim.x, tid = threadIdx.x;
  const int head = blockIdx.y, b = blockIdx.z;
  float* Ksr = sm;
  float* Ksc = Ksr + a.W * HD;
  float* ars = Ksc + a.H * HD;   // [W][T]
  float* acs = ars + a.W * T;    // [H][T]
  float* dar = acs + a.H * T;    // [W][T]
  float* dac = dar + a.W * T;    // [H][T]
  float* Vs = dac + a.H * T;
  for (int i = tid; i < a.W * HD; i += T)
    Ksr[i] = a.kr[((int64_t)b * a.W + i / HD) * a.E + head * HD + (i % HD)];
  for (int i = tid; i < a.H * HD; i += T)
    Ksc[i] = a.kc[((int64_t)b * a.H + i / HD) * a.E + head * HD + (i % HD)];
  const int q = blockIdx.x * T + tid;
  const bool ok = q < a.L;
  const int64_t bh = (int64_t)b * a.nh + head;
  float dov[HD];
  if (ok) {
    load32(a.d_o + ((int64_t)b * a.L + q) * a.E + head * HD, dov);
    for (int w = 0; w < a.W; ++w) ars[w * T + tid] = a.ar[(bh * a.W + w) * a.L + q];
    for (int h = 0; h < a.H; ++h) acs[h * T + tid] = a.ac[(bh * a.H + h) * a.L + q];
  } else {
#pragma unroll
    for (int j = 0; j < HD; ++j) dov[j] = 0.0f;
    for (int w = 0; w < a.W; ++w) ars[w * T + tid] = 0.0f;
    for (int h = 0; h < a.H; ++h) acs[h * T + tid] = 0.0f;
  }
  for (int w = 0; w < a.W; ++w) dar[w * T + tid] = 0.0f;
  for (int h0 = 0; h0 < a.H; h0 += a.Hc) {
    const int hc = min(a.Hc, a.H - h0);
    __syncthreads();
    load_v_chunk(a, b, head, h0, hc, Vs);
    __syncthreads();
    for (int hh = 0; hh < hc; ++hh) {
      const float cc = acs[(h0 + hh) * T + tid];
      float dcc = 0.0f;
      for (int w = 0; w < a.W; ++w) {
        const float4* vp = reinterpret_cast<const float4*>(Vs + (hh * a.W + w) * HD);
        float g0 = 0.0f, g1 = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float4 t = vp[j];
          const float4 u = vp[j + 1];
          g0 += dov[4 * j] * t.x + dov[4 * j + 1] * t.y + dov[4 * j + 2] * t.z + dov[4 * j + 3] * t.w;
          g1 += dov[4 * j + 4] * u.x + dov[4 * j + 5] * u.y + dov[4 * j + 6] * u.z + dov[4 * j + 7] * u.w;
        }
        const float g = g0 + g1;
        dcc += ars[w * T + tid] * g;
        dar[w * T + tid] += cc * g;
      }
      dac[(h0 + hh) * T + tid] = dcc;
    }
  }
  if (!ok) return;
  const float scale = rsqrtf((float)HD);
  // softmax backward + dq, row then column
  {
    float dot = 0.0f;
    for (int w = 0; w < a.W; ++w) dot += ars[w * T + tid] * dar[w * T + tid];
    float dq[HD];
#pragma unroll
    for (int j = 0; j < HD; ++j) dq[j] = 0.0f;
    for (int w = 0; w < a.W; ++w) {
      const float ds = ars[w * T + tid] * (dar[w * T + tid] - dot);
      a.dsr[(bh * a.W + w) * a.L + q] = ds;
      const float4* kp = reinterpret_cast<const float4*>(Ksr + w * HD);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = kp[j];
        dq[4 * j] += ds * t.x; dq[4 * j + 1] += ds * t.y; dq[4 * j + 2] += ds * t.z; dq[4 * j + 3] += ds * t.w;
      }
    }
#pragma unroll
    for (int j = 0; j < HD; ++j) dq[j] *= scale;
    const int64_t off = ((int64_t)b * a.L + q) * a.ld_g + head * HD;
    store_split32(a.dqr_hi + off, a.dqr_lo + off, dq);
  }
  {
    float dot = 0.0f;
    for (int h = 0; h < a.H; ++h) dot += acs[h * T + tid] * dac[h * T + tid];
    float dq[HD];
#pragma unroll
    for (int j = 0; j < HD; ++j) dq[j] = 0.0f;
    for (int h = 0; h < a.H; ++h) {
      const float ds = acs[h * T + tid] * (dac[h * T + tid] - dot);
      a.dsc[(bh * a.H + h) * a.L + q] = ds;
      const float4* kp = reinterpret_cast<const float4*>(Ksc + h * HD);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = kp[j];
        dq[4 * j] += ds * t.x; dq[4 * j + 1] += ds * t.y; dq[4 * j + 2] += ds * t.z; dq[4 * j + 3] += ds * t.w;
      }
    }
#pragma unroll
    for (int j = 0; j < HD; ++j) dq[j] *= scale;
    const int64_t off = ((int64_t)b * a.L + q) * a.ld_g + head * HD;
    store_split32(a.dqc_hi + off, a.dqc_lo + off, dq);
  }
}

// ------------------------------------------------------------------ backward 2 (per key position): dV
// dV[h,w,:] = sum_q A_c[q,h] A_r[q,w] dO[q,:];  grid (ceil(HW/T), nh, B); queries staged in tiles of TQ.
constexpr int TQ = 64;
__global__ void rcda_bwd_v_kernel(const RcdaArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int T = blockDim.x, tid = threadIdx.x;
  const int head = blockIdx.y, b = blockIdx.z;
  float* dos = sm;                          // [TQ][32]
  float* ars = dos + TQ * HD;               // [W][TQ+1]
  float* acs = ars + a.W * (TQ + 1);        // [H][TQ+1]
  const int p = blockIdx.x * T + tid;
  const bool ok = p < a.H * a.W;
  const int h = ok ? p / a.W : 0, w = ok ? p % a.W : 0;
  const int64_t bh = (int64_t)b * a.nh + head;
  float acc[HD];
#pragma unroll
  for (int j = 0; j < HD; ++j) acc[j] = 0.0f;
  for (int q0 = 0; q0 < a.L; q0 += TQ) {
    const int nq = min(TQ, a.L - q0);
    __syncthreads();
    for (int i = tid; i < TQ * 8; i += T) {
      const int qq = i >> 3, c4 = i & 7;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (qq < nq)
        t = __ldg(reinterpret_cast<const float4*>(a.d_o + ((int64_t)b * a.L + q0 + qq) * a.E + head * HD + c4 * 4));
      reinterpret_cast<float4*>(dos)[i] = t;
    }
    for (int i = tid; i < a.W * TQ; i += T) {
      const int ww = i / TQ, qq = i % TQ;
      ars[ww * (TQ + 1) + qq] = qq < nq ? a.ar[(bh * a.W + ww) * a.L + q0 + qq] : 0.0f;
    }
    for (int i = tid; i < a.H * TQ; i += T) {
      const int hh = i / TQ, qq = i % TQ;
      acs[hh * (TQ + 1) + qq] = qq < nq ? a.ac[(bh * a.H + hh) * a.L + q0 + qq] : 0.0f;
    }
    __syncthreads();
    for (int qq = 0; qq < nq; ++qq) {
      const float coef = acs[h * (TQ + 1) + qq] * ars[w * (TQ + 1) + qq];
      const float4* dp = reinterpret_cast<const float4*>(dos + qq * HD);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = dp[j];
        acc[4 * j] += coef * t.x; acc[4 * j + 1] += coef * t.y;
        acc[4 * j + 2] += coef * t.z; acc[4 * j + 3] += coef * t.w;
      }
    }
  }
  if (ok) {
    const int64_t off = (((int64_t)b * a.H + h) * a.W + w) * a.ld_g + head * HD;
    store_split32(a.dv_hi + off, a.dv_lo + off, acc);
  }
}

// ------------------------------------------------------------------ backward 3: dK_r / dK_c
// dK[k,d] = s * sum_q dS[k,q] * q[q,d];  grid (nh, B, 2): one CTA per (head, sample, row/column side).
constexpr int KMAX = 64;   // max keys per side handled by this kernel (H, W <= 64)
// On the tensor cores (mma.sync m16n8k16, split-bf16 3-pass): D[k, d] = sum_q dS[k, q] q[q, d] is a
// [n x L] x [L x 32] product per (sample, head, side).  The 8 warps of the CTA split the query range in steps of
// 16; fragments are loaded straight from global memory (dS rows are q-contiguous = row-major A, q rows give the
// "col" B operand pairwise), converted to hi/lo in registers, and the partial [n x 32] tiles are combined through
// shared-memory atomics.  ~100 instructions per 16 queries and warp instead of ~1500 FMAs + LDS.
__device__ __forceinline__ void mma_bf16_k(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int MT>   // key rows handled = 16 * MT (n <= 16 * MT)
__global__ void __launch_bounds__(256) rcda_bwd_k_mma_kernel(const RcdaArgs a) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[16 * MT][HD + 1];
  const int head = blockIdx.x, b = blockIdx.y, which = blockIdx.z;
  const int n = which == 0 ? a.W : a.H;
  const float* ds = which == 0 ? a.dsr : a.dsc;
  const float* qp = which == 0 ? a.qr : a.qc;
  __nv_bfloat16* ohi = which == 0 ? a.dkr_hi : a.dkc_hi;
  __nv_bfloat16* olo = which == 0 ? a.dkr_lo : a.dkc_lo;
  const int64_t bh = (int64_t)b * a.nh + head;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int i = threadIdx.x; i < 16 * MT * (HD + 1); i += 256) (&red[0][0])[i] = 0.0f;
  __syncthreads();
  float acc[MT][4][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.0f;
  const float* dsb = ds + bh * n * a.L;
  const float* qb = qp + (int64_t)b * a.L * a.E + head * HD;
  for (int q0 = warp * 16; q0 < a.L; q0 += 8 * 16) {
    // B fragments: (k = query 2t, 2t+1 [+8]; n = channel g) for the four 8-channel tiles
    uint32_t bh_[4][2], bl_[4][2];
    const int qa = q0 + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int c = nt * 8 + g;
      const float x0 = qa < a.L ? __ldg(qb + (int64_t)qa * a.E + c) : 0.0f;
      const float x1 = qa + 1 < a.L ? __ldg(qb + (int64_t)(qa + 1) * a.E + c) : 0.0f;
      const float x2 = qa + 8 < a.L ? __ldg(qb + (int64_t)(qa + 8) * a.E + c) : 0.0f;
      const float x3 = qa + 9 < a.L ? __ldg(qb + (int64_t)(qa + 9) * a.E + c) : 0.0f;
      split_bf16_pair(x0, x1, bh_[nt][0], bl_[nt][0]);
      split_bf16_pair(x2, x3, bh_[nt][1], bl_[nt][1]);
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      const float* p0 = dsb + (int64_t)r0 * a.L;
      const float* p1 = dsb + (int64_t)r1 * a.L;
      const bool k0 = r0 < n, k1 = r1 < n;
      const float y00 = (k0 && qa < a.L) ? __ldg(p0 + qa) : 0.0f, y01 = (k0 && qa + 1 < a.L) ? __ldg(p0 + qa + 1) : 0.0f;
      const float y10 = (k1 && qa < a.L) ? __ldg(p1 + qa) : 0.0f, y11 = (k1 && qa + 1 < a.L) ? __ldg(p1 + qa + 1) : 0.0f;
      const float y02 = (k0 && qa + 8 < a.L) ? __ldg(p0 + qa + 8) : 0.0f, y03 = (k0 && qa + 9 < a.L) ? __ldg(p0 + qa + 9) : 0.0f;
      const float y12 = (k1 && qa + 8 < a.L) ? __ldg(p1 + qa + 8) : 0.0f, y13 = (k1 && qa + 9 < a.L) ? __ldg(p1 + qa + 9) : 0.0f;
      uint32_t ah[4], al[4];
      split_bf16_pair(y00, y01, ah[0], al[0]);
      split_bf16_pair(y10, y11, ah[1], al[1]);
      split_bf16_pair(y02, y03, ah[2], al[2]);
      split_bf16_pair(y12, y13, ah[3], al[3]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        mma_bf16_k(acc[mt][nt], ah, bl_[nt][0], bl_[nt][1]);
        mma_bf16_k(acc[mt][nt], al, bh_[nt][0], bh_[nt][1]);
        mma_bf16_k(acc[mt][nt], ah, bh_[nt][0], bh_[nt][1]);
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int r = mt * 16 + g, c = nt * 8 + 2 * t;
      atomicAdd(&red[r][c], acc[mt][nt][0]);
      atomicAdd(&red[r][c + 1], acc[mt][nt][1]);
      atomicAdd(&red[r + 8][c], acc[mt][nt][2]);
      atomicAdd(&red[r + 8][c + 1], acc[mt][nt][3]);
    }
  __syncthreads();
  const float scale = rsqrtf((float)HD);
  for (int i = threadIdx.x; i < n * HD; i += 256) {
    const int k = i / HD, d = i % HD;
    const int64_t off = ((int64_t)b * n + k) * a.ld_g + head * HD + d;
    split_bf16(red[k][d] * scale, ohi[off], olo[off]);
  }
}
int launch_bwd_k(const RcdaArgs& a, cudaStream_t s) {
  const int n = a.H > a.W ? a.H : a.W;
  if (n <= 32) launch_light(rcda_bwd_k_mma_kernel<2>, dim3(a.nh, a.B, 2), dim3(256), 0, s, a);
  else launch_light(rcda_bwd_k_mma_kernel<4>, dim3(a.nh, a.B, 2), dim3(256), 0, s, a);
  return 0;
}

size_t fwd_smem(int H, int W, int T, int Hc) {
  return sizeof(float) * ((size_t)(W + H) * HD + (size_t)(W + H) * T + (size_t)Hc * W * HD);
}
size_t bwdq_smem(int H, int W, int T, int Hc) {
  return sizeof(float) * ((size_t)(W + H) * HD + 2 * (size_t)(W + H) * T + (size_t)Hc * W * HD);
}

}  // namespace

// C-ABI descriptor mirrors RcdaArgs with split views.
extern "C" int cdetr_rcda_fwd(int B, int L, int H, int W, int E, int nh, const float* qr, const float* qc,
                              const float* kr, const float* kc, const float* v, const uint8_t* mask_row,
                              const uint8_t* mask_col, float* ar, float* ac, cdetr_split_t o,
                              cdetr_stream_t s) {
  CDETR_CHECK_ARG(E == nh * HD, "rcda_fwd: head dim must be 32 (E=%d nh=%d)", E, nh);
  CDETR_CHECK_ARG(qr && qc && kr && kc && v && ar && ac && o.base, "rcda_fwd: null pointer");
  RcdaArgs a = {};
  a.B = B; a.L = L; a.H = H; a.W = W; a.E = E; a.nh = nh;
  a.qr = qr; a.qc = qc; a.kr = kr; a.kc = kc; a.v = v; a.mask_row = mask_row; a.mask_col = mask_col;
  a.ar = ar; a.ac = ac;
  a.o_hi = reinterpret_cast<__nv_bfloat16*>(o.base); a.o_lo = a.o_hi + o.plane; a.ld_o = o.ld;
  const size_t budget = 220 * 1024;
  int T = 256;
  while (T > 32 && fwd_smem(H, W, T, 1) > budget) T >>= 1;
  int Hc = H;
  while (Hc > 1 && fwd_smem(H, W, T, Hc) > budget) --Hc;
  CDETR_CHECK_ARG(fwd_smem(H, W, T, Hc) <= budget, "rcda_fwd: H=%d W=%d do not fit shared memory", H, W);
  a.Hc = Hc;
  const size_t smem = fwd_smem(H, W, T, Hc);
  { static DevAttrCache cfg = {}; CDETR_CHECK_CUDA(cdetr_ensure_smem(rcda_fwd_kernel, 227 * 1024, &cfg)); }
  dim3 grid(cdiv(L, T), nh, B);
  rcda_fwd_kernel<<<grid, T, smem, reinterpret_cast<cudaStream_t>(s)>>>(a);
  CDETR_CHECK_LAUNCH();
  return 0;
}

extern "C" int cdetr_rcda_bwd(int B, int L, int H, int W, int E, int nh, const float* qr, const float* qc,
                              const float* kr, const float* kc, const float* v, const float* ar,
                              const float* ac, const float* d_o, float* dsr, float* dsc, cdetr_split_t dqr,
                              cdetr_split_t dqc, cdetr_split_t dkr, cdetr_split_t dkc, cdetr_split_t dv,
                              cdetr_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  CDETR_CHECK_ARG(E == nh * HD, "rcda_bwd: head dim must be 32");
  CDETR_CHECK_ARG(H <= KMAX && W <= KMAX, "rcda_bwd: H, W must be <= 64");
  CDETR_CHECK_ARG(qr && qc && kr && kc && v && ar && ac && d_o && dsr && dsc && dqr.base && dqc.base &&
                      dkr.base && dkc.base && dv.base,
                  "rcda_bwd: null pointer");
  CDETR_CHECK_ARG(dqr.ld == dqc.ld && dqr.ld == dkr.ld && dqr.ld == dkc.ld && dqr.ld == dv.ld,
                  "rcda_bwd: gradient tensors must share ld");
  RcdaArgs a = {};
  a.B = B; a.L = L; a.H = H; a.W = W; a.E = E; a.nh = nh;
  a.qr = qr; a.qc = qc; a.kr = kr; a.kc = kc; a.v = v;
  a.ar = const_cast<float*>(ar); a.ac = const_cast<float*>(ac);
  a.d_o = d_o; a.dsr = dsr; a.dsc = dsc;
  auto hi = [](cdetr_split_t t) { return reinterpret_cast<__nv_bfloat16*>(t.base); };
  a.dqr_hi = hi(dqr); a.dqr_lo = hi(dqr) + dqr.plane;
  a.dqc_hi = hi(dqc); a.dqc_lo = hi(dqc) + dqc.plane;
  a.dkr_hi = hi(dkr); a.dkr_lo = hi(dkr) + dkr.plane;
  a.dkc_hi = hi(dkc); a.dkc_lo = hi(dkc) + dkc.plane;
  a.dv_hi = hi(dv); a.dv_lo = hi(dv) + dv.plane;
  a.ld_g = dqr.ld;
  const size_t budget = 220 * 1024;
  int T = 256;
  while (T > 32 && bwdq_smem(H, W, T, 1) > budget) T >>= 1;
  int Hc = H;
  while (Hc > 1 && bwdq_smem(H, W, T, Hc) > budget) --Hc;
  CDETR_CHECK_ARG(bwdq_smem(H, W, T, Hc) <= budget, "rcda_bwd: H=%d W=%d do not fit shared memory", H, W);
  a.Hc = Hc;
  { static DevAttrCache cfg = {}; CDETR_CHECK_CUDA(cdetr_ensure_smem(rcda_bwd_q_kernel, 227 * 1024, &cfg)); }
  rcda_bwd_q_kernel<<<dim3(cdiv(L, T), nh, B), T, bwdq_smem(H, W, T, Hc), s>>>(a);
  CDETR_CHECK_LAUNCH();
  const int TV = 256;
  const size_t smem_v = sizeof(float) * ((size_t)TQ * HD + (size_t)(W + H) * (TQ + 1));
  { static DevAttrCache cfg = {}; CDETR_CHECK_CUDA(cdetr_ensure_smem(rcda_bwd_v_kernel, 227 * 1024, &cfg)); }
  rcda_bwd_v_kernel<<<dim3(cdiv(H * W, TV), nh, B), TV, smem_v, s>>>(a);
  CDETR_CHECK_LAUNCH();
  launch_bwd_k(a, s);
  CDETR_CHECK_LAUNCH();
  return 0;
}

// dK_r / dK_c / dV given the dS maps (produced by cdetr_rcda_bwd or by the tensor-core query-side kernel).
extern "C" int cdetr_rcda_bwd_kv(int B, int L, int H, int W, int E, int nh, const float* qr, const float* qc,
                                 const float* ar, const float* ac, const float* d_o, const float* dsr,
                                 const float* dsc, cdetr_split_t dkr, cdetr_split_t dkc, cdetr_split_t dv,
                                 cdetr_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  CDETR_CHECK_ARG(E == nh * HD, "rcda_bwd_kv: head dim must be 32");
  CDETR_CHECK_ARG(H <= KMAX && W <= KMAX, "rcda_bwd_kv: H, W must be <= 64");
  CDETR_CHECK_ARG(qr && qc && ar && ac && d_o && dsr && dsc && dkr.base && dkc.base && dv.base, "rcda_bwd_kv: null pointer");
  CDETR_CHECK_ARG(dkr.ld == dkc.ld && dkr.ld == dv.ld, "rcda_bwd_kv: gradient tensors must share ld");
  RcdaArgs a = {};
  a.B = B; a.L = L; a.H = H; a.W = W; a.E = E; a.nh = nh;
  a.qr = qr; a.qc = qc; a.ar = const_cast<float*>(ar); a.ac = const_cast<float*>(ac);
  a.d_o = d_o; a.dsr = const_cast<float*>(dsr); a.dsc = const_cast<float*>(dsc);
  auto hi = [](cdetr_split_t t) { return reinterpret_cast<__nv_bfloat16*>(t.base); };
  a.dkr_hi = hi(dkr); a.dkr_lo = hi(dkr) + dkr.plane;
  a.dkc_hi = hi(dkc); a.dkc_lo = hi(dkc) + dkc.plane;
  a.dv_hi = hi(dv); a.dv_lo = hi(dv) + dv.plane;
  a.ld_g = dkr.ld;
  const int TV = 256;
  const size_t smem_v = sizeof(float) * ((size_t)TQ * HD + (size_t)(W + H) * (TQ + 1));
  { static DevAttrCache cfg = {}; CDETR_CHECK_CUDA(cdetr_ensure_smem(rcda_bwd_v_kernel, 227 * 1024, &cfg)); }
  rcda_bwd_v_kernel<<<dim3(cdiv(H * W, TV), nh, B), TV, smem_v, s>>>(a);
  CDETR_CHECK_LAUNCH();
  launch_bwd_k(a, s);
  CDETR_CHECK_LAUNCH();
  return 0;
}

// dK_r / dK_c only (the tensor-core path computes dV in cdetr_rcda_bwd_v_tc).
extern "C" int cdetr_rcda_bwd_k(int B, int L, int H, int W, int E, int nh, const float* qr, const float* qc,
                                const float* dsr, const float* dsc, cdetr_split_t dkr, cdetr_split_t dkc,
                                cdetr_stream_t s_) {
  CDETR_CHECK_ARG(E == nh * HD, "rcda_bwd_k: head dim must be 32");
  CDETR_CHECK_ARG(H <= KMAX && W <= KMAX, "rcda_bwd_k: H, W must be <= 64");
  CDETR_CHECK_ARG(qr && qc && dsr && dsc && dkr.base && dkc.base && dkr.ld == dkc.ld, "rcda_bwd_k: bad args");
  RcdaArgs a = {};
  a.B = B; a.L = L; a.H = H; a.W = W; a.E = E; a.nh = nh;
  a.qr = qr; a.qc = qc; a.dsr = const_cast<float*>(dsr); a.dsc = const_cast<float*>(dsc);
  a.dkr_hi = reinterpret_cast<__nv_bfloat16*>(dkr.base); a.dkr_lo = a.dkr_hi + dkr.plane;
  a.dkc_hi = reinterpret_cast<__nv_bfloat16*>(dkc.base); a.dkc_lo = a.dkc_hi + dkc.plane;
  a.ld_g = dkr.ld;
  launch_bwd_k(a, reinterpret_cast<cudaStream_t>(s_));
  CDETR_CHECK_LAUNCH();
  return 0;
}
