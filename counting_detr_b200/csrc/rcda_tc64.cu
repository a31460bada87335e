// RCDA forward and query-side backward on tcgen05 tensor cores for feature maps up to 64 x 64 (800 x 800 inputs:
// 50 x 50, BASELINE config C4), where a head's value slice (H*W*32 split-bf16 = 2 x 160 KB) no longer fits shared
// memory: V is STREAMED in blocks of 4 key rows through a 3-stage TMA ring while the contraction index (the key
// column w) is padded to 64.  Same math and same C ABI as rcda_tc.cu (which keeps V resident for maps <= 32 x 32):
//
//   O[q,:] = sum_h A_c[q,h] * ( sum_w A_r[q,w] * V[h,w,:] )           A2/models/row_column_decoupled_attention.py:210-291
//
//   warp 0      TMA producer: V[4 rows][64 cols][32 ch] hi + lo planes per stage (SWIZZLE_64B, OOB rows/cols -> 0)
//   warp 1      MMA issuer (one thread): logits S = q K^T ([128 x 32] x [32 x 64 keys]) per side and query tile, then
//               per V block D[128 x 128] = A_r[128 x 64] * V[64 x (4*32)] (forward; two TMEM buffers per query tile) or
//               G[128 x 256] = dO[128 x 32] * V^T[32 x (4*64)] (backward); three bf16 passes hi*lo + lo*hi + hi*hi
//   warps 2-9   two groups of four warps = two 128-query tiles, one query per thread (TMEM lane = query): softmaxes in
//               fp32, A_r -> split-bf16 SWIZZLE_128B operand, epilogue acc += A_c[q,h] * D[q,h,:] (forward) or
//               dA_c[h] = sum_w A_r G, dA_r[w] += A_c[h] G (backward), both softmax backwards, dq = dS K on the tensor core
// The reference's [B*heads, L, W, 32] intermediate (1.0 GB per encoder layer at 800 x 800) exists in TMEM only.
#include "common.cuh"
#include "../../include/cdetr.h"

namespace {

constexpr int HD = 32;      // head dim
constexpr int TQ = 128;     // queries per tile = TMEM lanes
constexpr int KP = 64;      // padded keys per side (= MMA K of the forward contraction)
constexpr int HB = 4;       // key rows per streamed V block
constexpr int NSTAGE = 3;
constexpr uint32_t V_PLANE = HB * KP * HD * 2;     // 16 KB: one plane of a V block, [4][64][32] bf16
constexpr uint32_t V_STAGE = 2 * V_PLANE;          // hi + lo
constexpr uint32_t AR_PLANE = TQ * KP * 2;         // 16 KB: [128][64] bf16, K-major SWIZZLE_128B
constexpr uint32_t Q_PLANE = TQ * HD * 2;          // 8 KB:  [128][32] bf16, K-major SWIZZLE_64B
constexpr uint32_t K_PLANE = KP * HD * 2;          // 4 KB:  [64 keys][32] bf16, SWIZZLE_64B

struct Fwd64Args {
  int B, L, H, W, E, nh;
  const float *qr, *qc, *kr, *kc;
  const uint8_t *mask_row, *mask_col;
  float *ar, *ac;              // [B,nh,W,L], [B,nh,H,L]
  __nv_bfloat16 *o_hi, *o_lo;
  int64_t ld_o;
  uint32_t idesc, idesc_s;
};

struct Bwd64Args {
  int B, L, H, W, E, nh;
  const float *kr, *kc, *ar, *ac, *d_o;
  float *dsr, *dsc;            // [B,nh,W,L], [B,nh,H,L]
  __nv_bfloat16 *dqr_hi, *dqr_lo, *dqc_hi, *dqc_lo;
  int64_t ld_g;
  uint32_t idesc, idesc_q;
};

// softmax over the first n of 64 logits (entries >= n or masked get probability 0)
__device__ __forceinline__ void softmax_row64(float (&p)[KP], int n, const uint8_t* __restrict__ mask) {
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    if (k >= n || (mask && mask[k])) p[k] = -INFINITY;
    mx = fmaxf(mx, p[k]);
  }
  float sum = 0.0f;
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    p[k] = expf(p[k] - mx);
    sum += p[k];
  }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int k = 0; k < KP; ++k) p[k] *= inv;
}
// 32 values -> split-bf16 K-major SWIZZLE_64B operand row r (64-byte rows)
__device__ __forceinline__ void write_row32_sw64(uint8_t* dst, uint32_t plane_bytes, int r, const float (&x)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_bf16_pair(x[8 * j + 2 * i], x[8 * j + 2 * i + 1], hw[i], lw[i]);
    const uint32_t off = (uint32_t)r * 64u + (uint32_t)((j ^ ((r >> 1) & 3)) << 4);
    sts128(smem_u32(dst + off), hw[0], hw[1], hw[2], hw[3]);
    sts128(smem_u32(dst + plane_bytes + off), lw[0], lw[1], lw[2], lw[3]);
  }
}
// 64 values -> split-bf16 K-major SWIZZLE_128B operand row r (128-byte rows, 16-byte chunk j at j ^ (r & 7))
__device__ __forceinline__ void write_row64_sw128(uint8_t* dst, uint32_t plane_bytes, int r, const float (&x)[KP]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_bf16_pair(x[8 * j + 2 * i], x[8 * j + 2 * i + 1], hw[i], lw[i]);
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
    sts128(smem_u32(dst + off), hw[0], hw[1], hw[2], hw[3]);
    sts128(smem_u32(dst + plane_bytes + off), lw[0], lw[1], lw[2], lw[3]);
  }
}
// the two key tiles K_r [W,32], K_c [H,32] of this (sample, head) -> [2 sides][2 planes][64 keys][32] SWIZZLE_64B
__device__ __forceinline__ void stage_key_tiles(uint8_t* Kb, const float* kr, const float* kc, int b, int head, int H, int W,
                                                int E, int ct) {
  for (int t = ct; t < 2 * KP * 4; t += 256) {
    const int side = t >> 8, k = (t >> 2) & (KP - 1), j = t & 3;
    const int n = side == 0 ? W : H;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0.0f;
    if (k < n) {
      const float* kp = (side == 0 ? kr : kc) + ((int64_t)b * n + k) * E + head * HD + j * 8;
      const float4 t0 = __ldg(reinterpret_cast<const float4*>(kp)), t1 = __ldg(reinterpret_cast<const float4*>(kp) + 1);
      x[0] = t0.x; x[1] = t0.y; x[2] = t0.z; x[3] = t0.w; x[4] = t1.x; x[5] = t1.y; x[6] = t1.z; x[7] = t1.w;
    }
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_bf16_pair(x[2 * i], x[2 * i + 1], hw[i], lw[i]);
    uint8_t* kd = Kb + (size_t)side * 2 * K_PLANE;
    const uint32_t off = (uint32_t)k * 64u + (uint32_t)((j ^ ((k >> 1) & 3)) << 4);
    sts128(smem_u32(kd + off), hw[0], hw[1], hw[2], hw[3]);
    sts128(smem_u32(kd + K_PLANE + off), lw[0], lw[1], lw[2], lw[3]);
  }
  fence_proxy_async();
}
__device__ __forceinline__ void store_split32_row(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&v)[32]) {
#pragma unroll
  for (int gg = 0; gg < 4; ++gg) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_bf16_pair(v[gg * 8 + 2 * j], v[gg * 8 + 2 * j + 1], hw[j], lw[j]);
    reinterpret_cast<uint4*>(hi)[gg] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    reinterpret_cast<uint4*>(lo)[gg] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

// =============================================================================================== forward
constexpr uint32_t FWD_SMEM = NSTAGE * V_STAGE + 4 * AR_PLANE + 4 * Q_PLANE + 4 * K_PLANE;   // 208 KB

__global__ void __launch_bounds__(320, 1)
rcda_fwd_tc64_kernel(const __grid_constant__ CUtensorMap tmV, const Fwd64Args a) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Vs = smem;                          // [NSTAGE][2 planes][4][64][32] bf16, SW64
  uint8_t* Ar = Vs + NSTAGE * V_STAGE;         // [2 tiles][2 planes][128][64] bf16, SW128 K-major (A_r operand)
  uint8_t* Qs = Ar + 4 * AR_PLANE;             // [2 tiles][2 planes][128][32] bf16, SW64 K-major (q_r, then q_c)
  uint8_t* Kb = Qs + 4 * Q_PLANE;              // [2 sides][2 planes][64][32] bf16, SW64
  uint64_t* bars = reinterpret_cast<uint64_t*>(Kb + 4 * K_PLANE);
  uint64_t* v_full = bars;            // [NSTAGE]
  uint64_t* v_empty = bars + 3;       // [NSTAGE]
  uint64_t* a_ready = bars + 6;       // [2]
  uint64_t* q_ready = bars + 8;       // [2]
  uint64_t* s_full = bars + 10;       // [2]
  uint64_t* t_full = bars + 12;       // [2 tiles][2 buffers]
  uint64_t* t_empty = bars + 16;      // [2 tiles][2 buffers]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int q_cta = blockIdx.x * 2 * TQ;
  const int nblk = (a.H + HB - 1) / HB;
  const int nks = (a.W + 15) / 16;             // k-steps of the contraction over w (columns >= W are zero on both sides)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&a_ready[g], 4);
      mbar_init(&q_ready[g], 4);
      mbar_init(&s_full[g], 1);
      for (int u = 0; u < 2; ++u) { mbar_init(&t_full[g * 2 + u], 1); mbar_init(&t_empty[g * 2 + u], 4); }
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int ngroups = (q_cta + TQ < a.L) ? 2 : 1;

  if (warp == 0) {
    if (lane == 0) {
      for (int p = 0; p < nblk; ++p) {
        const int st = p % NSTAGE;
        if (p >= NSTAGE) mbar_wait_single(&v_empty[st], (uint32_t)(p / NSTAGE - 1) & 1u);
        mbar_arrive_expect_tx(&v_full[st], V_STAGE);
        tma_load_5d(Vs + st * V_STAGE, &tmV, &v_full[st], head * HD, 0, p * HB, b, 0);          // box {32 c, 64 w, 4 h, 1, 1}
        tma_load_5d(Vs + st * V_STAGE + V_PLANE, &tmV, &v_full[st], head * HD, 0, p * HB, b, 1);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // logits S = q K^T: [128 x 32] x [32 x 64 keys] into the first 64 columns of the tile's TMEM region
      for (int side = 0; side < 2; ++side) {
        for (int g = 0; g < ngroups; ++g) {
          mbar_wait_single(&q_ready[g], (uint32_t)side);
          tc_fence_after();
          const uint32_t a_base = smem_u32(Qs) + (uint32_t)g * 2u * Q_PLANE;
          const uint32_t k_base = smem_u32(Kb) + (uint32_t)side * 2u * K_PLANE;
          const uint32_t d = tmem_base + (uint32_t)g * 256u;
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {
            const uint64_t a_hi = make_smem_desc(a_base + ks * 32, 16, 512, 4);
            const uint64_t a_lo = make_smem_desc(a_base + Q_PLANE + ks * 32, 16, 512, 4);
            const uint64_t b_hi = make_smem_desc(k_base + ks * 32, 16, 512, 4);
            const uint64_t b_lo = make_smem_desc(k_base + K_PLANE + ks * 32, 16, 512, 4);
            umma_bf16_ss(d, a_hi, b_lo, a.idesc_s, ks > 0 ? 1u : 0u);
            umma_bf16_ss(d, a_lo, b_hi, a.idesc_s, 1);
            umma_bf16_ss(d, a_hi, b_hi, a.idesc_s, 1);
          }
          umma_commit(&s_full[g]);
        }
      }
      // main contraction, V block by V block; two 128-column accumulator buffers per query tile
      for (int p = 0; p < nblk; ++p) {
        const int st = p % NSTAGE;
        mbar_wait_single(&v_full[st], (uint32_t)(p / NSTAGE) & 1u);
        const uint32_t v_base = smem_u32(Vs) + (uint32_t)st * V_STAGE;
        const int u = p & 1;
        for (int g = 0; g < ngroups; ++g) {
          if (p == 0) mbar_wait_single(&a_ready[g], 0);                 // A_r operand written AND both logit rows read out
          if (p >= 2) mbar_wait_single(&t_empty[g * 2 + u], (uint32_t)((p >> 1) - 1) & 1u);
          tc_fence_after();
          const uint32_t a_base = smem_u32(Ar) + (uint32_t)g * 2u * AR_PLANE;
          const uint32_t d = tmem_base + (uint32_t)g * 256u + (uint32_t)u * 128u;
          for (int ks = 0; ks < nks; ++ks) {
            // A_r: K-major SW128 (128-byte rows, 8-row groups 1024 B apart, k-step of 16 = +32 B)
            const uint64_t a_hi = make_smem_desc(a_base + ks * 32, 16, 1024, 2);
            const uint64_t a_lo = make_smem_desc(a_base + AR_PLANE + ks * 32, 16, 1024, 2);
            // V: MN-major SW64: MN chunk = one key row (32 channels, 64 B), chunks KP*64 B apart (LBO), 8 key columns
            // (the k index) per 512-byte group (SBO), k-step of 16 columns = +1024 B
            const uint32_t vb = v_base + ks * 1024;
            const uint64_t b_hi = make_smem_desc(vb, KP * 64, 512, 4);
            const uint64_t b_lo = make_smem_desc(vb + V_PLANE, KP * 64, 512, 4);
            umma_bf16_ss(d, a_hi, b_lo, a.idesc, ks > 0 ? 1u : 0u);
            umma_bf16_ss(d, a_lo, b_hi, a.idesc, 1);
            umma_bf16_ss(d, a_hi, b_hi, a.idesc, 1);
          }
          umma_commit(&t_full[g * 2 + u]);
        }
        umma_commit(&v_empty[st]);
      }
    }
  } else {
    const int ct = threadIdx.x - 64;       // 0..255
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;
    const int q = q_cta + g * TQ + r;
    const bool ok = q < a.L;
    const bool active = q_cta + g * TQ < a.L;
    stage_key_tiles(Kb, a.kr, a.kc, b, head, a.H, a.W, a.E, ct);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (active) {
      uint8_t* Qg = Qs + (size_t)g * 2 * Q_PLANE;
      uint8_t* Ag = Ar + (size_t)g * 2 * AR_PLANE;
      const float scale = rsqrtf((float)HD);
      const int64_t bh = (int64_t)b * a.nh + head;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)g * 256u;
      float ac[KP];
      {
        float qv[HD];
        // ---- row side: q_r (scaled) -> operand -> tensor core -> this query's 64 logits back from TMEM
#pragma unroll
        for (int j = 0; j < HD; ++j) qv[j] = 0.0f;
        if (ok) {
          const float4* qp = reinterpret_cast<const float4*>(a.qr + ((int64_t)b * a.L + q) * a.E + head * HD);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(qp + j);
            qv[4 * j] = t.x * scale; qv[4 * j + 1] = t.y * scale; qv[4 * j + 2] = t.z * scale; qv[4 * j + 3] = t.w * scale;
          }
        }
        write_row32_sw64(Qg, Q_PLANE, r, qv);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_ready[g]);
#pragma unroll
        for (int j = 0; j < HD; ++j) qv[j] = 0.0f;
        if (ok) {
          const float4* qp = reinterpret_cast<const float4*>(a.qc + ((int64_t)b * a.L + q) * a.E + head * HD);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(qp + j);
            qv[4 * j] = t.x * scale; qv[4 * j + 1] = t.y * scale; qv[4 * j + 2] = t.z * scale; qv[4 * j + 3] = t.w * scale;
          }
        }
        float ar[KP];
        mbar_wait(&s_full[g], 0);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t t[32];
          tmem_ld_32x32b_x32(taddr + (uint32_t)half * 32u, t);
          tmem_ld_wait();
#pragma unroll
          for (int w = 0; w < 32; ++w) ar[half * 32 + w] = __uint_as_float(t[w]);
        }
        tc_fence_before();
        // the row-logit MMAs have consumed q_r: the operand buffer takes q_c
        write_row32_sw64(Qg, Q_PLANE, r, qv);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_ready[g]);
        if (ok) {
          softmax_row64(ar, a.W, a.mask_row ? a.mask_row + (int64_t)b * a.W : nullptr);
#pragma unroll
          for (int w = 0; w < KP; ++w)
            if (w < a.W) a.ar[(bh * a.W + w) * a.L + q] = ar[w];
        } else {
#pragma unroll
          for (int w = 0; w < KP; ++w) ar[w] = 0.0f;
        }
        write_row64_sw128(Ag, AR_PLANE, r, ar);
        fence_proxy_async();
        // column logits (ar is dead from here on)
        mbar_wait(&s_full[g], 1);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t t[32];
          tmem_ld_32x32b_x32(taddr + (uint32_t)half * 32u, t);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 32; ++h) ac[half * 32 + h] = __uint_as_float(t[h]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[g]);     // A_r operand in place, logit columns of the TMEM region free
        if (ok) {
          softmax_row64(ac, a.H, a.mask_col ? a.mask_col + (int64_t)b * a.H : nullptr);
#pragma unroll
          for (int h = 0; h < KP; ++h)
            if (h < a.H) a.ac[(bh * a.H + h) * a.L + q] = ac[h];
        } else {
#pragma unroll
          for (int h = 0; h < KP; ++h) ac[h] = 0.0f;
        }
      }
      float acc[HD];
#pragma unroll
      for (int c = 0; c < HD; ++c) acc[c] = 0.0f;
#pragma unroll
      for (int p = 0; p < KP / HB; ++p) {
        if (p < nblk) {
          const int u = p & 1;
          mbar_wait(&t_full[g * 2 + u], (uint32_t)(p >> 1) & 1u);
          tc_fence_after();
#pragma unroll
          for (int hi = 0; hi < HB; ++hi) {
            uint32_t t[32];
            tmem_ld_32x32b_x32(taddr + (uint32_t)u * 128u + (uint32_t)hi * 32u, t);
            tmem_ld_wait();
            const float coef = ac[p * HB + hi];
            const float2 coef2 = make_float2(coef, coef);
#pragma unroll
            for (int c = 0; c < HD; c += 2) {   // packed FFMA2: two fp32 FMAs per issue slot, same rounding
              const float2 r2 = __ffma2_rn(coef2, make_float2(__uint_as_float(t[c]), __uint_as_float(t[c + 1])),
                                           make_float2(acc[c], acc[c + 1]));
              acc[c] = r2.x; acc[c + 1] = r2.y;
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[g * 2 + u]);
        }
      }
      if (ok) {
        const int64_t off = ((int64_t)b * a.L + q) * a.ld_o + head * HD;
        store_split32_row(a.o_hi + off, a.o_lo + off, acc);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =============================================================================================== backward (query side)
//   G[q,(h,w)] = sum_c dO[q,c] V[h,w,c]   (tcgen05: [128 x 32] x [32 x (4*64)] per V block)
//   dA_c[h] = sum_w A_r[w] G[h,w],  dA_r[w] += A_c[h] G[h,w];  dS = A o (dA - <A, dA>);  dq = s * dS K (tcgen05)
// dA_c[h] is parked in this query's slot of the dsc output (same thread writes and re-reads it) so that only A_r / dA_r
// (64 + 64 registers) stay live through the V stream.
constexpr uint32_t BWD_SMEM = NSTAGE * V_STAGE + 4 * Q_PLANE;    // 128 KB; K tiles + dS operands reuse the V ring

__global__ void __launch_bounds__(320, 1)
rcda_bwd_q_tc64_kernel(const __grid_constant__ CUtensorMap tmV, const Bwd64Args a) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Vs = smem;                          // V ring; afterwards: key tiles (16 KB) + dS operands (64 KB)
  uint8_t* As = Vs + NSTAGE * V_STAGE;         // dO operand [2 tiles][2 planes][128][32] bf16, SW64 K-major
  uint8_t* Kb = Vs;
  uint8_t* Ds = Vs + 4 * K_PLANE;              // [2 tiles][2 planes][128][64] bf16, SW128 K-major (dS rows)
  uint64_t* bars = reinterpret_cast<uint64_t*>(As + 4 * Q_PLANE);
  uint64_t* v_full = bars;            // [NSTAGE]
  uint64_t* v_empty = bars + 3;       // [NSTAGE]
  uint64_t* a_ready = bars + 6;       // [2]
  uint64_t* q_ready = bars + 8;       // [2]
  uint64_t* s_full = bars + 10;       // [2]
  uint64_t* t_full = bars + 12;       // [2]
  uint64_t* t_empty = bars + 14;      // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int q_cta = blockIdx.x * 2 * TQ;
  const int nblk = (a.H + HB - 1) / HB;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&a_ready[g], 4);
      mbar_init(&q_ready[g], 4);
      mbar_init(&s_full[g], 1);
      mbar_init(&t_full[g], 1);
      mbar_init(&t_empty[g], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int ngroups = (q_cta + TQ < a.L) ? 2 : 1;

  if (warp == 0) {
    if (lane == 0) {
      for (int p = 0; p < nblk; ++p) {
        const int st = p % NSTAGE;
        if (p >= NSTAGE) mbar_wait_single(&v_empty[st], (uint32_t)(p / NSTAGE - 1) & 1u);
        mbar_arrive_expect_tx(&v_full[st], V_STAGE);
        tma_load_5d(Vs + st * V_STAGE, &tmV, &v_full[st], head * HD, 0, p * HB, b, 0);
        tma_load_5d(Vs + st * V_STAGE + V_PLANE, &tmV, &v_full[st], head * HD, 0, p * HB, b, 1);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int p = 0; p < nblk; ++p) {
        const int st = p % NSTAGE;
        mbar_wait_single(&v_full[st], (uint32_t)(p / NSTAGE) & 1u);
        const uint32_t v_base = smem_u32(Vs) + (uint32_t)st * V_STAGE;
        for (int g = 0; g < ngroups; ++g) {
          if (p == 0) mbar_wait_single(&a_ready[g], 0);
          else mbar_wait_single(&t_empty[g], (uint32_t)(p - 1) & 1u);
          tc_fence_after();
          const uint32_t a_base = smem_u32(As) + (uint32_t)g * 2u * Q_PLANE;
          const uint32_t d = tmem_base + (uint32_t)g * 256u;
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {
            const uint64_t a_hi = make_smem_desc(a_base + ks * 32, 16, 512, 4);
            const uint64_t a_lo = make_smem_desc(a_base + Q_PLANE + ks * 32, 16, 512, 4);
            // V as K-major B: row n = h_local*64 + w at 64-byte pitch (8-row groups 512 B apart), k-step of 16 c = +32 B
            const uint64_t b_hi = make_smem_desc(v_base + ks * 32, 16, 512, 4);
            const uint64_t b_lo = make_smem_desc(v_base + V_PLANE + ks * 32, 16, 512, 4);
            umma_bf16_ss(d, a_hi, b_lo, a.idesc, ks > 0 ? 1u : 0u);
            umma_bf16_ss(d, a_lo, b_hi, a.idesc, 1);
            umma_bf16_ss(d, a_hi, b_hi, a.idesc, 1);
          }
          umma_commit(&t_full[g]);
        }
        umma_commit(&v_empty[st]);
      }
      // dq = dS K per side and tile: A = dS rows [128 x 64 keys] (K-major SW128), B = key tile [64 keys x 32] read as an
      // MN-major operand (row = key = k index, 64-byte rows), D = first 32 columns of the tile's TMEM region
      for (int side = 0; side < 2; ++side) {
        const int nks = ((side == 0 ? a.W : a.H) + 15) / 16;
        for (int g = 0; g < ngroups; ++g) {
          mbar_wait_single(&q_ready[g], (uint32_t)side);
          tc_fence_after();
          const uint32_t a_base = smem_u32(Ds) + (uint32_t)g * 2u * AR_PLANE;
          const uint32_t k_base = smem_u32(Kb) + (uint32_t)side * 2u * K_PLANE;
          const uint32_t d = tmem_base + (uint32_t)g * 256u;
          for (int ks = 0; ks < nks; ++ks) {
            const uint64_t a_hi = make_smem_desc(a_base + ks * 32, 16, 1024, 2);
            const uint64_t a_lo = make_smem_desc(a_base + AR_PLANE + ks * 32, 16, 1024, 2);
            const uint64_t b_hi = make_smem_desc(k_base + ks * 1024, K_PLANE, 512, 4);
            const uint64_t b_lo = make_smem_desc(k_base + K_PLANE + ks * 1024, K_PLANE, 512, 4);
            umma_bf16_ss(d, a_hi, b_lo, a.idesc_q, ks > 0 ? 1u : 0u);
            umma_bf16_ss(d, a_lo, b_hi, a.idesc_q, 1);
            umma_bf16_ss(d, a_hi, b_hi, a.idesc_q, 1);
          }
          umma_commit(&s_full[g]);
        }
      }
    }
  } else {
    const int ct = threadIdx.x - 64;
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int q = q_cta + g * TQ + r;
    const bool ok = q < a.L;
    const bool active = q_cta + g * TQ < a.L;
    const int64_t bh = (int64_t)b * a.nh + head;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)g * 256u;
    float ar[KP], dar[KP];
    if (active) {
      uint8_t* Ag = As + (size_t)g * 2 * Q_PLANE;
      {
        float dov[HD];
#pragma unroll
        for (int j = 0; j < HD; ++j) dov[j] = 0.0f;
        if (ok) {
          const float4* dp = reinterpret_cast<const float4*>(a.d_o + ((int64_t)b * a.L + q) * a.E + head * HD);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(dp + j);
            dov[4 * j] = t.x; dov[4 * j + 1] = t.y; dov[4 * j + 2] = t.z; dov[4 * j + 3] = t.w;
          }
        }
        write_row32_sw64(Ag, Q_PLANE, r, dov);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[g]);
      }
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        ar[k] = (ok && k < a.W) ? __ldg(a.ar + (bh * a.W + k) * a.L + q) : 0.0f;
        dar[k] = 0.0f;
      }
      for (int p = 0; p < nblk; ++p) {
        float cc[HB];
#pragma unroll
        for (int hi = 0; hi < HB; ++hi) {
          const int h = p * HB + hi;
          cc[hi] = (ok && h < a.H) ? __ldg(a.ac + (bh * a.H + h) * a.L + q) : 0.0f;
        }
        mbar_wait(&t_full[g], (uint32_t)p & 1u);
        tc_fence_after();
#pragma unroll
        for (int hi = 0; hi < HB; ++hi) {
          float2 s2 = make_float2(0.0f, 0.0f);
          const float2 cc2 = make_float2(cc[hi], cc[hi]);
#pragma unroll
          for (int part = 0; part < 4; ++part) {
            uint32_t t[16];
            tmem_ld_32x32b_x16(taddr + (uint32_t)hi * 64u + (uint32_t)part * 16u, t);
            tmem_ld_wait();
#pragma unroll
            for (int w = 0; w < 16; w += 2) {   // packed FFMA2: s += A_r pair * G pair; dA_r pair += A_c * G pair
              const int wi = part * 16 + w;
              const float2 g2 = make_float2(__uint_as_float(t[w]), __uint_as_float(t[w + 1]));
              s2 = __ffma2_rn(make_float2(ar[wi], ar[wi + 1]), g2, s2);
              const float2 d2 = __ffma2_rn(cc2, g2, make_float2(dar[wi], dar[wi + 1]));
              dar[wi] = d2.x; dar[wi + 1] = d2.y;
            }
          }
          const int h = p * HB + hi;
          if (ok && h < a.H) a.dsc[(bh * a.H + h) * a.L + q] = s2.x + s2.y;    // dA_c[h], parked (see header)
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[g]);
      }
    }
    // every G MMA has completed once both tiles are past their last t_full wait: the V ring takes the key tiles and the
    // dS operands; dq = scale * dS K then runs on the tensor core and comes back one query per thread
    asm volatile("bar.sync 1, 256;" ::: "memory");
    stage_key_tiles(Kb, a.kr, a.kc, b, head, a.H, a.W, a.E, ct);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (active) {
      uint8_t* Dg = Ds + (size_t)g * 2 * AR_PLANE;
      const float scale = rsqrtf((float)HD);
      const int64_t off = ((int64_t)b * a.L + q) * a.ld_g + head * HD;
      {   // row side: dS_r = A_r o (dA_r - <A_r, dA_r>)   (in place in dar)
        float dot = 0.0f;
#pragma unroll
        for (int w = 0; w < KP; ++w) dot += ar[w] * dar[w];
#pragma unroll
        for (int w = 0; w < KP; ++w) {
          dar[w] = ar[w] * (dar[w] - dot);
          if (ok && w < a.W) a.dsr[(bh * a.W + w) * a.L + q] = dar[w];
        }
      }
      write_row64_sw128(Dg, AR_PLANE, r, dar);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_ready[g]);
      {   // column side while the tensor core works on the row side: A_c and the parked dA_c come back (ar / dar reused)
        float dot = 0.0f;
#pragma unroll
        for (int h = 0; h < KP; ++h) {
          const bool in = ok && h < a.H;
          ar[h] = in ? __ldg(a.ac + (bh * a.H + h) * a.L + q) : 0.0f;
          dar[h] = in ? a.dsc[(bh * a.H + h) * a.L + q] : 0.0f;
          dot += ar[h] * dar[h];
        }
#pragma unroll
        for (int h = 0; h < KP; ++h) {
          dar[h] = ar[h] * (dar[h] - dot);
          if (ok && h < a.H) a.dsc[(bh * a.H + h) * a.L + q] = dar[h];
        }
      }
      mbar_wait(&s_full[g], 0);
      tc_fence_after();
      {
        uint32_t t[32];
        tmem_ld_32x32b_x32(taddr, t);
        tmem_ld_wait();
        tc_fence_before();
        // the row-side MMAs have consumed dS_r: the operand buffer takes dS_c
        write_row64_sw128(Dg, AR_PLANE, r, dar);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_ready[g]);
        if (ok) {
          float v[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(t[c]) * scale;
          store_split32_row(a.dqr_hi + off, a.dqr_lo + off, v);
        }
      }
      mbar_wait(&s_full[g], 1);
      tc_fence_after();
      {
        uint32_t t[32];
        tmem_ld_32x32b_x32(taddr, t);
        tmem_ld_wait();
        if (ok) {
          float v[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(t[c]) * scale;
          store_split32_row(a.dqc_hi + off, a.dqc_lo + off, v);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn64() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// V as a 5-D tensor (channel, w, h, sample, plane); one box = 32 channels of a head x 64 key columns x 4 key rows
int make_v_map64(CUtensorMap* tm, const cdetr_split_t& v, int B, int H, int W, int E) {
  EncodeTiledFn fn = encode_fn64();
  if (!fn) { cdetr_set_error("cuTensorMapEncodeTiled entry point unavailable"); return CDETR_ERR_CUDA; }
  cuuint64_t gdim[5] = {(cuuint64_t)E, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 2};
  cuuint64_t gstr[4] = {(cuuint64_t)v.ld * 2, (cuuint64_t)W * v.ld * 2, (cuuint64_t)H * W * v.ld * 2,
                        (cuuint64_t)v.plane * 2};
  cuuint32_t box[5] = {HD, KP, HB, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.base, gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { cdetr_set_error("rcda tc64: cuTensorMapEncodeTiled failed (%d)", (int)r); return CDETR_ERR_CUDA; }
  return 0;
}

}  // namespace

// Called by cdetr_rcda_fwd_tc / cdetr_rcda_bwd_q_tc (rcda_tc.cu) when 32 < max(H, W) <= 64.
int rcda_fwd_tc64_launch(int B, int L, int H, int W, int E, int nh, const float* qr, const float* qc, const float* kr,
                         const float* kc, cdetr_split_t v, const uint8_t* mask_row, const uint8_t* mask_col, float* ar,
                         float* ac, cdetr_split_t o, cudaStream_t s) {
  CUtensorMap tm;
  int rc = make_v_map64(&tm, v, B, H, W, E);
  if (rc) return rc;
  Fwd64Args a = {};
  a.B = B; a.L = L; a.H = H; a.W = W; a.E = E; a.nh = nh;
  a.qr = qr; a.qc = qc; a.kr = kr; a.kc = kc; a.mask_row = mask_row; a.mask_col = mask_col; a.ar = ar; a.ac = ac;
  a.o_hi = reinterpret_cast<__nv_bfloat16*>(o.base); a.o_lo = a.o_hi + o.plane; a.ld_o = o.ld;
  a.idesc = make_idesc_bf16_f32(TQ, HB * HD, 0, 1);
  a.idesc_s = make_idesc_bf16_f32(TQ, KP, 0, 0);
  const int smem = (int)FWD_SMEM + 256 + 1024;
  static DevAttrCache cfg = {};
  CDETR_CHECK_CUDA(cdetr_ensure_smem(rcda_fwd_tc64_kernel, smem, &cfg));
  rcda_fwd_tc64_kernel<<<dim3(cdiv(L, 2 * TQ), nh, B), 320, smem, s>>>(tm, a);
  CDETR_CHECK_LAUNCH();
  return 0;
}

int rcda_bwd_q_tc64_launch(int B, int L, int H, int W, int E, int nh, const float* kr, const float* kc, cdetr_split_t v,
                           const float* ar, const float* ac, const float* d_o, float* dsr, float* dsc, cdetr_split_t dqr,
                           cdetr_split_t dqc, cudaStream_t s) {
  CUtensorMap tm;
  int rc = make_v_map64(&tm, v, B, H, W, E);
  if (rc) return rc;
  Bwd64Args a = {};
  a.B = B; a.L = L; a.H = H; a.W = W; a.E = E; a.nh = nh;
  a.kr = kr; a.kc = kc; a.ar = ar; a.ac = ac; a.d_o = d_o; a.dsr = dsr; a.dsc = dsc;
  a.dqr_hi = reinterpret_cast<__nv_bfloat16*>(dqr.base); a.dqr_lo = a.dqr_hi + dqr.plane;
  a.dqc_hi = reinterpret_cast<__nv_bfloat16*>(dqc.base); a.dqc_lo = a.dqc_hi + dqc.plane;
  a.ld_g = dqr.ld;
  a.idesc = make_idesc_bf16_f32(TQ, HB * KP, 0, 0);
  a.idesc_q = make_idesc_bf16_f32(TQ, HD, 0, 1);
  const int smem = (int)BWD_SMEM + 256 + 1024;
  static DevAttrCache cfg = {};
  CDETR_CHECK_CUDA(cdetr_ensure_smem(rcda_bwd_q_tc64_kernel, smem, &cfg));
  rcda_bwd_q_tc64_kernel<<<dim3(cdiv(L, 2 * TQ), nh, B), 320, smem, s>>>(tm, a);
  CDETR_CHECK_LAUNCH();
  return 0;
}
