// RCDA forward on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), one CTA per (sample, head, 2 x 128 queries).
//
//   O[q,:] = sum_h A_c[q,h] * ( sum_w A_r[q,w] * V[h,w,:] )
//            `--- CUDA cores, TMEM->register epilogue ---'  `--- tcgen05.mma: [128 x W] x [W x (H*32)] ---'
//
//   warp 0      TMA: the head's V slice (hi + lo planes, [H][W][32] bf16, SWIZZLE_64B) -> shared memory
//   warp 1      MMA issuer: for each block of 8 key rows h: D[128 x 256] = A_r[128 x W] * V[W x (8*32)], three
//               bf16 passes (hi*lo, lo*hi, hi*hi), fp32 accumulation in one of two 256-column TMEM buffers
//   warps 2-9   two groups of four warps, each group owns one 128-query tile and one 256-column TMEM buffer (the
//               tensor core works on one tile while the other tile's epilogue runs; V is staged once for both);
//               one query per thread.  The logits q K^T of both sides are tensor-core products too: the thread
//               writes its scaled q row as a split-bf16 A operand row, the MMA warp multiplies the group's
//               [128 x 32] operand with the [32 keys x 32] key tile (staged once as a split-bf16 B operand) into the
//               first 32 TMEM columns of the group's buffer and the thread reads its logit row back
//               (tcgen05.ld.x32); the two softmaxes stay fp32 on the CUDA cores, A_r goes to shared memory as the
//               split-bf16 A operand of the main contraction (manual SWIZZLE_64B), then the epilogue:
//               tcgen05.ld of T[q, h, 0:32] and acc += A_c[q,h] * T   (the only CUDA-core FMAs: 1/32 of the MACs)
// The reference materialises [B*heads, L, W, 32] in HBM (A2/models/row_column_decoupled_attention.py:262-291);
// here it lives in TMEM only.  The kernels of this file keep the head's whole V slice resident: H, W <= 32 (512x512
// inputs); 32 < max(H, W) <= 64 (800x800: 50x50) dispatches to the streaming variants in rcda_tc64.cu.
#include "common.cuh"
#include "../../include/cdetr.h"

#if defined(CDETR_BWDV_SPIN) && CDETR_BWDV_SPIN
#define BWDV_WAIT mbar_wait
#else
#define BWDV_WAIT mbar_wait_sleep
#endif

namespace {

constexpr int HD = 32;    // head dim
constexpr int TQ = 128;   // queries per CTA = TMEM lanes
constexpr int HP = 32;    // padded key rows
constexpr int WP = 32;    // padded key columns (= MMA K, two k-steps of 16)
constexpr uint32_t V_PLANE_BYTES = HP * WP * HD * 2;  // 65536
constexpr uint32_t A_PLANE_BYTES = TQ * WP * 2;       // 8192

struct TcArgs {
  int B, L, H, W, E, nh;
  const float* qr;
  const float* qc;
  const float* kr;
  const float* kc;
  const uint8_t* mask_row;
  const uint8_t* mask_col;
  float* ar;  // [B,nh,W,L]
  float* ac;  // [B,nh,H,L]
  __nv_bfloat16* o_hi;
  __nv_bfloat16* o_lo;
  int64_t ld_o;
  uint32_t idesc;
  uint32_t idesc_s;   // logits MMA: [128 x 32] x [32 x 32]
};

// softmax over the first n entries of a logit row (entries >= n or masked get probability 0)
__device__ __forceinline__ void softmax_row32(float (&p)[32], int n, const uint8_t* __restrict__ mask) {
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    if (k >= n || (mask && mask[k])) p[k] = -INFINITY;
    mx = fmaxf(mx, p[k]);
  }
  float sum = 0.0f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    p[k] = expf(p[k] - mx);
    sum += p[k];
  }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int k = 0; k < 32; ++k) p[k] *= inv;
}
// one 32-value row -> split-bf16 K-major SWIZZLE_64B operand row r (64-byte rows): hi plane at dst, lo at dst + plane
__device__ __forceinline__ void write_operand_row32(uint8_t* dst, uint32_t plane_bytes, int r, const float (&x)[32]) {
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

__global__ void __launch_bounds__(320, 1)
rcda_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmV, const TcArgs a) {
  pdl_trigger();   // light successors (launch_light) may pre-launch; they wait for this grid to finish
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Vs = smem;                                   // [2][HP][WP][32] bf16, SW64
  uint8_t* As = Vs + 2 * V_PLANE_BYTES;                 // [2 tiles][2 planes][128][32] bf16, SW64 K-major
  // key tiles as split-bf16 K-major SW64 B operands of the logit MMAs: [2 sides][2 planes][32 keys][32 d] = 8 KB
  uint8_t* Kb = As + 4 * A_PLANE_BYTES;
  constexpr uint32_t K_PLANE_BYTES = 32 * HD * 2;   // 2048
  uint64_t* bars = reinterpret_cast<uint64_t*>(Kb + 4 * K_PLANE_BYTES);
  uint64_t* v_full = bars;
  uint64_t* a_ready = bars + 1;  // [2]
  uint64_t* t_full = bars + 3;   // [2]
  uint64_t* t_empty = bars + 5;  // [2]
  uint64_t* q_ready = bars + 7;  // [2]  q_r (phase 0) / q_c (phase 1) operand of the group written
  uint64_t* s_full = bars + 9;   // [2]  logits of the group in TMEM (phase 0: row side, phase 1: column side)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int q_cta = blockIdx.x * 2 * TQ;
  const int nblk = (a.H + 7) / 8;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmV);
    mbar_init(v_full, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&a_ready[g], 4);      // one elected lane per warp of the group
      mbar_init(&t_full[g], 1);
      mbar_init(&t_empty[g], 4);
      mbar_init(&q_ready[g], 4);
      mbar_init(&s_full[g], 1);
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

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(v_full, 2 * V_PLANE_BYTES);
      tma_load_5d(Vs, &tmV, v_full, head * HD, 0, 0, b, 0);                  // box {32 c, WP, HP, 1, 1}
      tma_load_5d(Vs + V_PLANE_BYTES, &tmV, v_full, head * HD, 0, 0, b, 1);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const int ngroups = (q_cta + TQ < a.L) ? 2 : 1;
      // logits: S = q K^T per side and group, [128 x 32] x [32 x 32] into the first 32 columns of the group's buffer
      for (int side = 0; side < 2; ++side) {
        for (int g = 0; g < ngroups; ++g) {
          mbar_wait_single(&q_ready[g], (uint32_t)side);
          tc_fence_after();
          const uint32_t a_base = smem_u32(As) + (uint32_t)g * 2u * A_PLANE_BYTES;
          const uint32_t k_base = smem_u32(Kb) + (uint32_t)side * 2u * K_PLANE_BYTES;
          const uint32_t d = tmem_base + (uint32_t)g * 256u;
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {
            const uint64_t a_hi = make_smem_desc(a_base + ks * 32, 16, 512, 4);
            const uint64_t a_lo = make_smem_desc(a_base + A_PLANE_BYTES + ks * 32, 16, 512, 4);
            const uint64_t b_hi = make_smem_desc(k_base + ks * 32, 16, 512, 4);
            const uint64_t b_lo = make_smem_desc(k_base + K_PLANE_BYTES + ks * 32, 16, 512, 4);
            umma_bf16_ss(d, a_hi, b_lo, a.idesc_s, ks > 0 ? 1u : 0u);
            umma_bf16_ss(d, a_lo, b_hi, a.idesc_s, 1);
            umma_bf16_ss(d, a_hi, b_hi, a.idesc_s, 1);
          }
          umma_commit(&s_full[g]);
        }
      }
      mbar_wait_single(v_full, 0);
      const uint32_t v_base = smem_u32(Vs);
      for (int p = 0; p < nblk; ++p) {
        for (int g = 0; g < ngroups; ++g) {
          if (p == 0) mbar_wait_single(&a_ready[g], 0);
          else mbar_wait_single(&t_empty[g], (uint32_t)(p - 1) & 1u);
          tc_fence_after();
          const uint32_t a_base = smem_u32(As) + (uint32_t)g * 2u * A_PLANE_BYTES;
          const uint32_t d = tmem_base + (uint32_t)g * 256u;
#pragma unroll
          for (int ks = 0; ks < WP / 16; ++ks) {
            // A_r: K-major SW64 (64-byte rows, 8-row groups 512 B apart, k-step +32 B)
            const uint64_t a_hi = make_smem_desc(a_base + ks * 32, 16, 512, 4);
            const uint64_t a_lo = make_smem_desc(a_base + A_PLANE_BYTES + ks * 32, 16, 512, 4);
            // V: MN-major SW64: MN chunk = one key row h (32 channels, 64 B); chunks WP*64 B apart (LBO);
            // 8 key columns w (the k index) per 512-byte group (SBO); one k-step = 16 w = +1024 B
            const uint32_t vb = v_base + (uint32_t)p * 8u * (WP * 64) + ks * 1024;
            const uint64_t b_hi = make_smem_desc(vb, WP * 64, 512, 4);
            const uint64_t b_lo = make_smem_desc(vb + V_PLANE_BYTES, WP * 64, 512, 4);
            umma_bf16_ss(d, a_hi, b_lo, a.idesc, ks > 0 ? 1u : 0u);
            umma_bf16_ss(d, a_lo, b_hi, a.idesc, 1);
            umma_bf16_ss(d, a_hi, b_hi, a.idesc, 1);
          }
          umma_commit(&t_full[g]);
        }
      }
    }
  } else {
    // ------------------------------ compute warps: one query per thread ------------------------------
    const int ct = threadIdx.x - 64;       // 0..255
    const int g = (warp - 2) >> 2;         // tile / TMEM buffer of this warp group
    const int quarter = warp & 3;          // TMEM lane quarter accessible to this warp
    const int r = quarter * 32 + lane;     // row inside the 128-query tile == TMEM lane
    const int q = q_cta + g * TQ + r;
    const bool ok = q < a.L;
    const bool active = q_cta + g * TQ < a.L;   // whole group idle when its tile is past the end
    {   // thread -> (side, key, 8-channel chunk): 2 x 32 x 4 = 256 chunks of the two key tiles
      const int side = ct >> 7, k = (ct >> 2) & 31, j = ct & 3;
      const int n = side == 0 ? a.W : a.H;
      const float* kp = (side == 0 ? a.kr : a.kc) + ((int64_t)b * n + k) * a.E + head * HD + j * 8;
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = 0.0f;
      if (k < n) {
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(kp)), t1 = __ldg(reinterpret_cast<const float4*>(kp) + 1);
        x[0] = t0.x; x[1] = t0.y; x[2] = t0.z; x[3] = t0.w; x[4] = t1.x; x[5] = t1.y; x[6] = t1.z; x[7] = t1.w;
      }
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_bf16_pair(x[2 * i], x[2 * i + 1], hw[i], lw[i]);
      uint8_t* kd = Kb + (size_t)side * 2 * K_PLANE_BYTES;
      const uint32_t off = (uint32_t)k * 64u + (uint32_t)((j ^ ((k >> 1) & 3)) << 4);
      sts128(smem_u32(kd + off), hw[0], hw[1], hw[2], hw[3]);
      sts128(smem_u32(kd + K_PLANE_BYTES + off), lw[0], lw[1], lw[2], lw[3]);
      fence_proxy_async();
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (active) {
      uint8_t* Ag = As + (size_t)g * 2 * A_PLANE_BYTES;
      const float scale = rsqrtf((float)HD);
      const int64_t bh = (int64_t)b * a.nh + head;
      float ac[32];
      const uint32_t taddr_s = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)g * 256u;
      {
        float ar[32], qv[HD];
        // ---- row side: q_r (scaled) -> A operand -> tensor core -> logits of this query back from TMEM
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
        write_operand_row32(Ag, A_PLANE_BYTES, r, qv);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_ready[g]);
        // column-side query while the row logits are computed
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
        mbar_wait(&s_full[g], 0);
        tc_fence_after();
        {
          uint32_t t[32];
          tmem_ld_32x32b_x32(taddr_s, t);
          tmem_ld_wait();
#pragma unroll
          for (int w = 0; w < 32; ++w) ar[w] = __uint_as_float(t[w]);
        }
        tc_fence_before();
        // the row-logit MMAs have consumed q_r: the operand buffer takes q_c
        write_operand_row32(Ag, A_PLANE_BYTES, r, qv);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_ready[g]);
        if (ok) {
          softmax_row32(ar, a.W, a.mask_row ? a.mask_row + (int64_t)b * a.W : nullptr);
#pragma unroll
          for (int w = 0; w < 32; ++w)
            if (w < a.W) a.ar[(bh * a.W + w) * a.L + q] = ar[w];
        } else {
#pragma unroll
          for (int w = 0; w < 32; ++w) ar[w] = 0.0f;
        }
        mbar_wait(&s_full[g], 1);
        tc_fence_after();
        {
          uint32_t t[32];
          tmem_ld_32x32b_x32(taddr_s, t);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 32; ++h) ac[h] = __uint_as_float(t[h]);
        }
        tc_fence_before();
        // the column-logit MMAs have consumed q_c: the operand buffer takes A_r for the main contraction
        write_operand_row32(Ag, A_PLANE_BYTES, r, ar);
        fence_proxy_async();   // make the generic-proxy smem writes visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[g]);
        if (ok) {
          softmax_row32(ac, a.H, a.mask_col ? a.mask_col + (int64_t)b * a.H : nullptr);
#pragma unroll
          for (int h = 0; h < 32; ++h)
            if (h < a.H) a.ac[(bh * a.H + h) * a.L + q] = ac[h];
        } else {
#pragma unroll
          for (int h = 0; h < 32; ++h) ac[h] = 0.0f;
        }
      }
      float acc[HD];
#pragma unroll
      for (int c = 0; c < HD; ++c) acc[c] = 0.0f;
      const uint32_t taddr_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)g * 256u;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (p < nblk) {
          mbar_wait(&t_full[g], (uint32_t)p & 1u);
          tc_fence_after();
#pragma unroll
          for (int hi = 0; hi < 8; ++hi) {
            uint32_t t[32];
            tmem_ld_32x32b_x32(taddr_row + (uint32_t)hi * 32u, t);
            tmem_ld_wait();
            const float coef = ac[p * 8 + hi];
            const float2 coef2 = make_float2(coef, coef);
#pragma unroll
            for (int c = 0; c < HD; c += 2) {   // packed FFMA2 (sm_100): two fp32 FMAs per issue slot, same rounding
              const float2 r2 = __ffma2_rn(coef2, make_float2(__uint_as_float(t[c]), __uint_as_float(t[c + 1])),
                                           make_float2(acc[c], acc[c + 1]));
              acc[c] = r2.x; acc[c + 1] = r2.y;
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[g]);
        }
      }
      if (ok) {
        const int64_t off = ((int64_t)b * a.L + q) * a.ld_o + head * HD;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            split_bf16_pair(acc[gg * 8 + 2 * j], acc[gg * 8 + 2 * j + 1], hw[j], lw[j]);
          }
          reinterpret_cast<uint4*>(a.o_hi + off)[gg] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          reinterpret_cast<uint4*>(a.o_lo + off)[gg] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
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

// ---------------------------------------------------------------------------------------------
// Backward (query side) on tensor cores:  G[q,(h,w)] = sum_c dO[q,c] V[h,w,c]  -> tcgen05.mma
//   [128 x 32] x [32 x (8*32)] per block of 8 key rows; epilogue per query thread:
//   dA_c[h] = sum_w A_r[w] G[h,w],  dA_r[w] += A_c[h] G[h,w];  then both softmax backwards, dS stored
//   (transposed) for the key-side kernel, and dq = s * dS K written as split-bf16.
// Same V staging (TMA, SWIZZLE_64B) as the forward; here V is the K-major B operand (rows = (h,w)).
struct TcBwdArgs {
  int B, L, H, W, E, nh;
  const float* kr;
  const float* kc;
  const float* ar;   // [B,nh,W,L]
  const float* ac;   // [B,nh,H,L]
  const float* d_o;  // [B,L,E]
  float* dsr;        // [B,nh,W,L]
  float* dsc;        // [B,nh,H,L]
  __nv_bfloat16 *dqr_hi, *dqr_lo, *dqc_hi, *dqc_lo;
  int64_t ld_g;
  uint32_t idesc;
  uint32_t idesc_q;   // dq MMA: [128 x 32 keys] x [32 keys x 32 d], B MN-major
};

__global__ void __launch_bounds__(320, 1)
rcda_bwd_q_tc_kernel(const __grid_constant__ CUtensorMap tmV, const TcBwdArgs a) {
  pdl_trigger();   // light successors (launch_light) may pre-launch; they wait for this grid to finish
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Vs = smem;                     // V hi/lo; after the last MMA its first 8 KB are reused for the K slices
  uint8_t* As = Vs + 2 * V_PLANE_BYTES;   // dO tiles [2 tiles][2 planes][128][32] bf16, SW64 K-major
  float* acs = reinterpret_cast<float*>(As + 4 * A_PLANE_BYTES);   // [2][32][128] A_c per query thread
  float* dacs = acs + 2 * 32 * TQ;                                  // [2][32][128] dA_c
  uint64_t* bars = reinterpret_cast<uint64_t*>(dacs + 2 * 32 * TQ);
  uint64_t* v_full = bars;
  uint64_t* a_ready = bars + 1;  // [2]
  uint64_t* t_full = bars + 3;   // [2]
  uint64_t* t_empty = bars + 5;  // [2]
  uint64_t* q_ready = bars + 7;  // [2]  dS_r (phase 0) / dS_c (phase 1) operand of the group written
  uint64_t* s_full = bars + 9;   // [2]  dq of the group in TMEM
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 11);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int q_cta = blockIdx.x * 2 * TQ;
  const int nblk = (a.H + 7) / 8;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmV);
    mbar_init(v_full, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&a_ready[g], 4);      // one elected lane per warp of the group
      mbar_init(&t_full[g], 1);
      mbar_init(&t_empty[g], 4);
      mbar_init(&q_ready[g], 4);
      mbar_init(&s_full[g], 1);
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

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(v_full, 2 * V_PLANE_BYTES);
      tma_load_5d(Vs, &tmV, v_full, head * HD, 0, 0, b, 0);
      tma_load_5d(Vs + V_PLANE_BYTES, &tmV, v_full, head * HD, 0, 0, b, 1);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const int ngroups = (q_cta + TQ < a.L) ? 2 : 1;
      mbar_wait_single(v_full, 0);
      const uint32_t v_base = smem_u32(Vs);
      for (int p = 0; p < nblk; ++p) {
        for (int g = 0; g < ngroups; ++g) {
          if (p == 0) mbar_wait_single(&a_ready[g], 0);
          else mbar_wait_single(&t_empty[g], (uint32_t)(p - 1) & 1u);
          tc_fence_after();
          const uint32_t a_base = smem_u32(As) + (uint32_t)g * 2u * A_PLANE_BYTES;
          const uint32_t d = tmem_base + (uint32_t)g * 256u;
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {
            const uint64_t a_hi = make_smem_desc(a_base + ks * 32, 16, 512, 4);
            const uint64_t a_lo = make_smem_desc(a_base + A_PLANE_BYTES + ks * 32, 16, 512, 4);
            // V as K-major B: row n = (h_local*32 + w) at 64 B pitch, 8-row groups 512 B apart, k-step (16 c) = +32 B
            const uint32_t vb = v_base + (uint32_t)p * 256u * 64u + ks * 32;
            const uint64_t b_hi = make_smem_desc(vb, 16, 512, 4);
            const uint64_t b_lo = make_smem_desc(vb + V_PLANE_BYTES, 16, 512, 4);
            umma_bf16_ss(d, a_hi, b_lo, a.idesc, ks > 0 ? 1u : 0u);
            umma_bf16_ss(d, a_lo, b_hi, a.idesc, 1);
            umma_bf16_ss(d, a_hi, b_hi, a.idesc, 1);
          }
          umma_commit(&t_full[g]);
        }
      }
      // dq = dS K per side and group: A = dS rows [128 x 32 keys] (K-major), B = key tile [32 keys x 32 d] read as
      // an MN-major operand (row = key = k index, 64-byte rows of 32 channels), D = first 32 columns of the buffer
      for (int side = 0; side < 2; ++side) {
        for (int g = 0; g < ngroups; ++g) {
          mbar_wait_single(&q_ready[g], (uint32_t)side);
          tc_fence_after();
          const uint32_t a_base = smem_u32(As) + (uint32_t)g * 2u * A_PLANE_BYTES;
          const uint32_t k_base = smem_u32(Vs) + (uint32_t)side * 4096u;
          const uint32_t d = tmem_base + (uint32_t)g * 256u;
#pragma unroll
          for (int ks = 0; ks < WP / 16; ++ks) {
            const uint64_t a_hi = make_smem_desc(a_base + ks * 32, 16, 512, 4);
            const uint64_t a_lo = make_smem_desc(a_base + A_PLANE_BYTES + ks * 32, 16, 512, 4);
            const uint64_t b_hi = make_smem_desc(k_base + ks * 1024, 2048, 512, 4);
            const uint64_t b_lo = make_smem_desc(k_base + 2048 + ks * 1024, 2048, 512, 4);
            umma_bf16_ss(d, a_hi, b_lo, a.idesc_q, ks > 0 ? 1u : 0u);
            umma_bf16_ss(d, a_lo, b_hi, a.idesc_q, 1);
            umma_bf16_ss(d, a_hi, b_hi, a.idesc_q, 1);
          }
          umma_commit(&s_full[g]);
        }
      }
    }
  } else {
    const int ct = threadIdx.x - 64;       // 0..255
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int q = q_cta + g * TQ + r;
    const bool ok = q < a.L;
    const bool active = q_cta + g * TQ < a.L;
    const int64_t bh = (int64_t)b * a.nh + head;
    // per-thread columns of the two [32][128] fp32 arrays, addressed as explicit shared memory (LDS / STS, not generic)
    const uint32_t acg = smem_u32(acs + g * 32 * TQ) + 4u * (uint32_t)r;
    const uint32_t dacg = smem_u32(dacs + g * 32 * TQ) + 4u * (uint32_t)r;
    float ar[32], dar[32];
    if (active) {
      uint8_t* Ag = As + (size_t)g * 2 * A_PLANE_BYTES;
      {  // dO row -> split-bf16 A operand (SWIZZLE_64B)
        float dov[HD];
        if (ok) {
          const float4* dp = reinterpret_cast<const float4*>(a.d_o + ((int64_t)b * a.L + q) * a.E + head * HD);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(dp + j);
            dov[4 * j] = t.x; dov[4 * j + 1] = t.y; dov[4 * j + 2] = t.z; dov[4 * j + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < HD; ++j) dov[j] = 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            split_bf16_pair(dov[8 * j + 2 * i], dov[8 * j + 2 * i + 1], hw[i], lw[i]);
          }
          const uint32_t off = (uint32_t)r * 64u + (uint32_t)((j ^ ((r >> 1) & 3)) << 4);
          sts128(smem_u32(Ag + off), hw[0], hw[1], hw[2], hw[3]);
          sts128(smem_u32(Ag + A_PLANE_BYTES + off), lw[0], lw[1], lw[2], lw[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[g]);
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        ar[k] = (ok && k < a.W) ? __ldg(a.ar + (bh * a.W + k) * a.L + q) : 0.0f;
        sts32(acg + 4u * TQ * k, (ok && k < a.H) ? __ldg(a.ac + (bh * a.H + k) * a.L + q) : 0.0f);
        sts32(dacg + 4u * TQ * k, 0.0f);
        dar[k] = 0.0f;
      }
      const uint32_t taddr_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)g * 256u;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (p < nblk) {
          mbar_wait(&t_full[g], (uint32_t)p & 1u);
          tc_fence_after();
#pragma unroll
          for (int hi = 0; hi < 8; ++hi) {
            uint32_t t[32];
            tmem_ld_32x32b_x32(taddr_row + (uint32_t)hi * 32u, t);
            tmem_ld_wait();
            const float cc = lds32(acg + 4u * TQ * (p * 8 + hi));
            float2 s2 = make_float2(0.0f, 0.0f);
            const float2 cc2 = make_float2(cc, cc);
#pragma unroll
            for (int w = 0; w < 32; w += 2) {   // packed FFMA2: (s0, s1) += A_r pair * G pair; dA_r pair += A_c * G pair
              const float2 g2 = make_float2(__uint_as_float(t[w]), __uint_as_float(t[w + 1]));
              s2 = __ffma2_rn(make_float2(ar[w], ar[w + 1]), g2, s2);
              const float2 d2 = __ffma2_rn(cc2, g2, make_float2(dar[w], dar[w + 1]));
              dar[w] = d2.x; dar[w + 1] = d2.y;
            }
            sts32(dacg + 4u * TQ * (p * 8 + hi), s2.x + s2.y);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[g]);
        }
      }
    }
    // every G MMA has completed once both groups are past their last t_full wait: the V region takes the two key
    // tiles as split-bf16 operands ([2 sides][2 planes][32 keys][32 d], SW64 rows of 64 bytes) and the dO operand
    // buffers take the dS rows; dq = scale * dS K then runs on the tensor core and comes back one query per thread
    asm volatile("bar.sync 1, 256;" ::: "memory");
    {
      const int side = ct >> 7, k = (ct >> 2) & 31, j = ct & 3;
      const int n = side == 0 ? a.W : a.H;
      const float* kp = (side == 0 ? a.kr : a.kc) + ((int64_t)b * n + k) * a.E + head * HD + j * 8;
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = 0.0f;
      if (k < n) {
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(kp)), t1 = __ldg(reinterpret_cast<const float4*>(kp) + 1);
        x[0] = t0.x; x[1] = t0.y; x[2] = t0.z; x[3] = t0.w; x[4] = t1.x; x[5] = t1.y; x[6] = t1.z; x[7] = t1.w;
      }
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_bf16_pair(x[2 * i], x[2 * i + 1], hw[i], lw[i]);
      uint8_t* kd = Vs + (size_t)side * 4096;
      const uint32_t off = (uint32_t)k * 64u + (uint32_t)((j ^ ((k >> 1) & 3)) << 4);
      sts128(smem_u32(kd + off), hw[0], hw[1], hw[2], hw[3]);
      sts128(smem_u32(kd + 2048 + off), lw[0], lw[1], lw[2], lw[3]);
      fence_proxy_async();
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (active) {
      uint8_t* Ag = As + (size_t)g * 2 * A_PLANE_BYTES;
      const float scale = rsqrtf((float)HD);
      const uint32_t taddr_s = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)g * 256u;
      const int64_t off = ((int64_t)b * a.L + q) * a.ld_g + head * HD;
      float ds[32];
      {   // row side: dS_r = A_r o (dA_r - <A_r, dA_r>)
        float dot = 0.0f;
#pragma unroll
        for (int w = 0; w < 32; ++w) dot += ar[w] * dar[w];
#pragma unroll
        for (int w = 0; w < 32; ++w) {
          ds[w] = ar[w] * (dar[w] - dot);
          if (ok && w < a.W) a.dsr[(bh * a.W + w) * a.L + q] = ds[w];
        }
      }
      write_operand_row32(Ag, A_PLANE_BYTES, r, ds);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_ready[g]);
      {   // column side while the tensor core works on the row side
        float dot = 0.0f;
#pragma unroll
        for (int h = 0; h < 32; ++h) dot += lds32(acg + 4u * TQ * h) * lds32(dacg + 4u * TQ * h);
#pragma unroll
        for (int h = 0; h < 32; ++h) {
          ds[h] = lds32(acg + 4u * TQ * h) * (lds32(dacg + 4u * TQ * h) - dot);
          if (ok && h < a.H) a.dsc[(bh * a.H + h) * a.L + q] = ds[h];
        }
      }
      mbar_wait(&s_full[g], 0);
      tc_fence_after();
      {
        uint32_t t[32];
        tmem_ld_32x32b_x32(taddr_s, t);
        tmem_ld_wait();
        tc_fence_before();
        // the row-side MMAs have consumed dS_r: the operand buffer takes dS_c
        write_operand_row32(Ag, A_PLANE_BYTES, r, ds);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_ready[g]);
        if (ok) {
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              split_bf16_pair(__uint_as_float(t[gg * 8 + 2 * j]) * scale, __uint_as_float(t[gg * 8 + 2 * j + 1]) * scale, hw[j], lw[j]);
            reinterpret_cast<uint4*>(a.dqr_hi + off)[gg] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            reinterpret_cast<uint4*>(a.dqr_lo + off)[gg] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
      }
      mbar_wait(&s_full[g], 1);
      tc_fence_after();
      {
        uint32_t t[32];
        tmem_ld_32x32b_x32(taddr_s, t);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              split_bf16_pair(__uint_as_float(t[gg * 8 + 2 * j]) * scale, __uint_as_float(t[gg * 8 + 2 * j + 1]) * scale, hw[j], lw[j]);
            reinterpret_cast<uint4*>(a.dqc_hi + off)[gg] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            reinterpret_cast<uint4*>(a.dqc_lo + off)[gg] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
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

// ---------------------------------------------------------------------------------------------
// Backward (value side) on tensor cores:  dV[(h,w), c] = sum_q A_c[q,h] A_r[q,w] dO[q,c]
// One CTA per (sample, head, slice of key positions).  Per block of 64 queries: dO tile by TMA (split planes,
// SWIZZLE_64B, MN-major B operand), the attention maps of those queries staged in shared memory (cp.async, one block
// ahead), and for each 128-row tile of key positions the compute warps build P[(h,w), q] = A_c A_r as a split-bf16
// A operand IN TENSOR MEMORY (tcgen05.st, one key position per TMEM lane = per thread, two bf16 per column), which one
// thread multiplies (tcgen05.mma with the A operand read from TMEM) into the tile's 32 accumulator columns.  The
// [L, H*W] outer-product matrix never touches shared memory: round 1 wrote every P tile to shared memory and had the
// tensor core read it back three times (hi twice, lo once), and ncu showed that kernel bound by the shared-memory
// pipe (l1tex data-pipe wavefronts 74 % of peak while active, tensor pipe 9 %); per tile that was 32 KB of stores +
// 48 KB of operand reads beside 64 KB of map loads, now only the map loads (and 12 KB of dO operand reads) remain.
// TMEM plan: tiles_per_cta * 32 accumulator columns + p_bufs * 64 operand columns (hi plane | lo plane, 32 columns
// = 64 queries each); 256 columns per CTA keep two CTAs per SM.  H, W <= 64.
constexpr int VK = 64;                                   // queries per k block
constexpr uint32_t DO_PLANE_BYTES = VK * HD * 2;         // 4 KB
constexpr int MAP_LD = 68;                               // map row = 64 queries + 4: conflict-free float4 reads of 8 rows

struct TcBwdVArgs {
  int B, L, H, W, E, nh;
  const float* ar;  // [B,nh,W,L]
  const float* ac;  // [B,nh,H,L]
  __nv_bfloat16 *dv_hi, *dv_lo;
  int64_t ld_g;
  uint32_t idesc;
  int tiles_per_cta;   // key-position tiles (of 128) per CTA: blockIdx.z selects the slice of positions
  int kp;              // rows of the staged attention maps: max(H, W) rounded up to 8 (<= 64)
  int map_bufs;        // 2: the next query block's maps are prefetched (cp.async) while this block's tiles are built
  int d_bufs;          // dO stages
  int p_bufs;          // P operand buffers in TMEM (2 when the accumulators leave 128 columns free)
  int tmem_cols;       // allocation: 256 (two CTAs per SM) or 512
  int vec4;            // L % 4 == 0 and 16-byte aligned maps: stage them with 16-byte cp.async
};


__global__ void __launch_bounds__(320, 2)
rcda_bwd_v_tc_kernel(const __grid_constant__ CUtensorMap tmD, const TcBwdVArgs a) {
  pdl_trigger();   // light successors (launch_light) may pre-launch; they wait for this grid to finish
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Ds = smem;                                    // [d_bufs][2 planes][64 q][32 c] bf16, SW64
  float* maps = reinterpret_cast<float*>(Ds + a.d_bufs * 2 * DO_PLANE_BYTES);   // [map_bufs][A_r | A_c][kp][MAP_LD]
  const int map_stride = 2 * a.kp * MAP_LD;
  uint64_t* bars = reinterpret_cast<uint64_t*>(maps + a.map_bufs * map_stride);
  uint64_t* d_full = bars;       // [2]
  uint64_t* d_empty = bars + 2;  // [2]
  uint64_t* p_full = bars + 4;   // [2]
  uint64_t* p_empty = bars + 6;  // [2]
  uint64_t* acc_full = bars + 8;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 9);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x, b = blockIdx.y;
  const int HW = a.H * a.W;
  // positions are split across gridDim.z CTAs (each rebuilds only its own rows of P; dO blocks are re-read from L2)
  const int mt_begin = blockIdx.z * a.tiles_per_cta;
  const int mt_end = min((HW + 127) / 128, mt_begin + a.tiles_per_cta);
  const int nkb = (a.L + VK - 1) / VK;
  const uint32_t p_col0 = (uint32_t)a.tiles_per_cta * 32u;   // first operand column behind the accumulators
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmD);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 1);
      mbar_init(&d_empty[i], 1);
      mbar_init(&p_full[i], 8);        // one elected lane per compute warp
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, (uint32_t)a.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.d_bufs;
        if (kb >= a.d_bufs) BWDV_WAIT(&d_empty[s], (uint32_t)(kb / a.d_bufs - 1) & 1u);
        mbar_arrive_expect_tx(&d_full[s], 2 * DO_PLANE_BYTES);
        tma_load_3d(Ds + s * 2 * DO_PLANE_BYTES, &tmD, &d_full[s], head * HD, b * a.L + kb * VK, 0);
        tma_load_3d(Ds + s * 2 * DO_PLANE_BYTES + DO_PLANE_BYTES, &tmD, &d_full[s], head * HD, b * a.L + kb * VK, 1);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int u = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.d_bufs;
        BWDV_WAIT(&d_full[s], (uint32_t)(kb / a.d_bufs) & 1u);
        const uint32_t d_base = smem_u32(Ds) + (uint32_t)s * 2u * DO_PLANE_BYTES;
        for (int mt = mt_begin; mt < mt_end; ++mt, ++u) {
          const int pb = u % a.p_bufs;
          BWDV_WAIT(&p_full[pb], (uint32_t)(u / a.p_bufs) & 1u);
          tc_fence_after();
          const uint32_t p_base = tmem_base + p_col0 + (uint32_t)pb * 64u;   // hi plane; lo plane 32 columns further
          const uint32_t d = tmem_base + (uint32_t)(mt - mt_begin) * 32u;
#pragma unroll
          for (int ks = 0; ks < VK / 16; ++ks) {
            const uint32_t a_hi = p_base + (uint32_t)ks * 8u;                // 16 queries = 8 packed columns
            const uint32_t a_lo = a_hi + 32u;
            const uint64_t b_hi = make_smem_desc(d_base + ks * 1024, 4096, 512, 4);              // MN-major SW64
            const uint64_t b_lo = make_smem_desc(d_base + DO_PLANE_BYTES + ks * 1024, 4096, 512, 4);
            umma_bf16_ts(d, a_hi, b_lo, a.idesc, (kb > 0 || ks > 0) ? 1u : 0u);
            umma_bf16_ts(d, a_lo, b_hi, a.idesc, 1);
            umma_bf16_ts(d, a_hi, b_hi, a.idesc, 1);
          }
          umma_commit(&p_empty[pb]);
        }
        umma_commit(&d_empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    const int ct = threadIdx.x - 64;        // 0..255
    const int quarter = warp & 3;           // TMEM lane quarter this warp may access
    const int ml = quarter * 32 + lane;     // row of the P tile (= TMEM lane) built by this thread
    const int qh = (warp - 2) >> 2;         // which half (32 queries) of the k block
    const int64_t bh = (int64_t)b * a.nh + head;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int u = 0;
    // maps of one block of 64 queries -> shared memory, asynchronously (cp.async, 4 bytes each: rows of the transposed
    // maps are only 4-byte aligned for odd L); out-of-range entries are zeros
    auto stage_maps = [&](int kb2, float* dst) {
      const int q0 = kb2 * VK;
      if (a.vec4) {     // L % 4 == 0: 16-byte chunks (4 queries), each entirely inside or outside the sequence
        const int c4 = (ct & 15) * 4, q = q0 + c4;
        const uint32_t nb = q < a.L ? 16u : 0u;
        const float* sr = a.ar + bh * a.W * a.L + min(q, a.L - 4);
        const float* sc = a.ac + bh * a.H * a.L + min(q, a.L - 4);
        float* d = dst + c4;
        for (int k = ct >> 4; k < a.kp; k += 16) {
          cp_async_16_zfill(d + k * MAP_LD, sr + (int64_t)min(k, a.W - 1) * a.L, k < a.W ? nb : 0u);
          cp_async_16_zfill(d + (a.kp + k) * MAP_LD, sc + (int64_t)min(k, a.H - 1) * a.L, k < a.H ? nb : 0u);
        }
        return;
      }
      for (int i = ct; i < a.kp * VK; i += 256) {
        const int k = i / VK, qq = i % VK;
        const bool qok = q0 + qq < a.L;
        const int mo = k * MAP_LD + qq;
        if (qok && k < a.W) cp_async_4(dst + mo, a.ar + (bh * a.W + k) * a.L + q0 + qq); else dst[mo] = 0.0f;
        if (qok && k < a.H) cp_async_4(dst + a.kp * MAP_LD + mo, a.ac + (bh * a.H + k) * a.L + q0 + qq);
        else dst[a.kp * MAP_LD + mo] = 0.0f;
      }
    };
    if (a.map_bufs == 2) {
      stage_maps(0, maps);
      cp_async_wait_all();
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    for (int kb = 0; kb < nkb; ++kb) {
      float* ars = maps + ((a.map_bufs == 2) ? (kb & 1) * map_stride : 0);
      float* acs = ars + a.kp * MAP_LD;
      if (a.map_bufs == 1) {
        asm volatile("bar.sync 1, 256;" ::: "memory");     // previous block's tiles are all built
        stage_maps(kb, maps);
        cp_async_wait_all();
        asm volatile("bar.sync 1, 256;" ::: "memory");
      } else if (kb + 1 < nkb) {
        stage_maps(kb + 1, maps + ((kb + 1) & 1) * map_stride);   // in flight while this block's tiles are built
      }
      for (int mt = mt_begin; mt < mt_end; ++mt, ++u) {
        const int pb = u % a.p_bufs;
        if (u >= a.p_bufs) {
          BWDV_WAIT(&p_empty[pb], (uint32_t)(u / a.p_bufs - 1) & 1u);   // the MMAs that read this buffer are done
          tc_fence_after();
        }
        // rows past H*W only feed accumulator rows that are never stored: any in-range map row will do
        const int m = mt * 128 + ml;
        const int h = min(m / a.W, a.kp - 1), w = m % a.W;
        const uint32_t cr = smem_u32(acs + h * MAP_LD + qh * 32);     // explicit LDS: see lds128()
        const uint32_t rr = smem_u32(ars + w * MAP_LD + qh * 32);
        const uint32_t pt = lane_addr + p_col0 + (uint32_t)pb * 64u + (uint32_t)qh * 16u;
#pragma unroll
        for (int j = 0; j < 2; ++j) {        // 2 groups of 16 queries = 8 packed columns per plane
          uint32_t hw[8], lw[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 c = lds128(cr + 64 * j + 16 * i);
            const float4 r = lds128(rr + 64 * j + 16 * i);
            const float2 p01 = __fmul2_rn(make_float2(c.x, c.y), make_float2(r.x, r.y));
            const float2 p23 = __fmul2_rn(make_float2(c.z, c.w), make_float2(r.z, r.w));
            split_bf16_pair(p01.x, p01.y, hw[2 * i], lw[2 * i]);
            split_bf16_pair(p23.x, p23.y, hw[2 * i + 1], lw[2 * i + 1]);
          }
          tmem_st_32x32b_x8(pt + 8u * j, hw);
          tmem_st_32x32b_x8(pt + 32u + 8u * j, lw);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[pb]);
      }
      if (a.map_bufs == 2) {     // next block's maps have landed; everyone is done reading this block's
        cp_async_wait_all();
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
    // ---- epilogue: accumulators -> dV (split)
    BWDV_WAIT(acc_full, 0);
    tc_fence_after();
    const int grp = (warp - 2) >> 2;
    for (int mt = mt_begin + grp; mt < mt_end; mt += 2) {
      uint32_t t[32];
      tmem_ld_32x32b_x32(lane_addr + (uint32_t)(mt - mt_begin) * 32u, t);
      tmem_ld_wait();
      const int m = mt * 128 + quarter * 32 + lane;
      if (m < HW) {
        const int64_t off = ((int64_t)b * HW + m) * a.ld_g + head * HD;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            split_bf16_pair(__uint_as_float(t[gg * 8 + 2 * j]), __uint_as_float(t[gg * 8 + 2 * j + 1]), hw[j], lw[j]);
          }
          reinterpret_cast<uint4*>(a.dv_hi + off)[gg] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          reinterpret_cast<uint4*>(a.dv_lo + off)[gg] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
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

}  // namespace

// rcda_tc64.cu: the same kernels for 32 < max(H, W) <= 64 (V streamed through a TMA ring instead of resident)
int rcda_fwd_tc64_launch(int B, int L, int H, int W, int E, int nh, const float* qr, const float* qc, const float* kr,
                         const float* kc, cdetr_split_t v, const uint8_t* mask_row, const uint8_t* mask_col, float* ar,
                         float* ac, cdetr_split_t o, cudaStream_t s);
int rcda_bwd_q_tc64_launch(int B, int L, int H, int W, int E, int nh, const float* kr, const float* kc, cdetr_split_t v,
                           const float* ar, const float* ac, const float* d_o, float* dsr, float* dsc, cdetr_split_t dqr,
                           cdetr_split_t dqc, cudaStream_t s);

// v: split [B*H*W, E] (the value projection); everything else as cdetr_rcda_fwd.  Requires H, W <= 64.
extern "C" int cdetr_rcda_fwd_tc(int B, int L, int H, int W, int E, int nh, const float* qr, const float* qc,
                                 const float* kr, const float* kc, cdetr_split_t v, const uint8_t* mask_row,
                                 const uint8_t* mask_col, float* ar, float* ac, cdetr_split_t o,
                                 cdetr_stream_t s) {
  CDETR_CHECK_ARG(E == nh * HD, "rcda_fwd_tc: head dim must be 32");
  CDETR_CHECK_ARG(H >= 1 && W >= 1 && H <= 64 && W <= 64, "rcda_fwd_tc: H, W must be <= 64 (got %d x %d)", H, W);
  CDETR_CHECK_ARG(qr && qc && kr && kc && v.base && ar && ac && o.base, "rcda_fwd_tc: null pointer");
  CDETR_CHECK_ARG(v.ld % 8 == 0 && v.plane % 8 == 0 && (reinterpret_cast<uintptr_t>(v.base) & 15) == 0,
                  "rcda_fwd_tc: V must be 16-byte aligned with ld/plane multiples of 8");
  if (H > HP || W > WP || (cdetr_tuning().rcda_stream & 1))
    return rcda_fwd_tc64_launch(B, L, H, W, E, nh, qr, qc, kr, kc, v, mask_row, mask_col, ar, ac, o,
                                reinterpret_cast<cudaStream_t>(s));
  EncodeTiledFn fn = encode_fn();
  if (!fn) { cdetr_set_error("cuTensorMapEncodeTiled entry point unavailable"); return CDETR_ERR_CUDA; }
  CUtensorMap tm;
  cuuint64_t gdim[5] = {(cuuint64_t)E, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 2};
  cuuint64_t gstr[4] = {(cuuint64_t)v.ld * 2, (cuuint64_t)W * v.ld * 2, (cuuint64_t)H * W * v.ld * 2,
                        (cuuint64_t)v.plane * 2};
  cuuint32_t box[5] = {HD, WP, HP, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.base, gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { cdetr_set_error("rcda_fwd_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return CDETR_ERR_CUDA; }
  TcArgs a = {};
  a.B = B; a.L = L; a.H = H; a.W = W; a.E = E; a.nh = nh;
  a.qr = qr; a.qc = qc; a.kr = kr; a.kc = kc; a.mask_row = mask_row; a.mask_col = mask_col; a.ar = ar; a.ac = ac;
  a.o_hi = reinterpret_cast<__nv_bfloat16*>(o.base); a.o_lo = a.o_hi + o.plane; a.ld_o = o.ld;
  a.idesc = make_idesc_bf16_f32(TQ, 256, 0, 1);
  a.idesc_s = make_idesc_bf16_f32(TQ, 32, 0, 0);
  const size_t smem = 2 * V_PLANE_BYTES + 4 * A_PLANE_BYTES + 2 * 32 * HD * sizeof(float) + 128 + 1024;
  static DevAttrCache cfg = {};
  CDETR_CHECK_CUDA(cdetr_ensure_smem(rcda_fwd_tc_kernel, (int)smem, &cfg));
  rcda_fwd_tc_kernel<<<dim3(cdiv(L, 2 * TQ), nh, B), 320, smem, reinterpret_cast<cudaStream_t>(s)>>>(tm, a);
  CDETR_CHECK_LAUNCH();
  return 0;
}

namespace {
int make_v_map(CUtensorMap* tm, const cdetr_split_t& v, int B, int H, int W, int E) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { cdetr_set_error("cuTensorMapEncodeTiled entry point unavailable"); return CDETR_ERR_CUDA; }
  cuuint64_t gdim[5] = {(cuuint64_t)E, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 2};
  cuuint64_t gstr[4] = {(cuuint64_t)v.ld * 2, (cuuint64_t)W * v.ld * 2, (cuuint64_t)H * W * v.ld * 2,
                        (cuuint64_t)v.plane * 2};
  cuuint32_t box[5] = {HD, WP, HP, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.base, gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { cdetr_set_error("rcda tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return CDETR_ERR_CUDA; }
  return 0;
}
}  // namespace

// Query-side backward on tensor cores: writes dsr/dsc (for the key-side kernel) and dqr/dqc (split).
extern "C" int cdetr_rcda_bwd_q_tc(int B, int L, int H, int W, int E, int nh, const float* kr, const float* kc,
                                   cdetr_split_t v, const float* ar, const float* ac, const float* d_o, float* dsr,
                                   float* dsc, cdetr_split_t dqr, cdetr_split_t dqc, cdetr_stream_t s) {
  CDETR_CHECK_ARG(E == nh * HD, "rcda_bwd_q_tc: head dim must be 32");
  CDETR_CHECK_ARG(H >= 1 && W >= 1 && H <= 64 && W <= 64, "rcda_bwd_q_tc: H, W must be <= 64");
  CDETR_CHECK_ARG(kr && kc && v.base && ar && ac && d_o && dsr && dsc && dqr.base && dqc.base, "rcda_bwd_q_tc: null pointer");
  CDETR_CHECK_ARG(dqr.ld == dqc.ld, "rcda_bwd_q_tc: gradient tensors must share ld");
  if (H > HP || W > WP || (cdetr_tuning().rcda_stream & 2))
    return rcda_bwd_q_tc64_launch(B, L, H, W, E, nh, kr, kc, v, ar, ac, d_o, dsr, dsc, dqr, dqc,
                                  reinterpret_cast<cudaStream_t>(s));
  CUtensorMap tm;
  int rc = make_v_map(&tm, v, B, H, W, E);
  if (rc) return rc;
  TcBwdArgs a = {};
  a.B = B; a.L = L; a.H = H; a.W = W; a.E = E; a.nh = nh;
  a.kr = kr; a.kc = kc; a.ar = ar; a.ac = ac; a.d_o = d_o; a.dsr = dsr; a.dsc = dsc;
  a.dqr_hi = reinterpret_cast<__nv_bfloat16*>(dqr.base); a.dqr_lo = a.dqr_hi + dqr.plane;
  a.dqc_hi = reinterpret_cast<__nv_bfloat16*>(dqc.base); a.dqc_lo = a.dqc_hi + dqc.plane;
  a.ld_g = dqr.ld;
  a.idesc = make_idesc_bf16_f32(TQ, 256, 0, 0);
  a.idesc_q = make_idesc_bf16_f32(TQ, 32, 0, 1);
  const size_t smem = 2 * V_PLANE_BYTES + 4 * A_PLANE_BYTES + 4 * 32 * TQ * sizeof(float) + 128 + 1024;
  static DevAttrCache cfg = {};
  CDETR_CHECK_CUDA(cdetr_ensure_smem(rcda_bwd_q_tc_kernel, (int)smem, &cfg));
  rcda_bwd_q_tc_kernel<<<dim3(cdiv(L, 2 * TQ), nh, B), 320, smem, reinterpret_cast<cudaStream_t>(s)>>>(tm, a);
  CDETR_CHECK_LAUNCH();
  return 0;
}

// Value-side backward on tensor cores.  d_o: split [B*L, E] (gradient of the attention output before out_proj).
extern "C" int cdetr_rcda_bwd_v_tc(int B, int L, int H, int W, int E, int nh, const float* ar, const float* ac,
                                   cdetr_split_t d_o, cdetr_split_t dv, cdetr_stream_t s) {
  CDETR_CHECK_ARG(E == nh * HD, "rcda_bwd_v_tc: head dim must be 32");
  CDETR_CHECK_ARG(H >= 1 && W >= 1 && H <= 64 && W <= 64, "rcda_bwd_v_tc: H, W must be <= 64");
  CDETR_CHECK_ARG(ar && ac && d_o.base && dv.base, "rcda_bwd_v_tc: null pointer");
  CDETR_CHECK_ARG(d_o.ld % 8 == 0 && d_o.plane % 8 == 0 && (reinterpret_cast<uintptr_t>(d_o.base) & 15) == 0,
                  "rcda_bwd_v_tc: dO must be 16-byte aligned with ld/plane multiples of 8");
  EncodeTiledFn fn = encode_fn();
  if (!fn) { cdetr_set_error("cuTensorMapEncodeTiled entry point unavailable"); return CDETR_ERR_CUDA; }
  CUtensorMap tm;
  cuuint64_t gdim[3] = {(cuuint64_t)E, (cuuint64_t)B * L, 2};
  cuuint64_t gstr[2] = {(cuuint64_t)d_o.ld * 2, (cuuint64_t)d_o.plane * 2};
  cuuint32_t box[3] = {HD, VK, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d_o.base, gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { cdetr_set_error("rcda_bwd_v_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return CDETR_ERR_CUDA; }
  TcBwdVArgs a = {};
  a.B = B; a.L = L; a.H = H; a.W = W; a.E = E; a.nh = nh; a.ar = ar; a.ac = ac;
  a.dv_hi = reinterpret_cast<__nv_bfloat16*>(dv.base); a.dv_lo = a.dv_hi + dv.plane; a.ld_g = dv.ld;
  a.idesc = make_idesc_bf16_f32(128, HD, 0, 1);
  a.kp = ((H > W ? H : W) + 7) / 8 * 8;
  a.vec4 = (L % 4 == 0 && L >= 4 && ((reinterpret_cast<uintptr_t>(ar) | reinterpret_cast<uintptr_t>(ac)) & 15) == 0) ? 1 : 0;
  // split the key positions over enough CTAs to fill the machine at two CTAs per SM (one wave).  TMEM: 32 accumulator
  // columns per tile + 64 per P operand buffer; 256 columns per CTA keep two CTAs per SM (<= 4 tiles with two operand
  // buffers, <= 6 with one), else one CTA takes all 512 (<= 12 tiles)
  int num_sms = 0;
  CDETR_CHECK_CUDA(cdetr_num_sms(&num_sms));
  const int ntile = (H * W + 127) / 128;
  int nsplit = 1;
  while (nsplit * 2 <= ntile && nh * B * nsplit * 2 <= 2 * num_sms) nsplit *= 2;
  while ((ntile + nsplit - 1) / nsplit > 12) ++nsplit;
  a.tiles_per_cta = (ntile + nsplit - 1) / nsplit;
  nsplit = (ntile + a.tiles_per_cta - 1) / a.tiles_per_cta;
  a.p_bufs = (a.tiles_per_cta <= 4 || a.tiles_per_cta > 6) ? 2 : 1;
  a.tmem_cols = a.tiles_per_cta * 32 + a.p_bufs * 64 <= 256 ? 256 : 512;
  // shared memory: dO stages + the staged maps (double-buffered: the next query block's maps are prefetched)
  const size_t map_bytes = 2 * (size_t)a.kp * MAP_LD * sizeof(float);
  a.map_bufs = 2;
  a.d_bufs = 2;
  const size_t smem = (size_t)a.d_bufs * 2 * DO_PLANE_BYTES + a.map_bufs * map_bytes + 128 + 1024;
  static DevAttrCache cfg = {};
  CDETR_CHECK_CUDA(cdetr_ensure_smem(rcda_bwd_v_tc_kernel, 4 * DO_PLANE_BYTES + 2 * 2 * 64 * MAP_LD * 4 + 128 + 1024, &cfg));
  rcda_bwd_v_tc_kernel<<<dim3(nh, B, nsplit), 320, smem, reinterpret_cast<cudaStream_t>(s)>>>(tm, a);
  CDETR_CHECK_LAUNCH();
  return 0;
}
