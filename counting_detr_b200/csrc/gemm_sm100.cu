// tcgen05 / TMA / TMEM GEMM for split-bf16 operands (see common.cuh for the number format).
//
//   mode 0 (TN): D[M,N] = A[M,K] * B[N,K]^T    both operands K-major   (forward / dgrad)
//   mode 1 (NT): D[M,N] = A[K,M]^T * B[K,N]    both operands MN-major  (wgrad)
//
// Persistent CTAs (<= 2 per SM) walk a static round-robin list of 128 x BN output tiles (x K-splits):
//   warp 0     TMA producer  : cp.async.bulk.tensor (SWIZZLE_128B) of the hi+lo planes of A and B
//   warp 1     MMA issuer    : per 64-wide k block, 4 k-steps x 3 tcgen05.mma (hi*lo, lo*hi, hi*hi)
//   warps 2-9  epilogue      : tcgen05.ld of the fp32 accumulator (one TMEM lane = one row per thread)
//                              -> row-scale / bias / residual / ReLU / ReLU-mask -> fp32 and/or split-bf16
// smem stages are recycled through full/empty mbarriers across tiles; the accumulator is double buffered in
// TMEM (tmem_full/tmem_empty barriers) so the epilogue of tile i overlaps the loads and MMAs of tile i+1.
//
// Replaces the ATen GEMM call sites listed in include/cdetr.h (cdetr_gemm).
#include "common.cuh"
#include "../../include/cdetr.h"
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int NUM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (2 per TMEM lane quarter)
constexpr int MAX_STAGES = 6;

struct EpilogueArgs {
  const float* row_scale;
  const float* bias;
  const __nv_bfloat16* add_hi;
  const __nv_bfloat16* add_lo;
  int64_t ld_add;
  const float* add_f32;
  int64_t ld_add_f32;
  const __nv_bfloat16* mask_hi;
  int64_t ld_mask;
  int relu;
  int atomic;
  float* out_f32;
  int64_t ld_out_f32;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  int64_t ld_out_split;
};

struct KernelArgs {
  int M, N, K;
  int block_n;
  int stages;
  int kb_per_split;  // k blocks (of 64) handled by one K-split
  int num_kb;
  int tiles_m, tiles_n, total_tiles;  // tile index = (split * tiles_m + mt) * tiles_n + nt
  int tile_m;                         // output rows per tile: 128, or 256 for a CTA pair (cta_group::2), or tile_rows
  int tile_rows;                      // valid rows of a tile's 128 TMEM lanes (< 128: implicit conv on maps whose width does
                                      // not divide 128 -- the tile is th image rows x tw pixels = tile_rows linear rows)
  int conv_wblk;                      // mode-1 conv, widths not dividing 64: k blocks (64 pixels, zero-filled past W) per image row
  uint32_t idesc;
  uint32_t pass_mask;  // which of the three split-bf16 products are issued: 1 hi*hi, 2 hi_a*lo_b, 4 lo_a*hi_b (7 = all)
  uint32_t tx_a, tx_b; // bytes one k-block's TMA loads deliver per operand (a lo plane no product reads is not loaded)
  uint32_t tmem_cols;  // columns of ONE accumulator buffer (two are allocated)
  // implicit 3x3 convolution (conv != 0): the conv operand (A in mode 0, B in mode 1) is an NHWC activation read
  // through a 5-D map {C, W, H, B, plane} with per-tap shifted windows; TMA zero-fills the out-of-image part
  int conv, cH, cW, cC, cdil, csign;
  // resident-B schedule (mode 0, K <= 256): the whole [BN x K] weight slab of an n-tile stays in shared memory while
  // the CTA walks a contiguous run of m-tiles, so only A streams through the TMA ring (half the L2 -> SM bytes)
  int resident_b, tiles_per_cta;
  uint32_t staging_off;  // byte offset of the epilogue staging area (after the operand stages)
  uint32_t staging_bytes;
  // TMA epilogue: residual / mask tiles arrive by TMA loads, outputs leave by TMA stores (or TMA reduce-add), one
  // [32 rows x 32 columns] chunk per epilogue warp at a time through per-warp swizzled staging buffers
  int epi_tma;
  int epi_nb;                // staging buffers per epilogue warp (2: inputs of chunk i+1 prefetched during chunk i)
  uint32_t epi_buf_bytes;    // bytes of one staging buffer: [io 4096][mask 2048]?[second output 4096]?
  uint32_t epi_off_mask, epi_off_out2;
  int epi_add_kind;          // 0 none, 1 split residual, 2 fp32 residual
  int epi_debug;             // diagnostics (CDETR_GEMM_EPI_DEBUG): 1 no TMA stores, 2 no staging either, 3 TMEM read only
  long long* dbg;            // debug timeline of CTA 0 (cdetr_gemm_debug_timeline), normally NULL
  EpilogueArgs ep;
};

__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define DBG_T(i)                                                       \
  do {                                                                 \
    if (args.dbg != nullptr && blockIdx.x == 0) args.dbg[i] = gtimer(); \
  } while (0)

struct TileCoord {
  int m0, n0, kb_begin, kb_end;
};
__device__ __forceinline__ TileCoord decode_tile(const KernelArgs& a, int rank, int tile) {
  TileCoord t;
  if (a.resident_b) {  // n-major: consecutive tiles of a CTA share the n-tile
    t.n0 = (tile / a.tiles_m) * a.block_n;
    t.m0 = (tile % a.tiles_m) * BM;
    t.kb_begin = 0;
    t.kb_end = a.num_kb;
  } else {
    t.n0 = (tile % a.tiles_n) * a.block_n;
    t.m0 = ((tile / a.tiles_n) % a.tiles_m) * a.tile_m + rank * BM;   // rank: CTA of the pair (0 otherwise)
    t.kb_begin = (tile / (a.tiles_n * a.tiles_m)) * a.kb_per_split;
    t.kb_end = min(a.num_kb, t.kb_begin + a.kb_per_split);
  }
  return t;
}

// MASKED = false: the product path, three back-to-back MMAs per k-step in straight-line code (no data-dependent branch
// in the single-thread issue loop); MASKED = true: precision-policy variants that drop cross terms (args.pass_mask).
template <bool NT, bool PAIR, bool MASKED>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_split_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOutF, const __grid_constant__ CUtensorMap tmOutS,
                  const __grid_constant__ CUtensorMap tmAdd, const __grid_constant__ CUtensorMap tmMask,
                  const KernelArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment of every tile base.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  const int BN = args.block_n;
  const uint32_t a_bytes = 2u * BM * 128u;               // hi+lo planes, 128 B per row/k-row
  // CTA pair: this CTA stages its own 128 rows of A and BN/2 rows of B; the 256 x BN MMA reads both CTAs' halves
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int BNL = PAIR ? BN / 2 : BN;                    // B rows staged by this CTA
  const uint32_t b_bytes = NT ? (uint32_t)((BN + 63) / 64) * 16384u : 2u * (uint32_t)BNL * 128u;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  constexpr int STG_LD = 33;                             // epilogue staging: 8 warps x [32][33] floats
  float* staging = reinterpret_cast<float*>(smem + args.staging_off);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + args.staging_off + args.staging_bytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2]
  uint64_t* slab_full_bar = tmem_empty_bar + 2;          // resident-B slab loaded / free to overwrite
  uint64_t* slab_empty_bar = slab_full_bar + 1;
  uint64_t* epi_bar = slab_empty_bar + 1;                // [8 epilogue warps][2 staging buffers]: inputs landed
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(epi_bar + 16);
  const bool RB = !NT && args.resident_b != 0;
  const uint32_t slab_bytes = RB ? (uint32_t)args.num_kb * b_bytes : 0u;
  // tile walk of this CTA: strided round-robin, or a contiguous run in the resident-B schedule
  const int t_begin = RB ? blockIdx.x * args.tiles_per_cta : (PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x);
  const int t_end = RB ? min(args.total_tiles, t_begin + args.tiles_per_cta) : args.total_tiles;
  const int t_step = RB ? 1 : (PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = args.stages;
  if (threadIdx.x == 0) DBG_T(0);
  // PDL: let the next kernel of the stream be scheduled as SMs free up (it waits for this grid to complete before
  // reading anything); our own set-up below (barriers, TMEM, descriptor prefetch) overlaps the predecessor's tail.
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 2 && lane == 0 && args.epi_tma) {
    if (args.ep.out_f32 != nullptr) tma_prefetch_desc(&tmOutF);
    if (args.ep.out_hi != nullptr) tma_prefetch_desc(&tmOutS);
    if (args.epi_add_kind != 0) tma_prefetch_desc(&tmAdd);
    if (args.ep.mask_hi != nullptr) tma_prefetch_desc(&tmMask);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], PAIR ? 16 : 8);   // one elected lane of each epilogue warp (of both CTAs)
    }
    mbar_init(slab_full_bar, 1);
    mbar_init(slab_empty_bar, 1);
    for (int i = 0; i < 16; ++i) mbar_init(&epi_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) {
      tmem_alloc_pair(tmem_holder, 2 * args.tmem_cols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_holder, 2 * args.tmem_cols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();   // everything the stream predecessor wrote (operands, residuals, accumulation targets) is now visible
  if (threadIdx.x == 0) DBG_T(1);

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int it = 0;
      int cur_n0 = -1;
      uint32_t slab_gen = 0;
      for (int tile = t_begin; tile < t_end; tile += t_step) {
        const TileCoord tc = decode_tile(args, rank, tile);
        const int n0 = tc.n0, m0 = tc.m0, kb_begin = tc.kb_begin, kb_end = tc.kb_end;
        if (RB && n0 != cur_n0) {   // new n-tile: (re)load the weight slab once every MMA on the old one is done
          if (slab_gen > 0) mbar_wait_single(slab_empty_bar, (slab_gen - 1) & 1u);
          mbar_arrive_expect_tx(slab_full_bar, slab_bytes);
          for (int kb = 0; kb < args.num_kb; ++kb)
            tma_load_3d(smem + (size_t)kb * b_bytes, &tmB, slab_full_bar, kb * BK, n0, 0);
          cur_n0 = n0;
          ++slab_gen;
        }
        const int hw = args.cH * args.cW;
        int cb = 0, ch0 = 0, cw0 = 0;    // mode 0 conv: image, first pixel row and first pixel column of this tile
        if (!NT && args.conv) {
          cb = m0 / hw;
          ch0 = (m0 - cb * hw) / args.cW;
          cw0 = m0 - cb * hw - ch0 * args.cW;   // != 0 only for tiles narrower than the image row (tile_rows < 128)
        }
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % stages;
          const uint32_t ph = (uint32_t)(it / stages) & 1u;
          mbar_wait_single(&empty_bar[s], ph ^ 1u);
          uint8_t* a_s = RB ? smem + slab_bytes + (size_t)s * a_bytes : smem + (size_t)s * stage_bytes;
          uint8_t* b_s = a_s + a_bytes;
          if (RB) {
            mbar_arrive_expect_tx(&full_bar[s], args.tx_a);
            tma_load_3d(a_s, &tmA, &full_bar[s], kb * BK, m0, 0);  // box {64, BM, 2}
            continue;
          }
          if (PAIR) {
            // the leader's barrier collects the bytes of both CTAs' loads (its own arrive carries the expectation)
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * (args.tx_a + args.tx_b));
            if (args.conv) {
              const int k0 = kb * BK;
              const int tap = k0 / args.cC;
              const int dh = (tap / 3 - 1) * args.cdil * args.csign, dw = (tap % 3 - 1) * args.cdil * args.csign;
              tma_load_5d_pair(a_s, &tmA, &full_bar[s], k0 - tap * args.cC, cw0 + dw, ch0 + dh, cb, 0);
            } else {
              tma_load_3d_pair(a_s, &tmA, &full_bar[s], kb * BK, m0, 0);           // box {64, 128, 2}
            }
            tma_load_3d_pair(b_s, &tmB, &full_bar[s], kb * BK, n0 + rank * BNL, 0);  // box {64, BN/2, 2}
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[s], args.tx_a + args.tx_b);
          if (!NT) {
            if (args.conv) {   // box {64 c, W, 128/W, 1, 2}: rows of the tile = pixels (h, w) of image cb, shifted by the tap
              const int k0 = kb * BK;
              const int tap = k0 / args.cC;
              const int dh = (tap / 3 - 1) * args.cdil * args.csign, dw = (tap % 3 - 1) * args.cdil * args.csign;
              if (args.tile_rows != BM) {   // narrow tile: one box per plane, the lo plane still lives 128 rows behind hi
                tma_load_5d(a_s, &tmA, &full_bar[s], k0 - tap * args.cC, cw0 + dw, ch0 + dh, cb, 0);
                if (args.pass_mask & 4u)
                  tma_load_5d(a_s + BM * 128, &tmA, &full_bar[s], k0 - tap * args.cC, cw0 + dw, ch0 + dh, cb, 1);
              } else {
                tma_load_5d(a_s, &tmA, &full_bar[s], k0 - tap * args.cC, cw0 + dw, ch0 + dh, cb, 0);
              }
            } else {
              tma_load_3d(a_s, &tmA, &full_bar[s], kb * BK, m0, 0);  // box {64, BM, 2}
            }
            tma_load_3d(b_s, &tmB, &full_bar[s], kb * BK, n0, 0);  // box {64, BN, 2}
          } else if (args.conv) {
            const int krow = kb * BK;                               // first pixel (= k index) of this k block
            if (args.conv_wblk > 0) {   // dy as {n_out, W, H, B}: the 64-pixel box is clipped (zero-filled) at the row end too
              const int rowi = kb / args.conv_wblk;
              const int ab_ = rowi / args.cH;
              for (int c = 0; c < BM / 64; ++c)
                tma_load_5d(a_s + c * 16384, &tmA, &full_bar[s], m0 + 64 * c, (kb - rowi * args.conv_wblk) * 64,
                            rowi - ab_ * args.cH, ab_, 0);
            } else {
              for (int c = 0; c < BM / 64; ++c)                       // box {64(mn), 64(k), 2}
                tma_load_3d(a_s + c * 16384, &tmA, &full_bar[s], m0 + 64 * c, krow, 0);
            }
            // k block = 64 consecutive pixels of one image: box {64 c, min(W,64), 64/min(W,64), 1, 2}
            int pb, ph0, pw0;
            if (args.conv_wblk > 0) {   // k block = 64 pixels of ONE image row starting at 64 * wb; past W the box is zeros
              const int rowi = kb / args.conv_wblk;
              pw0 = (kb - rowi * args.conv_wblk) * 64;
              pb = rowi / args.cH;
              ph0 = rowi - pb * args.cH;
            } else {
              const int p0 = kb * BK;
              pb = p0 / hw;
              const int rem = p0 - pb * hw;
              ph0 = rem / args.cW; pw0 = rem - ph0 * args.cW;
            }
            for (int c = 0; c < (BN + 63) / 64; ++c) {
              const int nn = n0 + 64 * c;
              const int tap = nn / args.cC;
              const int dh = (tap / 3 - 1) * args.cdil, dw = (tap % 3 - 1) * args.cdil;
              tma_load_5d(b_s + c * 16384, &tmB, &full_bar[s], nn - tap * args.cC, pw0 + dw, ph0 + dh, pb, 0);
            }
          } else {
            for (int c = 0; c < BM / 64; ++c)                       // box {64(mn), 64(k), 2}
              tma_load_3d(a_s + c * 16384, &tmA, &full_bar[s], m0 + 64 * c, kb * BK, 0);
            for (int c = 0; c < (BN + 63) / 64; ++c)
              tma_load_3d(b_s + c * 16384, &tmB, &full_bar[s], n0 + 64 * c, kb * BK, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only in a pair) -----
    if (lane == 0 && rank == 0) {
      int it = 0, ti = 0;
      int cur_n0 = -1;
      uint32_t slab_gen = 0;
      for (int tile = t_begin; tile < t_end; tile += t_step, ++ti) {
        const TileCoord tc = decode_tile(args, rank, tile);
        const int kb_begin = tc.kb_begin, kb_end = tc.kb_end;
        if (RB && tc.n0 != cur_n0) {
          mbar_wait_single(slab_full_bar, slab_gen & 1u);
          tc_fence_after();
          cur_n0 = tc.n0;
          ++slab_gen;
        }
        const int ab = ti & 1;
        if (ti >= 2) {  // the epilogue must have drained this accumulator buffer (used by tile ti-2)
          mbar_wait_single(&tmem_empty_bar[ab], (uint32_t)((ti >> 1) - 1) & 1u);
          tc_fence_after();
        }
        const uint32_t tmem_d = tmem_base + (uint32_t)ab * args.tmem_cols;
        uint32_t acc = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % stages;
          const uint32_t ph = (uint32_t)(it / stages) & 1u;
          mbar_wait_single(&full_bar[s], ph);
          tc_fence_after();
          if (it == 0) DBG_T(2);
          const uint32_t a_base = smem_u32(RB ? smem + slab_bytes + (size_t)s * a_bytes : smem + (size_t)s * stage_bytes);
          const uint32_t b_base = RB ? smem_u32(smem + (size_t)kb * b_bytes) : a_base + a_bytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            uint64_t a_hi, a_lo, b_hi, b_lo;
            if (!NT) {
              // K-major SW128: rows at 128 B pitch, 8-row groups 1024 B apart, k-step = +32 B.
              a_hi = make_smem_desc_sw128(a_base + k * 32, 16, 1024);
              a_lo = make_smem_desc_sw128(a_base + BM * 128 + k * 32, 16, 1024);
              b_hi = make_smem_desc_sw128(b_base + k * 32, 16, 1024);
              b_lo = make_smem_desc_sw128(b_base + BNL * 128 + k * 32, 16, 1024);
            } else {
              // MN-major SW128: 64-wide MN chunks 16 KB apart (LBO), 8-k groups 1024 B apart (SBO),
              // one k-step (16) = two k groups = +2048 B; lo plane 8 KB after hi inside a chunk.
              a_hi = make_smem_desc_sw128(a_base + k * 2048, 16384, 1024);
              a_lo = make_smem_desc_sw128(a_base + 8192 + k * 2048, 16384, 1024);
              b_hi = make_smem_desc_sw128(b_base + k * 2048, 16384, 1024);
              b_lo = make_smem_desc_sw128(b_base + 8192 + k * 2048, 16384, 1024);
            }
            // small cross terms first, hi*hi last
            if constexpr (!MASKED) {
              if (PAIR) {
                umma_bf16_ss_pair(tmem_d, a_hi, b_lo, args.idesc, acc);
                acc = 1;
                umma_bf16_ss_pair(tmem_d, a_lo, b_hi, args.idesc, 1);
                umma_bf16_ss_pair(tmem_d, a_hi, b_hi, args.idesc, 1);
              } else {
                umma_bf16_ss(tmem_d, a_hi, b_lo, args.idesc, acc);
                acc = 1;
                umma_bf16_ss(tmem_d, a_lo, b_hi, args.idesc, 1);
                umma_bf16_ss(tmem_d, a_hi, b_hi, args.idesc, 1);
              }
            } else {   // pass_mask drops cross terms (precision-policy measurements)
              if (PAIR) {
                if (args.pass_mask & 2u) { umma_bf16_ss_pair(tmem_d, a_hi, b_lo, args.idesc, acc); acc = 1; }
                if (args.pass_mask & 4u) { umma_bf16_ss_pair(tmem_d, a_lo, b_hi, args.idesc, acc); acc = 1; }
                umma_bf16_ss_pair(tmem_d, a_hi, b_hi, args.idesc, acc);
              } else {
                if (args.pass_mask & 2u) { umma_bf16_ss(tmem_d, a_hi, b_lo, args.idesc, acc); acc = 1; }
                if (args.pass_mask & 4u) { umma_bf16_ss(tmem_d, a_lo, b_hi, args.idesc, acc); acc = 1; }
                umma_bf16_ss(tmem_d, a_hi, b_hi, args.idesc, acc);
              }
              acc = 1;
            }
          }
          // frees this smem stage (in both CTAs of a pair) once the MMAs above have read it
          if (PAIR) umma_commit_pair(&empty_bar[s]);
          else umma_commit(&empty_bar[s]);
        }
        if (RB && (tile + 1 >= t_end || decode_tile(args, rank, tile + 1).n0 != tc.n0))
          umma_commit(slab_empty_bar);    // last MMA reading this weight slab: the producer may overwrite it
        if (PAIR) umma_commit_pair(&tmem_full_bar[ab]);   // accumulator of this tile complete (both CTAs' halves)
        else umma_commit(&tmem_full_bar[ab]);
        if (ti == 0) DBG_T(3);
      }
    }
  } else if (args.epi_tma) {
    // ------------------------------ epilogue (TMA staged) ----------------------
    // Each warp owns 32 accumulator rows (its TMEM lane quarter) and half of the tile's columns, processed in chunks
    // of 32 columns with ONE ROW PER THREAD end to end (no transposition): tcgen05.ld.x32 -> registers; residual /
    // ReLU-mask chunks were fetched by TMA into the warp's swizzled staging buffer (prefetched one chunk ahead);
    // results are written back into the same buffer (16-byte st.shared, conflict-free under the TMA swizzle) and
    // leave by one TMA store (or TMA reduce-add for split-K / accumulate) issued by lane 0.  Global memory is only
    // touched by the TMA unit, in full 32 B sectors, with no address arithmetic or exposed load latency in the warp.
    const EpilogueArgs& ep = args.ep;
    const int w = warp - 2;
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = w >> 2;
    const int cols_half = BN >> 1;      // multiple of 32 (BN % 64 == 0 on this path)
    const int nch = cols_half >> 5;
    const int nb = args.epi_nb;
    uint8_t* wbuf = smem + args.staging_off + (uint32_t)(w * nb) * args.epi_buf_bytes;
    uint64_t* in_bar = epi_bar + 2 * w;
    const int add_kind = args.epi_add_kind;
    const bool has_mask = ep.mask_hi != nullptr;
    const bool has_in = add_kind != 0 || has_mask;
    const uint32_t in_bytes = (add_kind != 0 ? 4096u : 0u) + (has_mask ? 2048u : 0u);
    // byte offset of 16-byte chunk c of this thread's row: 64 B rows (bf16, SWIZZLE_64B) / 128 B rows (fp32, SWIZZLE_128B)
    const uint32_t row64 = (uint32_t)lane * 64u, sw64 = (uint32_t)((lane >> 1) & 3);
    const uint32_t row128 = (uint32_t)lane * 128u, sw128 = (uint32_t)(lane & 7);

    auto chunk_valid = [&](int tile, int ch) -> bool {
      const TileCoord tc = decode_tile(args, rank, tile);
      return tc.m0 + q * 32 < args.M && tc.n0 + half * cols_half + ch * 32 < args.N;
    };
    auto advance = [&](int& tile, int& ch) {   // next chunk of this warp that touches the output at all
      do {
        if (++ch == nch) { ch = 0; tile += t_step; }
      } while (tile < t_end && !chunk_valid(tile, ch));
    };
    auto issue_in = [&](int tile, int ch, int b) {   // lane 0: TMA loads of one chunk's residual / mask
      const TileCoord tc = decode_tile(args, rank, tile);
      const int mr = tc.m0 + q * 32, nc = tc.n0 + half * cols_half + ch * 32;
      uint8_t* buf = wbuf + (uint32_t)b * args.epi_buf_bytes;
      mbar_arrive_expect_tx(&in_bar[b], in_bytes);
      if (add_kind == 1) tma_load_3d(buf, &tmAdd, &in_bar[b], nc, mr, 0);
      else if (add_kind == 2) tma_load_2d(buf, &tmAdd, &in_bar[b], nc, mr);
      if (has_mask) tma_load_2d(buf + args.epi_off_mask, &tmMask, &in_bar[b], nc, mr);
    };

    int ci = 0;   // chunks processed so far by this warp (buffer = ci & 1, barrier parity from ci)
    if (has_in && nb == 2 && lane == 0) {
      int t0 = t_begin, c0 = -1;
      if (t0 < t_end) {
        advance(t0, c0);
        if (t0 < t_end) issue_in(t0, c0, 0);
      }
    }
    int ti = 0;
    for (int tile = t_begin; tile < t_end; tile += t_step, ++ti) {
      const TileCoord tc = decode_tile(args, rank, tile);
      const int n0 = tc.n0, m0 = tc.m0;
      const int ab = ti & 1;
      mbar_wait(&tmem_full_bar[ab], (uint32_t)(ti >> 1) & 1u);
      tc_fence_after();
      if (ti == 0 && warp == 2 && lane == 0) DBG_T(4);
      const int mrow = m0 + q * 32;
      const int row_t = mrow + lane;
      const float rs = (ep.row_scale != nullptr && row_t < args.M) ? ep.row_scale[row_t] : 1.0f;
      const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ab * args.tmem_cols;
      for (int ch = 0; ch < nch; ++ch) {
        const int cl = half * cols_half + ch * 32;
        const int nbase = n0 + cl;
        if (mrow >= args.M || nbase >= args.N) continue;   // warp-uniform, same predicate as chunk_valid
        const int b = nb == 2 ? (ci & 1) : 0;
        uint8_t* buf = wbuf + (uint32_t)b * args.epi_buf_bytes;
        const uint32_t sbuf = smem_u32(buf);
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr_row + (uint32_t)cl, r);
        if (nb == 1) {   // single buffer: the previous chunk's store must have drained before its inputs may land
          if (lane == 0) {
            bulk_wait_group_read<0>();
            if (has_in) issue_in(tile, ch, 0);
          }
          __syncwarp();
        }
        uint4 ain[8], amk[4];
        if (has_in) {
          mbar_wait(&in_bar[b], (uint32_t)(nb == 2 ? (ci >> 1) : ci) & 1u);
          if (add_kind == 1) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t off = row64 + (((uint32_t)c ^ sw64) << 4);
              ain[c] = lds128u(sbuf + off);            // explicit LDS / STS throughout: `buf` has lost its address space
              ain[4 + c] = lds128u(sbuf + 2048 + off);
            }
          } else if (add_kind == 2) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              ain[c] = lds128u(sbuf + row128 + (((uint32_t)c ^ sw128) << 4));
          }
          if (has_mask) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              amk[c] = lds128u(sbuf + args.epi_off_mask + row64 + (((uint32_t)c ^ sw64) << 4));
          }
        }
        if (nb == 2) {
          if (lane == 0) {
            if (has_in) {
              bulk_wait_group_read<0>();   // the store of chunk ci-1 has drained buffer b^1
              int t1 = tile, c1 = ch;
              advance(t1, c1);
              if (t1 < t_end) issue_in(t1, c1, b ^ 1);
            } else {
              bulk_wait_group_read<1>();   // the store of chunk ci-2 has drained buffer b
            }
          }
          __syncwarp();
        }
        tmem_ld_wait();
        if (args.epi_debug >= 3) { ++ci; continue; }
        float v[32];
        if (ep.row_scale != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * rs;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        }
        if (ep.bias != nullptr) {
          if (nbase + 32 <= args.N && (reinterpret_cast<uintptr_t>(ep.bias + nbase) & 15) == 0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + nbase) + c);
              v[4 * c] += bv.x; v[4 * c + 1] += bv.y; v[4 * c + 2] += bv.z; v[4 * c + 3] += bv.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nbase + j < args.N) v[j] += __ldg(ep.bias + nbase + j);
          }
        }
        if (add_kind == 1) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t hw[4] = {ain[c].x, ain[c].y, ain[c].z, ain[c].w};
            const uint32_t lw[4] = {ain[4 + c].x, ain[4 + c].y, ain[4 + c].z, ain[4 + c].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[8 * c + 2 * j] += bf16_bits_to_float(hw[j] & 0xffffu) + bf16_bits_to_float(lw[j] & 0xffffu);
              v[8 * c + 2 * j + 1] += __uint_as_float(hw[j] & 0xffff0000u) + __uint_as_float(lw[j] & 0xffff0000u);
            }
          }
        } else if (add_kind == 2) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            v[4 * c] += __uint_as_float(ain[c].x); v[4 * c + 1] += __uint_as_float(ain[c].y);
            v[4 * c + 2] += __uint_as_float(ain[c].z); v[4 * c + 3] += __uint_as_float(ain[c].w);
          }
        }
        if (ep.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        if (has_mask) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t mw[4] = {amk[c].x, amk[c].y, amk[c].z, amk[c].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (!(bf16_bits_to_float(mw[j] & 0xffffu) > 0.0f)) v[8 * c + 2 * j] = 0.0f;
              if (!(__uint_as_float(mw[j] & 0xffff0000u) > 0.0f)) v[8 * c + 2 * j + 1] = 0.0f;
            }
          }
        }
        if (args.epi_debug == 2) {   // keep the math alive without staging it
          float sacc = 0.0f;
#pragma unroll
          for (int j = 0; j < 32; ++j) sacc += v[j];
          if (sacc == 123.456f) buf[0] = 1;
          ++ci;
          continue;
        }
        uint8_t* sdst = buf;
        uint32_t ssdst = sbuf;
        if (ep.out_f32 != nullptr) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            sts128(sbuf + row128 + (((uint32_t)c ^ sw128) << 4), __float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]),
                   __float_as_uint(v[4 * c + 2]), __float_as_uint(v[4 * c + 3]));
          sdst = buf + args.epi_off_out2;
          ssdst = sbuf + args.epi_off_out2;
        }
        if (ep.out_hi != nullptr) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) split_bf16_pair(v[8 * c + 2 * j], v[8 * c + 2 * j + 1], hw[j], lw[j]);
            const uint32_t off = row64 + (((uint32_t)c ^ sw64) << 4);
            sts128(ssdst + off, hw[0], hw[1], hw[2], hw[3]);
            sts128(ssdst + 2048 + off, lw[0], lw[1], lw[2], lw[3]);
          }
        }
        fence_proxy_async();   // generic-proxy writes above -> visible to the TMA unit
        __syncwarp();
        if (lane == 0 && args.epi_debug != 1) {
          if (ep.out_f32 != nullptr) {
            if (ep.atomic) tma_reduce_add_2d(&tmOutF, buf, nbase, mrow);
            else tma_store_2d(&tmOutF, buf, nbase, mrow);
          }
          if (ep.out_hi != nullptr) tma_store_3d(&tmOutS, sdst, nbase, mrow, 0);
          bulk_commit_group();
        }
        ++ci;
      }
      // all tcgen05.ld of this warp for this accumulator buffer are complete: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(&tmem_empty_bar[ab]);
        else mbar_arrive(&tmem_empty_bar[ab]);
      }
      if (ti == 0 && warp == 2 && lane == 0) DBG_T(5);
    }
    if (lane == 0) bulk_wait_group<0>();   // every store has been performed before the CTA retires
  } else {
    // ------------------------------ epilogue (generic fallback) ----------------
    // Each warp owns 32 accumulator rows (its TMEM lane quarter) and half of the tile's columns.  Per
    // 32-column chunk: TMEM -> registers (one row per thread) -> per-warp shared-memory staging
    // -> re-read with 4 lanes per row so that
    // every global access of the warp covers 8 rows x 128 contiguous bytes (fp32) / 64 bytes (bf16 plane).
    const EpilogueArgs& ep = args.ep;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // two warps share a quarter and split the tile's columns
    const int nchunks = BN / 16;
    const int c_begin = half == 0 ? 0 : ((nchunks + 1) / 2) * 16;
    const int c_end = half == 0 ? ((nchunks + 1) / 2) * 16 : BN;
    float* stg = staging + (warp - 2) * (32 * STG_LD);
    const bool vec_ok = ((ep.ld_out_f32 & 3) == 0) && ((ep.ld_out_split & 7) == 0) &&
                        ((ep.ld_add & 7) == 0) && ((ep.ld_add_f32 & 3) == 0) &&
                        ((ep.ld_mask & 7) == 0);
    const int g8 = (lane & 3) * 8;  // column group of this lane in the coalesced phase
    const int rr = lane >> 2;       // row sub-index 0..7
    int ti = 0;
    for (int tile = t_begin; tile < t_end; tile += t_step, ++ti) {
    const TileCoord tc = decode_tile(args, rank, tile);
    const int n0 = tc.n0, m0 = tc.m0;
    const int ab = ti & 1;
    mbar_wait(&tmem_full_bar[ab], (uint32_t)(ti >> 1) & 1u);
    tc_fence_after();
    if (ti == 0 && warp == 2 && lane == 0) DBG_T(4);
    const int row_t = m0 + q * 32 + lane;  // row held by this thread in the TMEM phase
    const float rs = (ep.row_scale != nullptr && row_t < args.M && q * 32 + lane < args.tile_rows) ? ep.row_scale[row_t] : 1.0f;
    const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ab * args.tmem_cols;
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
      const int cl = c0 + g8;     // tile-local first column of this lane's 8-vector in the coalesced phase
      const int nb = n0 + cl;
      const bool col_ok = cl < c_end && nb < args.N;
      const bool full = vec_ok && (nb + 8 <= args.N);
      // Prefetch this chunk's residual / mask vectors (global memory, L2 or HBM latency) BEFORE the TMEM reads and
      // the shared-memory transpose so that their latency overlaps that work instead of serialising after it.
      uint4 pa0[4], pa1[4], pm[4];
      const bool pre = col_ok && full;
      const bool pre_add = pre && (ep.add_hi != nullptr || ep.add_f32 != nullptr);
      const bool pre_mask = pre && ep.mask_hi != nullptr;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = m0 + q * 32 + rr + 8 * i;
        pa0[i] = make_uint4(0, 0, 0, 0); pa1[i] = pa0[i]; pm[i] = pa0[i];
        if (row < args.M && q * 32 + rr + 8 * i < args.tile_rows) {
          if (pre_add) {
            if (ep.add_hi != nullptr) {
              pa0[i] = __ldg(reinterpret_cast<const uint4*>(ep.add_hi + (int64_t)row * ep.ld_add + nb));
              pa1[i] = __ldg(reinterpret_cast<const uint4*>(ep.add_lo + (int64_t)row * ep.ld_add + nb));
            } else {
              const uint4* pf = reinterpret_cast<const uint4*>(ep.add_f32 + (int64_t)row * ep.ld_add_f32 + nb);
              pa0[i] = __ldg(pf);
              pa1[i] = __ldg(pf + 1);
            }
          }
          if (pre_mask) pm[i] = __ldg(reinterpret_cast<const uint4*>(ep.mask_hi + (int64_t)row * ep.ld_mask + nb));
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (c0 + 16 * u < c_end) {  // warp-uniform
          uint32_t r[16];
          tmem_ld_32x32b_x16(taddr_row + (uint32_t)(c0 + 16 * u), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) stg[lane * STG_LD + 16 * u + j] = __uint_as_float(r[j]) * rs;
        }
      }
      __syncwarp();
      if (col_ok) {
        float bias8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bias8[j] = (ep.bias != nullptr && nb + j < args.N) ? __ldg(ep.bias + nb + j) : 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rl = rr + 8 * i;
          const int row = m0 + q * 32 + rl;
          if (row >= args.M || q * 32 + rl >= args.tile_rows) continue;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = stg[rl * STG_LD + g8 + j] + bias8[j];
          if (ep.add_hi != nullptr) {
            if (full) {
              const uint32_t hw[4] = {pa0[i].x, pa0[i].y, pa0[i].z, pa0[i].w}, lw[4] = {pa1[i].x, pa1[i].y, pa1[i].z, pa1[i].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                v[2 * j] += bf16_bits_to_float(hw[j] & 0xffffu) + bf16_bits_to_float(lw[j] & 0xffffu);
                v[2 * j + 1] += bf16_bits_to_float(hw[j] >> 16) + bf16_bits_to_float(lw[j] >> 16);
              }
            } else {
              const __nv_bfloat16* ph = ep.add_hi + (int64_t)row * ep.ld_add + nb;
              const __nv_bfloat16* pl = ep.add_lo + (int64_t)row * ep.ld_add + nb;
              for (int j = 0; j < 8; ++j)
                if (nb + j < args.N) v[j] += join_bf16(ph[j], pl[j]);
            }
          }
          if (ep.add_f32 != nullptr) {
            const float* pa = ep.add_f32 + (int64_t)row * ep.ld_add_f32 + nb;
            if (full && ep.add_hi == nullptr) {   // prefetched
              v[0] += __uint_as_float(pa0[i].x); v[1] += __uint_as_float(pa0[i].y);
              v[2] += __uint_as_float(pa0[i].z); v[3] += __uint_as_float(pa0[i].w);
              v[4] += __uint_as_float(pa1[i].x); v[5] += __uint_as_float(pa1[i].y);
              v[6] += __uint_as_float(pa1[i].z); v[7] += __uint_as_float(pa1[i].w);
            } else if (full) {
              const float4 t0 = __ldg(reinterpret_cast<const float4*>(pa));
              const float4 t1 = __ldg(reinterpret_cast<const float4*>(pa) + 1);
              v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w;
              v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w;
            } else {
              for (int j = 0; j < 8; ++j)
                if (nb + j < args.N) v[j] += pa[j];
            }
          }
          if (ep.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
          }
          if (ep.mask_hi != nullptr) {
            const __nv_bfloat16* pmk = ep.mask_hi + (int64_t)row * ep.ld_mask + nb;
            if (full) {
              const uint32_t hw[4] = {pm[i].x, pm[i].y, pm[i].z, pm[i].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (!(bf16_bits_to_float(hw[j] & 0xffffu) > 0.0f)) v[2 * j] = 0.0f;
                if (!(bf16_bits_to_float(hw[j] >> 16) > 0.0f)) v[2 * j + 1] = 0.0f;
              }
            } else {
              for (int j = 0; j < 8; ++j)
                if (nb + j < args.N && !(__bfloat162float(pmk[j]) > 0.0f)) v[j] = 0.0f;
            }
          }
          if (ep.out_f32 != nullptr) {
            float* po = ep.out_f32 + (int64_t)row * ep.ld_out_f32 + nb;
            if (ep.atomic) {
              if (full) {
                atomicAdd(reinterpret_cast<float4*>(po), make_float4(v[0], v[1], v[2], v[3]));
                atomicAdd(reinterpret_cast<float4*>(po) + 1, make_float4(v[4], v[5], v[6], v[7]));
              } else {
                for (int j = 0; j < 8; ++j)
                  if (nb + j < args.N) atomicAdd(po + j, v[j]);
              }
            } else if (full) {
              reinterpret_cast<float4*>(po)[0] = make_float4(v[0], v[1], v[2], v[3]);
              reinterpret_cast<float4*>(po)[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else {
              for (int j = 0; j < 8; ++j)
                if (nb + j < args.N) po[j] = v[j];
            }
          }
          if (ep.out_hi != nullptr) {
            __nv_bfloat16* ph = ep.out_hi + (int64_t)row * ep.ld_out_split + nb;
            __nv_bfloat16* pl = ep.out_lo + (int64_t)row * ep.ld_out_split + nb;
            if (full) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                split_bf16_pair(v[2 * j], v[2 * j + 1], hw[j], lw[j]);
              }
              *reinterpret_cast<uint4*>(ph) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(pl) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            } else {
              for (int j = 0; j < 8; ++j)
                if (nb + j < args.N) split_bf16(v[j], ph[j], pl[j]);
            }
          }
        }
      }
      __syncwarp();
    }
    // all tcgen05.ld of this warp for this accumulator buffer are complete: hand it back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
        if (PAIR) mbar_arrive_leader(&tmem_empty_bar[ab]);
        else mbar_arrive(&tmem_empty_bar[ab]);
      }
    if (ti == 0 && warp == 2 && lane == 0) DBG_T(5);
    }  // tile loop
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) DBG_T(6);
  if (PAIR) cluster_sync_all();   // the leader's MMAs have read the peer's shared memory; both may release TMEM
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 2 * args.tmem_cols);
    else tmem_dealloc(tmem_base, 2 * args.tmem_cols);
    if (lane == 0) DBG_T(7);
  }
}

// -------------------------------- host side ---------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 3-D map over a split matrix: dim0 = contiguous index (extent inner), dim1 = rows, dim2 = plane.
int make_split_map(CUtensorMap* map, const cdetr_split_t& t, int64_t inner, int64_t rows,
                   int box_inner, int box_rows, int planes = 2) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    cdetr_set_error("cuTensorMapEncodeTiled entry point unavailable");
    return CDETR_ERR_CUDA;
  }
  CDETR_CHECK_ARG(t.base != nullptr, "gemm: null operand");
  CDETR_CHECK_ARG((reinterpret_cast<uintptr_t>(t.base) & 15) == 0, "gemm: operand not 16B aligned");
  CDETR_CHECK_ARG(t.ld % 8 == 0 && t.plane % 8 == 0 && t.ld >= inner && t.plane > 0,
                  "gemm: operand ld/plane must be multiples of 8 elements (ld=%lld plane=%lld)",
                  (long long)t.ld, (long long)t.plane);
  cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)rows, 2};
  cuuint64_t gstride[2] = {(cuuint64_t)t.ld * 2, (cuuint64_t)t.plane * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, (cuuint32_t)planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, t.base, gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    cdetr_set_error("cuTensorMapEncodeTiled failed (%d) inner=%lld rows=%lld ld=%lld box=%dx%d",
                    (int)r, (long long)inner, (long long)rows, (long long)t.ld, box_inner, box_rows);
    return CDETR_ERR_CUDA;
  }
  return CDETR_OK;
}

// 5-D map over an NHWC split activation [B*H*W, C]: dims {C, W, H, B, plane}; box {64, box_w, box_h, 1, 2}.
int make_conv_map(CUtensorMap* map, const cdetr_split_t& t, int C, int W, int H, int B, int box_w, int box_h,
                  int planes = 2) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    cdetr_set_error("cuTensorMapEncodeTiled entry point unavailable");
    return CDETR_ERR_CUDA;
  }
  CDETR_CHECK_ARG(t.base != nullptr, "gemm: null conv operand");
  CDETR_CHECK_ARG((reinterpret_cast<uintptr_t>(t.base) & 15) == 0, "gemm: conv operand not 16B aligned");
  CDETR_CHECK_ARG(t.ld % 8 == 0 && t.plane % 8 == 0 && t.ld >= C && t.plane > 0,
                  "gemm: conv operand ld/plane must be multiples of 8 elements");
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 2};
  cuuint64_t gstride[4] = {(cuuint64_t)t.ld * 2, (cuuint64_t)W * t.ld * 2, (cuuint64_t)H * W * t.ld * 2,
                           (cuuint64_t)t.plane * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, (cuuint32_t)planes};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, t.base, gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    cdetr_set_error("cuTensorMapEncodeTiled (conv) failed (%d) C=%d W=%d H=%d B=%d ld=%lld box=%dx%d", (int)r, C, W,
                    H, B, (long long)t.ld, box_w, box_h);
    return CDETR_ERR_CUDA;
  }
  return CDETR_OK;
}

// Epilogue maps: one [32 rows x 32 columns] chunk per TMA operation.  fp32 matrix: 2-D {N, M}, 128 B inner box,
// SWIZZLE_128B.  split matrix: 3-D {N, M, plane} bf16, 64 B inner box, SWIZZLE_64B (planes = 2: hi+lo in one box;
// planes = 1 as a 2-D map over plane 0: the ReLU mask).
int make_epi_map_f32(CUtensorMap* map, const float* base, int64_t ld, int N, int M) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    cdetr_set_error("cuTensorMapEncodeTiled entry point unavailable");
    return CDETR_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    cdetr_set_error("cuTensorMapEncodeTiled (epilogue f32) failed (%d) N=%d M=%d ld=%lld", (int)r, N, M, (long long)ld);
    return CDETR_ERR_CUDA;
  }
  return CDETR_OK;
}
int make_epi_map_bf16(CUtensorMap* map, const void* base, int64_t ld, int64_t plane, int N, int M, int planes) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    cdetr_set_error("cuTensorMapEncodeTiled entry point unavailable");
    return CDETR_ERR_CUDA;
  }
  cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)M, 2};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane * 2};
  cuuint32_t box[3] = {32, 32, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, planes == 2 ? 3 : 2, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    cdetr_set_error("cuTensorMapEncodeTiled (epilogue bf16) failed (%d) N=%d M=%d ld=%lld plane=%lld", (int)r, N, M,
                    (long long)ld, (long long)plane);
    return CDETR_ERR_CUDA;
  }
  return CDETR_OK;
}

long long* g_dbg_buf = nullptr;   // debug only: 8 timestamps (ns, %globaltimer) per launch, CTA 0
int g_dbg_cap = 0, g_dbg_next = 0;

}  // namespace

// Debug hook (not part of the reference-facing API): subsequent cdetr_gemm launches record a CTA-0 timeline
// {entry, setup done, first operands landed, MMAs of tile 0 issued, accumulator seen by the epilogue, epilogue of
// tile 0 done, all warps done, TMEM released} into buf[8 * launch]; pass NULL to stop.
extern "C" int cdetr_gemm_debug_timeline(long long* buf, int capacity_launches) {
  g_dbg_buf = buf;
  g_dbg_cap = buf != nullptr ? capacity_launches : 0;
  g_dbg_next = 0;
  return CDETR_OK;
}

extern "C" int cdetr_gemm(const cdetr_gemm_t* g, cdetr_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CDETR_CHECK_ARG(g != nullptr, "gemm: null descriptor");
  CDETR_CHECK_ARG(g->mode == 0 || g->mode == 1, "gemm: bad mode %d", g->mode);
  CDETR_CHECK_ARG(g->M > 0 && g->N > 0 && g->K > 0, "gemm: bad shape %d %d %d", g->M, g->N, g->K);
  CDETR_CHECK_ARG(g->out_f32 != nullptr || g->out_split.base != nullptr, "gemm: no output");
  const bool nt = g->mode == 1;

  int num_sms = 0;
  CDETR_CHECK_CUDA(cdetr_num_sms(&num_sms));
  const CdetrTuning& tune = cdetr_tuning();
  int bn = g->block_n;
  if (bn <= 0) {
    // measured on B200 (tools/gemm_sweep.py, profiles/r01_gemm_sweep_v8.txt).  The GEMM family is bound by the
    // L2 -> SM operand traffic (~8.5 TB/s aggregate for 4-byte split operands), so wide tiles pay whenever the K
    // loop is long enough to amortise them and there are enough tiles to fill the machine
    const int tiles_m = cdiv(g->M, BM);
    // 256-wide tiles need the small epilogue staging (no ReLU-mask input, one output) to keep two operand stages
    const bool wide_ok = !nt && g->mask.base == nullptr && !(g->out_f32 != nullptr && g->out_split.base != nullptr);
    const int tiles256 = tiles_m * cdiv(g->N, 256);
    if (!nt && g->N >= 256 && g->K >= 1024 && (g->N >= 512 || tiles256 >= (3 * num_sms) / 4)) bn = 256;
    else if (wide_ok && g->N >= 512 && g->K >= 256 && tiles256 >= 2 * num_sms) bn = 256;   // sweep v16: 8-18 % faster
    else if (g->N > 64) bn = 128;
    else if (g->N > 32) bn = 64;
    else if (g->N > 16) bn = 32;
    else bn = 16;
    // short-K problems with fewer tiles than SMs are latency-bound: spread them over more CTAs
    if (!nt && g->K <= 256)
      while (bn > 32 && tiles_m * cdiv(g->N, bn) < num_sms) bn >>= 1;
    if (nt && bn < 64) bn = 64;
  }
  // CTA pairs (cta_group::2): 256 x 256 output tiles over two SMs, each staging 128 rows of A and 128 rows of B per
  // k-block: 64 KB per CTA and k-block for twice the MMA work of a 128 x 128 tile (the K <= 1024 shapes are bound by
  // the L2 -> SM operand stream).  CDETR_GEMM_PAIR: 0 never, 1 whenever eligible, unset = heuristic below.
  bool pair = false;
  {
    const int pair_mode = tune.gemm_pair;
    const bool eligible = !nt && g->N >= 256 && g->M >= 256 && g->split_k <= 1 && (g->block_n <= 0 || g->block_n == 256);
    if (pair_mode == 1) pair = eligible;
    else if (pair_mode == -1)   // tools/pair_sweep.py (profiles/r01_pair_sweep_v18.txt): 1.05-1.8x for K >= 512, <= 1.0x below
      pair = eligible && g->N % 256 == 0 && g->K >= 512 && cdiv(g->M, 256) * (g->N / 256) >= num_sms / 4;
    if (pair) bn = 256;
  }
  CDETR_CHECK_ARG(bn >= 16 && bn <= 256 && bn % 16 == 0, "gemm: bad block_n %d", bn);
  CDETR_CHECK_ARG(!nt || bn % 64 == 0, "gemm: mode 1 needs block_n multiple of 64");

  int num_kb = cdiv(g->K, BK);
  int splits = g->split_k > 1 ? g->split_k : 1;
  if (splits > num_kb) splits = num_kb;
  int kb_per_split = cdiv(num_kb, splits);
  splits = cdiv(num_kb, kb_per_split);
  const bool atomic = g->accumulate != 0 || splits > 1;
  if (splits > 1)
    CDETR_CHECK_ARG(g->out_split.base == nullptr && g->bias == nullptr && !g->relu &&
                        g->add_split.base == nullptr && g->add_f32 == nullptr &&
                        g->mask.base == nullptr && g->out_f32 != nullptr,
                    "gemm: split_k only supports (row_scale, atomic out_f32) epilogues");
  if (g->accumulate) CDETR_CHECK_ARG(g->out_f32 != nullptr, "gemm: accumulate needs out_f32");

  const bool conv = g->conv_taps != 0;
  int conv_B = 0, conv_tw = 0, conv_th = 0, conv_wblk = 0;
  if (conv) {
    const int H = g->conv_H, W = g->conv_W, C = g->conv_C;
    CDETR_CHECK_ARG(g->conv_taps == 9 && H > 0 && W > 0 && C > 0 && C % 64 == 0 && g->conv_dil >= 1,
                    "gemm: implicit conv needs 9 taps and C %% 64 == 0 (C=%d)", C);
    CDETR_CHECK_ARG(g->conv_sign == 1 || g->conv_sign == -1, "gemm: conv_sign must be +1 or -1");
    const int64_t pixels = nt ? g->K : g->M;
    CDETR_CHECK_ARG(pixels % ((int64_t)H * W) == 0, "gemm: conv pixel count %lld is not a multiple of H*W",
                    (long long)pixels);
    conv_B = (int)(pixels / ((int64_t)H * W));
    if (!nt) {
      CDETR_CHECK_ARG(g->K == 9 * C, "gemm: conv mode 0 needs K == 9*C");
      // An M tile is th image rows x tw pixels = th * tw LINEAR rows of the [pixels, C] matrix.  Whole 128-row tiles when
      // W | 128 and 128 | H*W; otherwise the largest (tw, th) with tw | W, th | H (th > 1 only when tw == W) and
      // tw * th <= 128: the tile's remaining TMEM lanes are never stored (50 x 50 / 100 x 100 / 200 x 200 maps: 100 rows).
      if (W <= BM && BM % W == 0 && (H * W) % BM == 0) {
        conv_tw = W; conv_th = BM / W;
      } else if (W <= BM) {
        conv_tw = W; conv_th = 1;
        for (int t = BM / W; t >= 1; --t) if (H % t == 0) { conv_th = t; break; }
      } else {
        conv_th = 1; conv_tw = 0;
        for (int t = BM; t >= 1; --t) if (W % t == 0) { conv_tw = t; break; }
      }
      CDETR_CHECK_ARG(conv_tw * conv_th >= 64, "gemm: implicit conv finds no tile of >= 64 rows for a %d x %d map", H, W);
    } else {
      CDETR_CHECK_ARG(g->N == 9 * C && g->conv_sign == 1, "gemm: conv mode 1 needs N == 9*C, conv_sign == 1");
      const int bw = W < 64 ? W : 64;
      // k blocks of 64 consecutive pixels when W | 64 or 64 | W; otherwise 64-pixel blocks per image ROW, the part past W
      // zero-filled by the TMA unit (the matching dy rows then multiply zeros)
      if (!(64 % bw == 0 && W % bw == 0 && (H * W) % 64 == 0)) conv_wblk = (W + 63) / 64;
      CDETR_CHECK_ARG((9 * C) % bn == 0 && C % (bn < 64 ? bn : 64) == 0, "gemm: conv mode 1 needs block_n | 9*C");
    }
  }
  const int tile_rows = (conv && !nt) ? conv_tw * conv_th : BM;
  if (tile_rows != BM) pair = false;          // narrow tiles: single CTA, generic epilogue (long-K GEMMs: the main loop dominates)
  if (conv_wblk > 0) {                        // mode 1, padded k blocks: the contraction runs over B * H * wblk blocks
    num_kb = conv_B * g->conv_H * conv_wblk;
    splits = g->split_k > 1 ? g->split_k : 1;
    if (splits > num_kb) splits = num_kb;
    kb_per_split = cdiv(num_kb, splits);
    splits = cdiv(num_kb, kb_per_split);
  }

  CUtensorMap tmA, tmB;
  int rc;
  // precision policy: an operand's lo plane is only loaded when a product reads it (bit 2: lo_a * hi_b, bit 1: hi_a * lo_b)
  const uint32_t pmask = (g->pass_mask > 0 && g->pass_mask <= 7) ? ((uint32_t)g->pass_mask | 1u) : 7u;
  const int pl_a = (pmask & 4u) ? 2 : 1, pl_b = (pmask & 2u) ? 2 : 1;
  if (!nt) {
    if (conv) {
      if ((rc = make_conv_map(&tmA, g->a, g->conv_C, g->conv_W, g->conv_H, conv_B, conv_tw, conv_th,
                              conv_tw * conv_th != BM ? 1 : pl_a)) != 0)
        return rc;
    } else if ((rc = make_split_map(&tmA, g->a, g->K, g->M, BK, BM, pl_a)) != 0) return rc;
    if ((rc = make_split_map(&tmB, g->b, g->K, g->N, BK, pair ? bn / 2 : bn, pl_b)) != 0) return rc;
  } else {
    if (conv && conv_wblk > 0) {   // dy [pixels, n_out] seen as {n_out, W, H, B}: boxes of 64 pixels clipped at the row end
      CDETR_CHECK_ARG(g->M % 64 == 0, "gemm: conv mode 1 on narrow maps needs n_out %% 64 == 0 (%d)", g->M);
      if ((rc = make_conv_map(&tmA, g->a, g->M, g->conv_W, g->conv_H, conv_B, 64, 1, pl_a)) != 0) return rc;
    } else if ((rc = make_split_map(&tmA, g->a, g->M, g->K, 64, BK, pl_a)) != 0) return rc;
    if (conv) {
      const int bw = conv_wblk > 0 ? 64 : (g->conv_W < 64 ? g->conv_W : 64);
      if ((rc = make_conv_map(&tmB, g->b, g->conv_C, g->conv_W, g->conv_H, conv_B, bw, 64 / bw, pl_b)) != 0) return rc;
    } else if ((rc = make_split_map(&tmB, g->b, g->N, g->K, 64, BK, pl_b)) != 0) return rc;
  }

  KernelArgs ka;
  ka.M = g->M; ka.N = g->N; ka.K = g->K;
  ka.block_n = bn;
  ka.kb_per_split = kb_per_split;
  ka.num_kb = num_kb;
  ka.conv = conv ? 1 : 0;
  ka.cH = conv ? g->conv_H : 1; ka.cW = conv ? g->conv_W : 1; ka.cC = conv ? g->conv_C : 1;
  ka.cdil = g->conv_dil; ka.csign = g->conv_sign;
  ka.idesc = make_idesc_bf16_f32(pair ? 2 * BM : BM, bn, nt ? 1 : 0, nt ? 1 : 0);
  ka.pass_mask = pmask;
  ka.tile_m = pair ? 2 * BM : tile_rows;
  ka.tile_rows = tile_rows;
  ka.conv_wblk = conv_wblk;
  uint32_t cols = 32;
  while ((int)cols < bn) cols <<= 1;
  ka.tmem_cols = cols;
  const uint32_t a_bytes = 2u * BM * 128u;
  const uint32_t b_bytes = nt ? (uint32_t)((bn + 63) / 64) * 16384u : 2u * (uint32_t)(pair ? bn / 2 : bn) * 128u;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  ka.tx_a = a_bytes / 2u * (uint32_t)pl_a;     // both operands are laid out [hi | lo] per tile / per 64-wide chunk
  if (tile_rows != BM) ka.tx_a = (uint32_t)tile_rows * 128u * (uint32_t)pl_a;   // the box delivers tile_rows rows per plane
  ka.tx_b = b_bytes / 2u * (uint32_t)pl_b;
  const uint32_t tail_bytes = (2 * MAX_STAGES + 6 + 16) * 8 + 16;
  const uint32_t smem_max = 227u * 1024u - 1024u - tail_bytes;   // dynamic smem minus alignment slack and barriers

  // ---- epilogue flavour.  TMA-staged (see the kernel) whenever every epilogue tensor is TMA-addressable; the generic
  // per-thread epilogue remains for tiny / unaligned outputs (heads with N = 2, 4) and block_n < 64.
  auto tma_ok_f32 = [](const float* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0; };
  auto tma_ok_split = [](const cdetr_split_t& t) {
    return (reinterpret_cast<uintptr_t>(t.base) & 15) == 0 && t.ld % 8 == 0 && t.plane % 8 == 0 && t.plane > 0;
  };
  const bool has_of = g->out_f32 != nullptr, has_os = g->out_split.base != nullptr;
  const bool has_as = g->add_split.base != nullptr, has_af = g->add_f32 != nullptr, has_mk = g->mask.base != nullptr;
  bool tma_epi = bn % 64 == 0 && !(has_as && has_af) && (!has_of || tma_ok_f32(g->out_f32, g->ld_out_f32)) &&
                 (!has_os || tma_ok_split(g->out_split)) && (!has_as || tma_ok_split(g->add_split)) &&
                 (!has_af || tma_ok_f32(g->add_f32, g->ld_add_f32)) &&
                 (!has_mk || ((reinterpret_cast<uintptr_t>(g->mask.base) & 15) == 0 && g->mask.ld % 8 == 0));
  if (tune.gemm_tma_epi == 0 || tile_rows != BM) tma_epi = false;
  const uint32_t old_staging = 8u * 32u * 33u * 4u;
  uint32_t epi_buf = 4096u + (has_mk ? 2048u : 0u) + ((has_of && has_os) ? 4096u : 0u);
  int epi_nb = 2;
  uint32_t staging_bytes = tma_epi ? 8u * (uint32_t)epi_nb * epi_buf : old_staging;
  if (tma_epi && (smem_max - staging_bytes) / stage_bytes < 2) {   // keep two operand stages: single staging buffer
    epi_nb = 1;
    staging_bytes = 8u * epi_buf;
  }
  if (tma_epi && smem_max < staging_bytes + stage_bytes) {
    tma_epi = false;
    staging_bytes = old_staging;
  }
  const uint32_t smem_budget = smem_max - staging_bytes;
  // Two accumulator buffers per CTA (the epilogue of tile i overlaps the main loop of tile i+1).  One persistent CTA
  // per SM: the kernel's register footprint (168 x 320 threads) does not leave room for a second one, so all of the
  // shared memory goes to the TMA ring (a single-stage ring serialises the k-blocks on the TMA round trip:
  // tools/gemm_floor.py measured 0.9 us per k-block).
  int ctas_per_sm = 1;
  int stages = (int)(smem_budget / stage_bytes);
  if (stages < 1) stages = 1;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2 && smem_budget >= 2 * stage_bytes) stages = 2;
  if (tune.gemm_stages > 0) {  // tuning hook (tools/gemm_sweep.py)
    const int f = tune.gemm_stages;
    if (f >= 1 && f <= MAX_STAGES && (uint32_t)f * stage_bytes <= smem_budget) stages = f;
  }
  CDETR_CHECK_ARG((uint32_t)stages * stage_bytes <= smem_budget, "gemm: tile does not fit shared memory");
  ka.tiles_m = cdiv(g->M, ka.tile_m);
  ka.tiles_n = cdiv(g->N, bn);
  ka.total_tiles = ka.tiles_m * ka.tiles_n * splits;
  // Resident-B schedule: short-K problems with many m-tiles per n-tile keep the [bn x K] weight slab in shared memory
  // and stream only A (the L2 -> SM operand traffic halves).  One CTA per SM.
  // Measured (profiles/r01_bench_c3_v11*.json): with the generic epilogue as the limiter of short-K tiles, one
  // resident-B CTA per SM (8 epilogue warps) lost to two streaming CTAs (16 epilogue warps): 23.0 vs 21.6 ms of GEMM
  // per C3 step.  Off unless CDETR_GEMM_RESIDENT=1.
  bool resident = false;
  {
    const int f = tune.gemm_resident;
    if (f == 1) resident = pmask == 7u && !pair && !nt && !conv && splits == 1 && (uint32_t)num_kb * b_bytes + 2 * a_bytes + 8u * epi_buf <= smem_max;
  }
  ka.resident_b = 0;
  ka.tiles_per_cta = 0;
  size_t operand_bytes = (size_t)stages * stage_bytes;
  if (resident) {
    const uint32_t slab = (uint32_t)num_kb * b_bytes;
    uint32_t budget = smem_budget;
    if (tma_epi && epi_nb == 2 && budget < slab + 2 * a_bytes) {   // trade the second staging buffer for A stages
      epi_nb = 1;
      staging_bytes = 8u * epi_buf;
      budget = smem_max - staging_bytes;
    }
    int st = budget > slab ? (int)((budget - slab) / a_bytes) : 0;
    if (st > MAX_STAGES) st = MAX_STAGES;
    if (tune.gemm_stages > 0) {
      const int f = tune.gemm_stages;
      if (f >= 1 && f <= st) st = f;
    }
    if (st >= 2) {
      stages = st;
      ctas_per_sm = 1;
      ka.resident_b = 1;
      operand_bytes = (size_t)slab + (size_t)stages * a_bytes;
    }
  }
  ka.stages = stages;
  ka.staging_off = (uint32_t)operand_bytes;
  ka.staging_bytes = staging_bytes;
  ka.epi_tma = tma_epi ? 1 : 0;
  ka.epi_nb = epi_nb;
  ka.epi_buf_bytes = epi_buf;
  ka.epi_off_mask = 4096u;
  ka.epi_off_out2 = 4096u + (has_mk ? 2048u : 0u);
  ka.epi_add_kind = has_as ? 1 : (has_af ? 2 : 0);
  ka.epi_debug = tune.gemm_epi_debug;
  ka.dbg = (g_dbg_buf != nullptr && g_dbg_next < g_dbg_cap) ? g_dbg_buf + 8 * (g_dbg_next++) : nullptr;
  const size_t smem_bytes = operand_bytes + staging_bytes + tail_bytes + 1024;

  CUtensorMap tmOutF, tmOutS, tmAdd, tmMask;
  memset(&tmOutF, 0, sizeof(tmOutF)); memset(&tmOutS, 0, sizeof(tmOutS));
  memset(&tmAdd, 0, sizeof(tmAdd)); memset(&tmMask, 0, sizeof(tmMask));
  if (tma_epi) {
    if (has_of && (rc = make_epi_map_f32(&tmOutF, g->out_f32, g->ld_out_f32, g->N, g->M)) != 0) return rc;
    if (has_os && (rc = make_epi_map_bf16(&tmOutS, g->out_split.base, g->out_split.ld, g->out_split.plane, g->N, g->M, 2)) != 0)
      return rc;
    if (has_as && (rc = make_epi_map_bf16(&tmAdd, g->add_split.base, g->add_split.ld, g->add_split.plane, g->N, g->M, 2)) != 0)
      return rc;
    if (has_af && (rc = make_epi_map_f32(&tmAdd, g->add_f32, g->ld_add_f32, g->N, g->M)) != 0) return rc;
    if (has_mk && (rc = make_epi_map_bf16(&tmMask, g->mask.base, g->mask.ld, 8, g->N, g->M, 1)) != 0) return rc;
  }

  EpilogueArgs& ep = ka.ep;
  ep.row_scale = g->row_scale;
  ep.bias = g->bias;
  ep.add_hi = reinterpret_cast<const __nv_bfloat16*>(g->add_split.base);
  ep.add_lo = ep.add_hi ? ep.add_hi + g->add_split.plane : nullptr;
  ep.ld_add = ep.add_hi ? g->add_split.ld : 0;
  ep.add_f32 = g->add_f32;
  ep.ld_add_f32 = g->add_f32 ? g->ld_add_f32 : 0;
  ep.mask_hi = reinterpret_cast<const __nv_bfloat16*>(g->mask.base);
  ep.ld_mask = ep.mask_hi ? g->mask.ld : 0;
  ep.relu = g->relu;
  ep.atomic = atomic ? 1 : 0;
  ep.out_f32 = g->out_f32;
  ep.ld_out_f32 = g->out_f32 ? g->ld_out_f32 : 0;
  ep.out_hi = reinterpret_cast<__nv_bfloat16*>(g->out_split.base);
  ep.out_lo = ep.out_hi ? ep.out_hi + g->out_split.plane : nullptr;
  ep.ld_out_split = ep.out_hi ? g->out_split.ld : 0;
  // vector paths need 16-byte aligned bases
  auto mis = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; };
  if (mis(ep.add_hi) || mis(ep.add_lo) || mis(ep.add_f32) || mis(ep.mask_hi) || mis(ep.out_f32) ||
      mis(ep.out_hi) || mis(ep.out_lo)) {
    cdetr_set_error("gemm: epilogue tensors must be 16-byte aligned");
    return CDETR_ERR_ARG;
  }

  int nctas = num_sms * ctas_per_sm;
  if (nctas > ka.total_tiles) nctas = ka.total_tiles;
  if (pair) {   // clusters of two CTAs, one 256-row tile per cluster at a time
    nctas = 2 * ka.total_tiles;
    if (nctas > (num_sms & ~1)) nctas = num_sms & ~1;
  }
  if (ka.resident_b) {   // contiguous runs of tiles_per_cta tiles (n-major order)
    ka.tiles_per_cta = cdiv(ka.total_tiles, nctas);
    nctas = cdiv(ka.total_tiles, ka.tiles_per_cta);
  }
  dim3 grid(nctas);
  const bool masked = ka.pass_mask != 7u;
  auto kern = masked ? (nt ? gemm_split_kernel<true, false, true>
                           : (pair ? gemm_split_kernel<false, true, true> : gemm_split_kernel<false, false, true>))
                     : (nt ? gemm_split_kernel<true, false, false>
                           : (pair ? gemm_split_kernel<false, true, false> : gemm_split_kernel<false, false, false>));
  static DevAttrCache configured[6] = {};
  CDETR_CHECK_CUDA(cdetr_ensure_smem(kern, 227 * 1024, &configured[(masked ? 3 : 0) + (nt ? 1 : (pair ? 2 : 0))]));
  const int use_pdl = tune.pdl;   // opt-in: measured +0.6 ms on the C3 step (early CTAs of the successor crowd the side streams)
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CDETR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOutF, tmOutS, tmAdd, tmMask, ka));
  return CDETR_OK;
}
