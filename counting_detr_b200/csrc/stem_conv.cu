// ResNet stem as ONE kernel: conv 7x7 stride 2 pad 3 (3 -> 64 channels) + FrozenBN + ReLU
// (A2/models/resnet.py:263-271 conv1/bn1/relu, A2/models/backbone.py:22-60 FrozenBatchNorm2d), NCHW fp32 image ->
// NHWC split-bf16 activation.
//
// Round 1 lowered the stem as im2col + GEMM: the im2col matrix of a 16 x 512 x 512 batch is 1 M rows x 152 split-bf16
// columns = 637 MB written and read back for 20 GFLOP of work (0.55 ms of a 18.7 ms step, 0.4 ms of it on the critical
// path).  Here the im2col rows never exist in memory: a CTA takes an 8 x 16 patch of output pixels (= 128 GEMM rows =
// 128 TMEM lanes), stages the 21 x 37 x 3 input window it touches in shared memory (cp.async, next patch prefetched),
// and every thread builds the 147 (padded to 160) window values of ITS pixel as a split-bf16 A operand row directly
// in tensor memory (tcgen05.st); one thread multiplies it with the FrozenBN-folded weights ([64 x 160] split bf16,
// K-major SWIZZLE_128B in shared memory, staged once per CTA) by tcgen05.mma with the A operand read from TMEM (three
// bf16 products hi*lo + lo*hi + hi*hi, fp32 accumulation); the epilogue adds the FrozenBN shift, applies the ReLU and
// leaves through a swizzled shared-memory staging tile and ONE TMA store per plane and patch (box {64 c, 16 x, 8 y}, edges
// clipped by the TMA unit): per-thread 16-byte global stores (32 rows per warp instruction) kept the L1TEX pipe busy
// enough to delay the window loads behind them.  HBM traffic: the image once (+ halo re-reads from L2) and the
// activation once.
// One CTA per SM, 16 builder warps (4 per TMEM lane quarter, a quarter of the window values each) + the MMA warp; the
// operand rows and the accumulator are double-buffered in TMEM (2 x 160 + 2 x 64 of the 512 columns), so the tensor core
// works on tile t while the builders store tile t-1 and build tile t+1: nobody waits for an MMA in steady state
// (the first version - 8 builder warps, single buffers, two CTAs per SM - spent 23 % of its stall samples and 20 % of
// its issued instructions in barrier spins: 207 us; profiles/r02_final_ncu_attn_stem_summary.txt).
#include "common.cuh"
#include "../../include/cdetr.h"

#if defined(CDETR_STEM_SPIN) && CDETR_STEM_SPIN
#define STEM_WAIT mbar_wait
#else
#define STEM_WAIT mbar_wait_sleep
#endif

namespace {

constexpr int PH = 8, PW = 16;                 // output patch = 128 pixels
constexpr int IH = 2 * PH + 5;                 // 21 input rows
constexpr int ROW_LD = 40;                     // floats per staged input row: [even 19 | pad | odd 18 at +20]; 2*40 = 16 mod 32
constexpr int PATCH_FLOATS = 2528;              // 3 * 21 * 40 = 2520 floats, padded to a 128-byte multiple
constexpr int KP = 160;                        // 147 window values padded to 10 k-steps of 16
constexpr int KSTEPS = KP / 16;
constexpr uint32_t W_PLANE_BYTES = 3 * 64 * 128;   // 3 k-blocks of [64 rows x 64 k] bf16, SW128 K-major
constexpr uint32_t COL_ACC = 0, COL_A = 128, A_COLS = KP;   // acc buffer i at 64 i; operand buffer i at 128 + 160 i: hi 80 | lo 80
constexpr uint32_t OUT_PLANE_BYTES = 128 * 64 * 2;          // one plane of a patch's output tile: 16 KB
constexpr int NPATCH = 3;                                   // input-window buffers: two tiles of cp.async lead
constexpr int NBUILD = 16;                                  // builder warps

struct StemArgs {
  const float* img;
  const __nv_bfloat16 *w_hi, *w_lo;   // [64][ld_w] FrozenBN-folded weights, k = (r*7 + s)*3 + c
  int64_t ld_w;
  const float* shift;                 // [64]
  __nv_bfloat16 *o_hi, *o_lo;
  int64_t ld_o;
  int B, H, W, Ho, Wo, tiles_y, tiles_x, ntiles;
  uint32_t idesc;
};

// Staged window layouts (floats).  The 37 input columns a patch touches are de-interleaved by parity so that the 16
// pixels of a patch row read consecutive words (the two patch rows of a warp land 16 banks apart).
//   [c][r][ROW_LD = 40]: even columns at +0, odd at +20.  (A TMA box cannot do this: the TMA unit has no element stride
//   along the innermost dimension; the window is staged with 4-byte cp.async, two tiles ahead.)
// window value k of the pixel whose de-interleaved window starts at `base` (byte address in the staged patch)
template <int K>
__device__ __forceinline__ float window_value(uint32_t base) {
  if (K >= 147) return 0.0f;
  constexpr int c = K % 3, t = K / 3, s = t % 7, r = t / 7;
  return lds32(base + 4u * (uint32_t)((c * IH + r) * ROW_LD + (s & 1) * 20 + (s >> 1)));
}

template <int K0>
__device__ __forceinline__ void build_k8(uint32_t base, uint32_t taddr_hi, uint32_t taddr_lo) {
  uint32_t hw[4], lw[4];
  split_bf16_pair(window_value<K0 + 0>(base), window_value<K0 + 1>(base), hw[0], lw[0]);
  split_bf16_pair(window_value<K0 + 2>(base), window_value<K0 + 3>(base), hw[1], lw[1]);
  split_bf16_pair(window_value<K0 + 4>(base), window_value<K0 + 5>(base), hw[2], lw[2]);
  split_bf16_pair(window_value<K0 + 6>(base), window_value<K0 + 7>(base), hw[3], lw[3]);
  tmem_st_32x32b_x4(taddr_hi + (uint32_t)(K0 / 2), hw);
  tmem_st_32x32b_x4(taddr_lo + (uint32_t)(K0 / 2), lw);
}
template <int K0>
__device__ __forceinline__ void build_k40(uint32_t base, uint32_t t_hi, uint32_t t_lo) {
  build_k8<K0>(base, t_hi, t_lo);      build_k8<K0 + 8>(base, t_hi, t_lo);  build_k8<K0 + 16>(base, t_hi, t_lo);
  build_k8<K0 + 24>(base, t_hi, t_lo); build_k8<K0 + 32>(base, t_hi, t_lo);
}

__global__ void __launch_bounds__(32 * NBUILD + 32, 1)
stem_conv_kernel(const __grid_constant__ CUtensorMap tmOut, const StemArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Os = smem;                                               // [2 buffers][2 planes][128 pixels][64 c] bf16, SW128: output staging
  uint8_t* Ws = Os + 4 * OUT_PLANE_BYTES;                           // [2 planes][3 k-blocks][64][64] bf16 SW128
  float* patch = reinterpret_cast<float*>(Ws + 2 * W_PLANE_BYTES);  // [NPATCH buffers][3][IH][ROW_LD]
  uint64_t* bars = reinterpret_cast<uint64_t*>(patch + NPATCH * PATCH_FLOATS);
  uint64_t* a_full = bars;        // [2] operand rows of the tile are in TMEM buffer i (16 builder warps)
  uint64_t* mma_done = bars + 2;  // [2] accumulator i complete, operand buffer i free
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 4);
  float* shift_s = reinterpret_cast<float*>(bars + 6);    // FrozenBN shift, read back as warp-wide broadcasts
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], NBUILD);
      mbar_init(&mma_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == NBUILD) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  if (threadIdx.x < 64) shift_s[threadIdx.x] = __ldg(a.shift + threadIdx.x);
  // tiles of this CTA (blockIdx.x, + gridDim.x, ...), decoded once
  int2* tiles = reinterpret_cast<int2*>(shift_s + 64);
  const int n_my = a.ntiles > (int)blockIdx.x ? (a.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  for (int i = threadIdx.x; i < n_my; i += blockDim.x) {
    const int tile = blockIdx.x + i * gridDim.x;
    const int tpi = a.tiles_y * a.tiles_x;
    const int b = tile / tpi, tr = tile % tpi;
    tiles[i] = make_int2(b, ((tr / a.tiles_x) << 16) | (tr % a.tiles_x));
  }
  // weights -> shared memory in the K-major SWIZZLE_128B operand layout (16-byte chunks; k >= ld_w is zero padding)
  for (int i = threadIdx.x; i < 2 * 64 * 24; i += blockDim.x) {
    const int j = i % 24, n = (i / 24) % 64, pl = i / (24 * 64);     // chunk j = k 8j .. 8j+7
    uint4 v = make_uint4(0, 0, 0, 0);
    if (8 * j + 8 <= a.ld_w) v = __ldg(reinterpret_cast<const uint4*>((pl ? a.w_lo : a.w_hi) + n * a.ld_w + 8 * j));
    const int kb = j >> 3, jj = j & 7;
    const uint32_t off = (uint32_t)pl * W_PLANE_BYTES + (uint32_t)kb * 8192u + (uint32_t)n * 128u + (uint32_t)((jj ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(Ws + off) = v;
  }
  fence_proxy_async();      // generic-proxy writes of the weights -> visible to the tensor core's shared-memory reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == NBUILD) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      const uint32_t w_base = smem_u32(Ws);
      for (int it = 0; it < n_my; ++it) {
        const uint32_t buf = (uint32_t)it & 1u;
        STEM_WAIT(&a_full[buf], (uint32_t)(it >> 1) & 1u);
        tc_fence_after();
        const uint32_t acc = tmem_base + COL_ACC + 64u * buf;
        const uint32_t a0 = tmem_base + COL_A + A_COLS * buf;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const uint32_t boff = (uint32_t)(ks >> 2) * 8192u + (uint32_t)(ks & 3) * 32u;
          const uint64_t b_hi = make_smem_desc_sw128(w_base + boff, 16, 1024);
          const uint64_t b_lo = make_smem_desc_sw128(w_base + W_PLANE_BYTES + boff, 16, 1024);
          const uint32_t a_hi = a0 + (uint32_t)ks * 8u;
          const uint32_t a_lo = a_hi + KP / 2;
          umma_bf16_ts(acc, a_hi, b_lo, a.idesc, ks > 0 ? 1u : 0u);
          umma_bf16_ts(acc, a_lo, b_hi, a.idesc, 1);
          umma_bf16_ts(acc, a_hi, b_hi, a.idesc, 1);
        }
        umma_commit(&mma_done[buf]);
      }
    }
  } else {
    // ------------------------------ builders / epilogue (16 warps) ------------------------------
    const int quarter = warp & 3;           // TMEM lane quarter of this warp
    const int kq = warp >> 2;               // k 40 kq .. 40 kq + 39 (build), output channels 16 kq .. 16 kq + 15 (epilogue)
    const int ml = quarter * 32 + lane;     // pixel of the patch = TMEM lane
    const int py = ml / PW, px = ml % PW;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // this CTA's tiles, decoded once: {sample b, (ty << 16) | tx}; no divisions in the tile loop
    auto tile_info = [&](int i) -> int2 { return tiles[i]; };

    // input window of one patch -> shared memory (zeros outside the image).  Warp w takes the (channel, row) pairs
    // w, w + 16, ... of the 63; lane l the columns l and l + 32.  Everything that does not depend on the tile is hoisted.
    const int sx0 = lane, sx1 = lane + 32;
    const uint32_t so0 = 4u * (uint32_t)((sx0 & 1) * 20 + (sx0 >> 1)), so1 = 4u * (uint32_t)((sx1 & 1) * 20 + (sx1 >> 1));
    int prow[4], pimg[4];                      // window row r and c * H * W of this warp's pairs
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pr = warp + NBUILD * j;
      prow[j] = pr % IH;
      pimg[j] = (pr / IH) * a.H * a.W;
    }
    auto stage = [&](int i, uint32_t dst) {    // dst: shared-memory address of the patch buffer
      const int2 ti = tile_info(i);
      const int iy0 = (ti.y >> 16) * (PH * 2) - 3, ix0 = (ti.y & 0xffff) * (PW * 2) - 3;
      const bool ok0 = ix0 + sx0 >= 0 && ix0 + sx0 < a.W;
      const bool ok1 = sx1 < 37 && ix0 + sx1 >= 0 && ix0 + sx1 < a.W;
      const float* img_b = a.img + ti.x * (3 * a.H * a.W) + ix0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pr = warp + NBUILD * j;
        if (pr < 3 * IH) {
          const int iy = iy0 + prow[j];
          const bool rok = iy >= 0 && iy < a.H;
          const float* src = img_b + (pimg[j] + iy * a.W);
          const uint32_t d = dst + (uint32_t)pr * (ROW_LD * 4u);
          if (rok && ok0) cp_async_4s(d + so0, src + sx0); else sts32(d + so0, 0.0f);
          if (sx1 < 37) { if (rok && ok1) cp_async_4s(d + so1, src + sx1); else sts32(d + so1, 0.0f); }
        }
      }
    };
    // accumulator row of this pixel (16 channels per thread) of tile i from accumulator buffer `buf` -> staging tile `buf`
    // (row = pixel, 128 bytes per plane, 16-byte chunks XOR-swizzled by the row as the TMA store's SWIZZLE_128B expects)
    const uint32_t os = smem_u32(Os);
    auto store_tile = [&](uint32_t buf) {
      uint32_t t[16];
      tmem_ld_32x32b_x16(lane_addr + COL_ACC + 64u * buf + (uint32_t)kq * 16u, t);
      tmem_ld_wait();
      tc_fence_before();      // a later tile's MMAs (ordered behind an a_full arrival of this warp) overwrite these columns
      const uint32_t row = os + buf * (2u * OUT_PLANE_BYTES) + (uint32_t)ml * 128u;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float4 s0 = lds128(smem_u32(shift_s + kq * 16 + g * 8));
        const float4 s1 = lds128(smem_u32(shift_s + kq * 16 + g * 8 + 4));
        const float sh[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v0 = fmaxf(__uint_as_float(t[g * 8 + 2 * j]) + sh[2 * j], 0.0f);
          const float v1 = fmaxf(__uint_as_float(t[g * 8 + 2 * j + 1]) + sh[2 * j + 1], 0.0f);
          split_bf16_pair(v0, v1, hw[j], lw[j]);
        }
        const uint32_t off = (uint32_t)(((2 * kq + g) ^ (ml & 7)) << 4);
        sts128(row + off, hw[0], hw[1], hw[2], hw[3]);
        sts128(row + OUT_PLANE_BYTES + off, lw[0], lw[1], lw[2], lw[3]);
      }
      fence_proxy_async();    // generic-proxy writes -> visible to the TMA unit (after the CTA barrier that follows)
    };
    // one thread: TMA stores of tile i's staging tile (both planes); coordinates {c, x, y, sample, plane}
    auto issue_store = [&](int i) {
      const int2 ti = tile_info(i);
      const uint32_t src = os + (uint32_t)(i & 1) * (2u * OUT_PLANE_BYTES);
      tma_store_5d(&tmOut, src, 0, (ti.y & 0xffff) * PW, (ti.y >> 16) * PH, ti.x, 0);
      tma_store_5d(&tmOut, src + OUT_PLANE_BYTES, 0, (ti.y & 0xffff) * PW, (ti.y >> 16) * PH, ti.x, 1);
      bulk_commit_group();
    };
    const uint32_t patch_s = smem_u32(patch);
    // windows travel two tiles ahead (cp.async path: one group per tile, empty groups past the end keep the count uniform)
    if (n_my > 0) stage(0, patch_s);
    cp_async_commit();
    if (n_my > 1) stage(1, patch_s + PATCH_FLOATS * 4u);
    cp_async_commit();
    for (int it = 0; it < n_my; ++it) {
      const uint32_t buf = (uint32_t)it & 1u;
      const uint32_t pbuf = (uint32_t)(it % NPATCH);
      if (threadIdx.x == 0) bulk_wait_group_read<0>();     // the stores issued one iteration ago have left their staging tile
      cp_async_wait_group<1>();                            // this tile's window has landed (the next one may be in flight)
      asm volatile("bar.sync 1, 512;" ::: "memory");       // ... for every thread; window buffer (it+2) % 3 is no longer read;
                                                           // the staging tile of tile it-2 is complete
      if (threadIdx.x == 0 && it >= 2) issue_store(it - 2);
      // operand buffer `buf` was read by the MMAs of tile it-2: this thread saw them complete before it staged tile it-2
      // ---- build this pixel's operand row: window value k = patch[c][2 py + r][2 px + s]
      const uint32_t base = patch_s + pbuf * (PATCH_FLOATS * 4u) + 4u * (uint32_t)(2 * py * ROW_LD + px);
      const uint32_t t_hi = lane_addr + COL_A + A_COLS * buf, t_lo = t_hi + KP / 2;
      if (kq == 0) build_k40<0>(base, t_hi, t_lo);
      else if (kq == 1) build_k40<40>(base, t_hi, t_lo);
      else if (kq == 2) build_k40<80>(base, t_hi, t_lo);
      else build_k40<120>(base, t_hi, t_lo);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[buf]);
      if (it + 2 < n_my) stage(it + 2, patch_s + (uint32_t)((it + 2) % NPATCH) * (PATCH_FLOATS * 4u));
      cp_async_commit();
      // ---- stage the PREVIOUS tile's output: its MMAs ran while this one was built; its staging tile (buf ^ 1) was last
      // read by the TMA stores of tile it-3, drained before this iteration's barrier
      if (it > 0) {
        STEM_WAIT(&mma_done[buf ^ 1u], (uint32_t)((it - 1) >> 1) & 1u);
        tc_fence_after();
        store_tile(buf ^ 1u);
      }
    }
    // drain: tiles n_my-2 (staged, not yet stored) and n_my-1 (MMAs in flight)
    if (threadIdx.x == 0) bulk_wait_group_read<0>();
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (threadIdx.x == 0 && n_my >= 2) issue_store(n_my - 2);
    if (n_my > 0) {
      const uint32_t buf = (uint32_t)(n_my - 1) & 1u;
      STEM_WAIT(&mma_done[buf], (uint32_t)((n_my - 1) >> 1) & 1u);
      tc_fence_after();
      store_tile(buf);
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (threadIdx.x == 0) issue_store(n_my - 1);
    }
    if (threadIdx.x == 0) bulk_wait_group<0>();            // every store has been performed before the CTA retires
    cp_async_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NBUILD) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*StemEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
StemEncodeFn stem_encode_fn() {
  static StemEncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<StemEncodeFn>(p);
  }
  return fn;
}

}  // namespace

// out[(b, oy, ox), 0:64] = relu(conv7x7s2p3(img)[b, :, oy, ox] + shift): w = FrozenBN-scaled weights [64, ld_w >= 152]
// split bf16 with k = (r*7 + s)*3 + c (the layout cdetr_mt_pack_weights writes for the stem), shift fp32 [64].
extern "C" int cdetr_stem_conv(const float* img, int B, int H, int W, cdetr_split_t w, const float* shift,
                               cdetr_split_t out, cdetr_stream_t s) {
  CDETR_CHECK_ARG(img && w.base && shift && out.base && B > 0 && H > 0 && W > 0, "stem_conv: bad args");
  CDETR_CHECK_ARG(w.ld >= 152 && w.ld % 8 == 0 && w.plane % 8 == 0 && (reinterpret_cast<uintptr_t>(w.base) & 15) == 0,
                  "stem_conv: weights must be [64, ld >= 152] split bf16, 16-byte aligned rows");
  CDETR_CHECK_ARG(out.ld % 8 == 0 && out.ld >= 64 && out.plane % 8 == 0 && (reinterpret_cast<uintptr_t>(out.base) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(shift) & 15) == 0,
                  "stem_conv: output rows / shift must be 16-byte aligned");
  StemArgs a = {};
  a.img = img;
  a.w_hi = reinterpret_cast<const __nv_bfloat16*>(w.base); a.w_lo = a.w_hi + w.plane; a.ld_w = w.ld;
  a.shift = shift;
  a.o_hi = reinterpret_cast<__nv_bfloat16*>(out.base); a.o_lo = a.o_hi + out.plane; a.ld_o = out.ld;
  a.B = B; a.H = H; a.W = W;
  a.Ho = (H + 6 - 7) / 2 + 1; a.Wo = (W + 6 - 7) / 2 + 1;
  a.tiles_y = (a.Ho + PH - 1) / PH; a.tiles_x = (a.Wo + PW - 1) / PW;
  a.ntiles = B * a.tiles_y * a.tiles_x;
  a.idesc = make_idesc_bf16_f32(128, 64, 0, 0);
  CDETR_CHECK_ARG((int64_t)B * 3 * H * W < (int64_t)1 << 31 && a.tiles_y < 65536 && a.tiles_x < 65536, "stem_conv: batch too large for 32-bit image offsets");
  int num_sms = 0;
  CDETR_CHECK_CUDA(cdetr_num_sms(&num_sms));
  const int grid = a.ntiles < num_sms ? a.ntiles : num_sms;
  const int n_my_max = (a.ntiles + grid - 1) / grid;
  const size_t smem = 4 * OUT_PLANE_BYTES + 2 * W_PLANE_BYTES + NPATCH * PATCH_FLOATS * sizeof(float) + 96 + 256 + (size_t)n_my_max * 8 + 1024;
  CDETR_CHECK_ARG(smem <= 200 * 1024, "stem_conv: too many tiles per CTA (%d)", n_my_max);
  // output map {64 c, Wo, Ho, B, 2 planes}: one box {64, 16, 8, 1, 1} = a patch's tile of one plane, edges clipped
  StemEncodeFn enc = stem_encode_fn();
  if (!enc) { cdetr_set_error("cuTensorMapEncodeTiled entry point unavailable"); return CDETR_ERR_CUDA; }
  CDETR_CHECK_ARG(out.ld == 64, "stem_conv: output rows must be dense (ld == 64)");
  CUtensorMap tm;
  {
    cuuint64_t gdim[5] = {64, (cuuint64_t)a.Wo, (cuuint64_t)a.Ho, (cuuint64_t)B, 2};
    cuuint64_t gstr[4] = {(cuuint64_t)out.ld * 2, (cuuint64_t)a.Wo * out.ld * 2, (cuuint64_t)a.Ho * a.Wo * out.ld * 2,
                          (cuuint64_t)out.plane * 2};
    cuuint32_t box[5] = {64, PW, PH, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, out.base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { cdetr_set_error("stem_conv: cuTensorMapEncodeTiled failed (%d) Ho=%d Wo=%d", (int)r, a.Ho, a.Wo); return CDETR_ERR_CUDA; }
  }
  static DevAttrCache cfg = {};
  CDETR_CHECK_CUDA(cdetr_ensure_smem(stem_conv_kernel, 200 * 1024, &cfg));
  stem_conv_kernel<<<grid, 32 * NBUILD + 32, smem, reinterpret_cast<cudaStream_t>(s)>>>(tm, a);
  CDETR_CHECK_LAUNCH();
  return 0;
}
