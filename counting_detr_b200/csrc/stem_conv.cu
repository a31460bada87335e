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
// writes split-bf16 NHWC rows.  HBM traffic: the image once (+ halo re-reads from L2) and the activation once.
// TMEM per CTA: 64 accumulator + 80 + 80 operand columns (256 allocated: two CTAs per SM overlap build / MMA / store).
#include "common.cuh"
#include "../../include/cdetr.h"

namespace {

constexpr int PH = 8, PW = 16;                 // output patch = 128 pixels
constexpr int IH = 2 * PH + 5;                 // 21 input rows
constexpr int IWH = PW + 3;                    // 19 columns per parity (37 input columns de-interleaved: even | odd)
constexpr int ROW_LD = 40;                     // floats per staged input row: [even 19 | pad | odd 18 at +20]; 2*40 = 16 mod 32
constexpr int PATCH_FLOATS = 3 * IH * ROW_LD;  // 2520
constexpr int KP = 160;                        // 147 window values padded to 10 k-steps of 16
constexpr int KSTEPS = KP / 16;
constexpr uint32_t W_PLANE_BYTES = 3 * 64 * 128;   // 3 k-blocks of [64 rows x 64 k] bf16, SW128 K-major
constexpr uint32_t COL_ACC = 0, COL_AHI = 64, COL_ALO = 64 + KP / 2;

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

// window value k of the pixel whose de-interleaved window starts at `base` (float index in the staged patch)
template <int K>
__device__ __forceinline__ float window_value(uint32_t base) {
  if (K >= 147) return 0.0f;
  constexpr int c = K % 3, t = K / 3, s = t % 7, r = t / 7;
  return lds32(base + 4u * (uint32_t)((c * IH + r) * ROW_LD + (s & 1) * 20 + (s >> 1)));
}

template <int K0>
__device__ __forceinline__ void build_k16(uint32_t base, uint32_t taddr_hi, uint32_t taddr_lo) {
  uint32_t hw[8], lw[8];
  float v[16];
  // template recursion by hand: 16 consecutive k
  v[0] = window_value<K0 + 0>(base);   v[1] = window_value<K0 + 1>(base);
  v[2] = window_value<K0 + 2>(base);   v[3] = window_value<K0 + 3>(base);
  v[4] = window_value<K0 + 4>(base);   v[5] = window_value<K0 + 5>(base);
  v[6] = window_value<K0 + 6>(base);   v[7] = window_value<K0 + 7>(base);
  v[8] = window_value<K0 + 8>(base);   v[9] = window_value<K0 + 9>(base);
  v[10] = window_value<K0 + 10>(base); v[11] = window_value<K0 + 11>(base);
  v[12] = window_value<K0 + 12>(base); v[13] = window_value<K0 + 13>(base);
  v[14] = window_value<K0 + 14>(base); v[15] = window_value<K0 + 15>(base);
#pragma unroll
  for (int i = 0; i < 8; ++i) split_bf16_pair(v[2 * i], v[2 * i + 1], hw[i], lw[i]);
  tmem_st_32x32b_x8(taddr_hi + (uint32_t)(K0 / 2), hw);
  tmem_st_32x32b_x8(taddr_lo + (uint32_t)(K0 / 2), lw);
}

__global__ void __launch_bounds__(288, 2)
stem_conv_kernel(const StemArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Ws = smem;                                               // [2 planes][3 k-blocks][64][64] bf16 SW128
  float* patch = reinterpret_cast<float*>(Ws + 2 * W_PLANE_BYTES);  // [2 buffers][3][IH][ROW_LD]
  uint64_t* bars = reinterpret_cast<uint64_t*>(patch + 2 * PATCH_FLOATS);
  uint64_t* a_full = bars;        // operand rows of the tile are in TMEM (8 builder warps)
  uint64_t* mma_done = bars + 1;  // accumulator complete, operand columns free
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(a_full, 8);
    mbar_init(mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_holder, 256);
    tmem_relinquish();
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

  const int t_begin = blockIdx.x, t_step = gridDim.x;
  if (warp == 8) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      const uint32_t w_base = smem_u32(Ws);
      int it = 0;
      for (int tile = t_begin; tile < a.ntiles; tile += t_step, ++it) {
        mbar_wait_sleep(a_full, (uint32_t)it & 1u);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const uint32_t boff = (uint32_t)(ks >> 2) * 8192u + (uint32_t)(ks & 3) * 32u;
          const uint64_t b_hi = make_smem_desc_sw128(w_base + boff, 16, 1024);
          const uint64_t b_lo = make_smem_desc_sw128(w_base + W_PLANE_BYTES + boff, 16, 1024);
          const uint32_t a_hi = tmem_base + COL_AHI + (uint32_t)ks * 8u;
          const uint32_t a_lo = tmem_base + COL_ALO + (uint32_t)ks * 8u;
          umma_bf16_ts(tmem_base + COL_ACC, a_hi, b_lo, a.idesc, ks > 0 ? 1u : 0u);
          umma_bf16_ts(tmem_base + COL_ACC, a_lo, b_hi, a.idesc, 1);
          umma_bf16_ts(tmem_base + COL_ACC, a_hi, b_hi, a.idesc, 1);
        }
        umma_commit(mma_done);
      }
    }
  } else {
    // ------------------------------ builders / epilogue (8 warps) ------------------------------
    const int ct = threadIdx.x;             // 0..255
    const int quarter = warp & 3;           // TMEM lane quarter of this warp
    const int khalf = warp >> 2;            // k 0..79 / 80..159 (build), output channels 0..31 / 32..63 (epilogue)
    const int ml = quarter * 32 + lane;     // pixel of the patch = TMEM lane
    const int py = ml / PW, px = ml % PW;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int tiles_per_img = a.tiles_y * a.tiles_x;

    auto stage = [&](int tile, float* dst) {     // input window of one patch -> shared memory (zeros outside the image)
      const int b = tile / tiles_per_img, tr = tile % tiles_per_img;
      const int iy0 = (tr / a.tiles_x) * PH * 2 - 3, ix0 = (tr % a.tiles_x) * PW * 2 - 3;
      for (int i = ct; i < 3 * IH * 37; i += 256) {
        const int x = i % 37, r = (i / 37) % IH, c = i / (37 * IH);
        const int iy = iy0 + r, ix = ix0 + x;
        float* d = dst + (c * IH + r) * ROW_LD + (x & 1) * 20 + (x >> 1);
        if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) cp_async_4(d, a.img + (((int64_t)b * 3 + c) * a.H + iy) * a.W + ix);
        else *d = 0.0f;
      }
    };

    if (t_begin < a.ntiles) stage(t_begin, patch);
    int it = 0;
    for (int tile = t_begin; tile < a.ntiles; tile += t_step, ++it) {
      float* cur = patch + (it & 1) * PATCH_FLOATS;
      cp_async_wait_all();
      asm volatile("bar.sync 1, 256;" ::: "memory");       // this patch has landed; the other buffer is no longer read
      if (tile + t_step < a.ntiles) stage(tile + t_step, patch + ((it + 1) & 1) * PATCH_FLOATS);
      // ---- build this pixel's operand row: window value k = patch[c][2 py + r][2 px + s]
      const uint32_t base = smem_u32(cur) + 4u * (uint32_t)(2 * py * ROW_LD + px);
      const uint32_t t_hi = lane_addr + COL_AHI, t_lo = lane_addr + COL_ALO;
      if (khalf == 0) {
        build_k16<0>(base, t_hi, t_lo);  build_k16<16>(base, t_hi, t_lo); build_k16<32>(base, t_hi, t_lo);
        build_k16<48>(base, t_hi, t_lo); build_k16<64>(base, t_hi, t_lo);
      } else {
        build_k16<80>(base, t_hi, t_lo);  build_k16<96>(base, t_hi, t_lo); build_k16<112>(base, t_hi, t_lo);
        build_k16<128>(base, t_hi, t_lo); build_k16<144>(base, t_hi, t_lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
      // ---- epilogue: accumulator row of this pixel, 32 channels per thread
      mbar_wait_sleep(mma_done, (uint32_t)it & 1u);
      tc_fence_after();
      uint32_t t[32];
      tmem_ld_32x32b_x32(lane_addr + COL_ACC + (uint32_t)khalf * 32u, t);
      tmem_ld_wait();
      tc_fence_before();      // the next tile's MMAs (ordered behind our a_full arrival) overwrite these columns
      const int b = tile / tiles_per_img, tr = tile % tiles_per_img;
      const int oy = (tr / a.tiles_x) * PH + py, ox = (tr % a.tiles_x) * PW + px;
      if (oy < a.Ho && ox < a.Wo) {
        const int64_t off = (((int64_t)b * a.Ho + oy) * a.Wo + ox) * a.ld_o + khalf * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(a.shift + khalf * 32 + g * 8));
          const float4 s1 = __ldg(reinterpret_cast<const float4*>(a.shift + khalf * 32 + g * 8 + 4));
          const float sh[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float v0 = fmaxf(__uint_as_float(t[g * 8 + 2 * j]) + sh[2 * j], 0.0f);
            const float v1 = fmaxf(__uint_as_float(t[g * 8 + 2 * j + 1]) + sh[2 * j + 1], 0.0f);
            split_bf16_pair(v0, v1, hw[j], lw[j]);
          }
          reinterpret_cast<uint4*>(a.o_hi + off)[g] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          reinterpret_cast<uint4*>(a.o_lo + off)[g] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
    }
    cp_async_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
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
  const size_t smem = 2 * W_PLANE_BYTES + 2 * PATCH_FLOATS * sizeof(float) + 64 + 1024;
  static DevAttrCache cfg = {};
  CDETR_CHECK_CUDA(cdetr_ensure_smem(stem_conv_kernel, smem, &cfg));
  int num_sms = 0;
  CDETR_CHECK_CUDA(cdetr_num_sms(&num_sms));
  const int grid = a.ntiles < 2 * num_sms ? a.ntiles : 2 * num_sms;
  stem_conv_kernel<<<grid, 288, smem, reinterpret_cast<cudaStream_t>(s)>>>(a);
  CDETR_CHECK_LAUNCH();
  return 0;
}
