// Hungarian set matching on the device: per-image cost block + exact linear-sum-assignment.
//
//  cost (A2/models/matcher.py:221-242; box ops A2/util/box_ops.py:17-67): fp32, same operation order
//  as the reference (explicit __f*_rn intrinsics so nvcc cannot contract into FMAs); only the diagonal
//  [Q x T_b] block of each image is computed (the reference builds the full cross-batch matrix).
//
//  assignment (A2/models/matcher.py:243-247 -> scipy.optimize.linear_sum_assignment): one CTA per image
//  runs scipy's shortest-augmenting-path solver (Crouse 2016) in fp64 with scipy's exact scan order,
//  tie rule and output ordering (C restatement in oracle/lsap.c), so the indices are
//  bit-identical to the reference's host call while removing the per-step D2H copy + host solve.
//  The column scan of one step is spread over the CTA's threads and the arg-min is a lexicographic
//  reduction on (value, assigned?, scan position) which reproduces the sequential tie rule.
#include "common.cuh"
#include "../../include/cdetr.h"
#include <math.h>

namespace {

__device__ __forceinline__ float giou_rn(float ax0, float ay0, float ax1, float ay1, float bx0, float by0,
                                         float bx1, float by1) {
  const float area_a = __fmul_rn(__fsub_rn(ax1, ax0), __fsub_rn(ay1, ay0));
  const float area_b = __fmul_rn(__fsub_rn(bx1, bx0), __fsub_rn(by1, by0));
  const float iw = fmaxf(__fsub_rn(fminf(ax1, bx1), fmaxf(ax0, bx0)), 0.0f);
  const float ih = fmaxf(__fsub_rn(fminf(ay1, by1), fmaxf(ay0, by0)), 0.0f);
  const float inter = __fmul_rn(iw, ih);
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  const float iou = __fdiv_rn(inter, uni);
  const float cw = fmaxf(__fsub_rn(fmaxf(ax1, bx1), fminf(ax0, bx0)), 0.0f);
  const float ch = fmaxf(__fsub_rn(fmaxf(ay1, by1), fminf(ay0, by0)), 0.0f);
  const float hull = __fmul_rn(cw, ch);
  return __fsub_rn(iou, __fdiv_rn(__fsub_rn(hull, uni), hull));
}

// cost[b] is [T_b, Q] (transposed, when T_b < Q: rows of the LSAP = targets) or [Q, T_b] otherwise;
// each image owns a slab of Q*Tmax floats.
__global__ void match_cost_kernel(const float* __restrict__ logits, int num_logits,
                                  const float* __restrict__ boxes, const float* __restrict__ tgt,
                                  const int* __restrict__ tgt_off, int Q, int Tmax, float w_class,
                                  float w_bbox, float w_giou, float* __restrict__ cost) {
  const int b = blockIdx.y;
  const int t0 = tgt_off[b], T = tgt_off[b + 1] - t0;
  float* cb = cost + (int64_t)b * Q * Tmax;
  const bool transposed = T < Q;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Q * T; i += gridDim.x * blockDim.x) {
    // iterate with q fastest so that both the box reads and (transposed) cost writes coalesce
    const int q = i % Q, t = i / Q;
    const float x = logits[((int64_t)b * Q + q) * num_logits];
    const float p = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
    const float neg = __fmul_rn(__fmul_rn(0.75f, __fmul_rn(p, p)),
                                -logf(__fadd_rn(__fsub_rn(1.0f, p), 1e-8f)));
    const float omp = __fsub_rn(1.0f, p);
    const float pos = __fmul_rn(__fmul_rn(0.25f, __fmul_rn(omp, omp)), -logf(__fadd_rn(p, 1e-8f)));
    const float c_class = __fsub_rn(pos, neg);
    const float4 pb = *reinterpret_cast<const float4*>(boxes + ((int64_t)b * Q + q) * 4);
    const float4 tb = *reinterpret_cast<const float4*>(tgt + (int64_t)(t0 + t) * 4);
    float c_bbox = fabsf(__fsub_rn(pb.x, tb.x));
    c_bbox = __fadd_rn(c_bbox, fabsf(__fsub_rn(pb.y, tb.y)));
    c_bbox = __fadd_rn(c_bbox, fabsf(__fsub_rn(pb.z, tb.z)));
    c_bbox = __fadd_rn(c_bbox, fabsf(__fsub_rn(pb.w, tb.w)));
    const float hwp = __fmul_rn(0.5f, pb.z), hhp = __fmul_rn(0.5f, pb.w);
    const float hwt = __fmul_rn(0.5f, tb.z), hht = __fmul_rn(0.5f, tb.w);
    const float g = giou_rn(__fsub_rn(pb.x, hwp), __fsub_rn(pb.y, hhp), __fadd_rn(pb.x, hwp),
                            __fadd_rn(pb.y, hhp), __fsub_rn(tb.x, hwt), __fsub_rn(tb.y, hht),
                            __fadd_rn(tb.x, hwt), __fadd_rn(tb.y, hht));
    const float c = __fadd_rn(__fadd_rn(__fmul_rn(w_bbox, c_bbox), __fmul_rn(w_class, c_class)),
                              __fmul_rn(w_giou, -g));
    if (transposed) cb[(int64_t)t * Q + q] = c; else cb[(int64_t)q * T + t] = c;
  }
}

// ---------------------------------------------------------------------------------------------
struct Key {  // lexicographic (val, pri, pos): smaller wins
  double val;
  int pri;  // 0: column unassigned (preferred among ties), 1: assigned
  int pos;  // unassigned: -scan position (so the LAST wins); assigned: +scan position (FIRST wins)
};
__device__ __forceinline__ bool key_less(const Key& a, const Key& b) {
  if (a.val != b.val) return a.val < b.val;
  if (a.pri != b.pri) return a.pri < b.pri;
  return a.pos < b.pos;
}
__device__ __forceinline__ Key key_shfl_xor(const Key& k, int o) {
  Key r;
  r.val = __shfl_xor_sync(0xffffffffu, k.val, o);
  r.pri = __shfl_xor_sync(0xffffffffu, k.pri, o);
  r.pos = __shfl_xor_sync(0xffffffffu, k.pos, o);
  return r;
}

template <bool ONE_WARP>
__device__ __forceinline__ void bsync() {
  if (ONE_WARP) __syncwarp(); else __syncthreads();
}

// One CTA per image.  Shared layout (nr <= nc): u[nr] v[nc] spc[nc] (double) | path col4row row4col
// remaining (int) | SR[nr] SC[nc] (uint8) | reduction scratch.
template <bool ONE_WARP>
__global__ void lsap_kernel(const float* __restrict__ cost_all, const int* __restrict__ tgt_off, int Q,
                            int Tmax, int64_t* __restrict__ out_q, int64_t* __restrict__ out_t,
                            int* __restrict__ out_n, int* __restrict__ status, int stage_cost, int stage_off) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int b = blockIdx.x;
  const int T = tgt_off[b + 1] - tgt_off[b];
  const bool transposed = T < Q;  // rows = targets (scipy transposes iff nr > nc)
  const int nr = transposed ? T : Q, nc = transposed ? Q : T;
  const int ncap = max(Q, Tmax);
  const float* cost = cost_all + (int64_t)b * Q * Tmax;  // [nr, nc] row-major
  // small problems (300 x 50: 60 KB) keep the whole cost slab in shared memory: every augmentation step re-reads one
  // row, and the L2 round trip of that read was the longest link of the dependent chain (3.4 us per step)
  float* cost_s = reinterpret_cast<float*>(smraw + stage_off);
  double* u = reinterpret_cast<double*>(smraw);
  double* v = u + ncap;
  double* spc = v + ncap;
  int* path = reinterpret_cast<int*>(spc + ncap);
  int* col4row = path + ncap;
  int* row4col = col4row + ncap;
  int* remaining = row4col + ncap;
  unsigned char* SR = reinterpret_cast<unsigned char*>(remaining + ncap);
  unsigned char* SC = SR + ncap;
  Key* red = reinterpret_cast<Key*>(SC + ncap + ((16 - (2 * ncap) % 16) % 16));
  __shared__ int s_i, s_sink, s_nrem, s_fail;
  __shared__ double s_minval;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5;
  int64_t* oq = out_q + (int64_t)b * min(Q, Tmax);
  int64_t* ot = out_t + (int64_t)b * min(Q, Tmax);
  if (nr == 0) {
    if (tid == 0) out_n[b] = 0;
    return;
  }
  for (int i = tid; i < nr; i += nt) { u[i] = 0.0; col4row[i] = -1; }
  for (int j = tid; j < nc; j += nt) { v[j] = 0.0; row4col[j] = -1; }
  if (tid == 0) s_fail = 0;
  if (stage_cost) {
    const int n4 = (nr * nc) >> 2;
    if ((reinterpret_cast<uintptr_t>(cost) & 15) == 0) {
      for (int i = tid; i < n4; i += nt) reinterpret_cast<float4*>(cost_s)[i] = __ldg(reinterpret_cast<const float4*>(cost) + i);
      for (int i = 4 * n4 + tid; i < nr * nc; i += nt) cost_s[i] = __ldg(cost + i);
    } else {
      for (int i = tid; i < nr * nc; i += nt) cost_s[i] = __ldg(cost + i);
    }
  }
  bsync<ONE_WARP>();

  for (int cur = 0; cur < nr; ++cur) {
    for (int j = tid; j < nc; j += nt) {
      remaining[j] = nc - j - 1;  // reverse fill, as scipy
      spc[j] = INFINITY;
      SC[j] = 0;
    }
    for (int i = tid; i < nr; i += nt) SR[i] = 0;
    if (tid == 0) { s_i = cur; s_sink = -1; s_nrem = nc; s_minval = 0.0; }
    bsync<ONE_WARP>();
    while (true) {
      const int i = s_i;
      const int nrem = s_nrem;
      const double minval = s_minval;
      const double ui = u[i];
      const float* crow = stage_cost ? cost_s + (int64_t)i * nc : cost + (int64_t)i * nc;
      Key best;
      best.val = INFINITY; best.pri = 2; best.pos = 0x7fffffff;
      for (int it = tid; it < nrem; it += nt) {
        const int j = remaining[it];
        const double r = __dsub_rn(__dsub_rn(__dadd_rn(minval, (double)crow[j]), ui), v[j]);
        double sj = spc[j];
        if (r < sj) {
          sj = r;
          spc[j] = r;
          path[j] = i;
        }
        Key k;
        k.val = sj;
        const bool unassigned = row4col[j] == -1;
        k.pri = unassigned ? 0 : 1;
        k.pos = unassigned ? -it : it;
        if (key_less(k, best)) best = k;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const Key other = key_shfl_xor(best, o);
        if (key_less(other, best)) best = other;
      }
      if (!ONE_WARP) {
        if (lane == 0) red[warp] = best;
        __syncthreads();
        if (warp == 0) {
          Key k2;
          k2.val = INFINITY; k2.pri = 2; k2.pos = 0x7fffffff;
          if (lane < (nt >> 5)) k2 = red[lane];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const Key other = key_shfl_xor(k2, o);
            if (key_less(other, k2)) k2 = other;
          }
          best = k2;
        }
      }
      if (tid == 0) {
        SR[i] = 1;
        if (best.val == INFINITY) {  // infeasible (cannot happen for finite costs)
          s_fail = 1;
          s_sink = -2;
        } else {
          const int index = best.pri == 0 ? -best.pos : best.pos;
          const int j = remaining[index];
          s_minval = best.val;
          if (row4col[j] == -1) s_sink = j; else s_i = row4col[j];
          SC[j] = 1;
          remaining[index] = remaining[nrem - 1];
          s_nrem = nrem - 1;
        }
      }
      bsync<ONE_WARP>();
      if (s_sink != -1) break;
    }
    if (s_fail) break;
    // dual update (scipy order of arithmetic), then augment along the alternating path
    const double minval = s_minval;
    for (int i = tid; i < nr; i += nt) {
      if (i == cur) u[i] = __dadd_rn(u[i], minval);
      else if (SR[i]) u[i] = __dadd_rn(u[i], __dsub_rn(minval, spc[col4row[i]]));
    }
    for (int j = tid; j < nc; j += nt)
      if (SC[j]) v[j] = __dsub_rn(v[j], __dsub_rn(minval, spc[j]));
    bsync<ONE_WARP>();
    if (tid == 0) {
      int j = s_sink;
      while (true) {
        const int i = path[j];
        row4col[j] = i;
        const int tmp = col4row[i];
        col4row[i] = j;
        j = tmp;
        if (i == cur) break;
      }
    }
    bsync<ONE_WARP>();
  }
  if (s_fail) {
    if (tid == 0) { out_n[b] = 0; atomicExch(status, 1); }
    return;
  }
  // output: (query, target) pairs with query ascending (scipy ordering, incl. the transposed case)
  if (!transposed) {
    for (int i = tid; i < nr; i += nt) { oq[i] = i; ot[i] = col4row[i]; }
  } else {
    for (int t = tid; t < nr; t += nt) {
      const int q = col4row[t];
      int rank = 0;
      for (int t2 = 0; t2 < nr; ++t2) rank += col4row[t2] < q;
      oq[rank] = q;
      ot[rank] = t;
    }
  }
  if (tid == 0) out_n[b] = nr;
}

// Single-warp solver with the per-column state in REGISTERS (problems up to 384 columns: lane l owns columns l, l + 32, ...).
// Same algorithm, same arithmetic order and the same tie rule as lsap_kernel (= scipy's _lsap): the scan position of a
// column in scipy's `remaining` array is tracked per column (`pos`, updated for the one column the swap-with-last
// removal moves), so the lexicographic (value, assigned?, +-position) arg-min sees exactly scipy's keys.  What changes is
// where the state lives: lsap_kernel re-reads remaining / cost / v / spc / row4col from shared memory in a dependent
// chain per column and serialises the bookkeeping of every step on lane 0 (1.7 us per augmentation step at 300 x 50:
// 188 us per call, all of it on the critical path between the forward and the backward); here a step is ten register
// updates per lane, one warp arg-min and three uniform shared-memory reads.
constexpr int LSAP_CPL_MAX = 12;   // columns per lane (template parameter: the unrolled per-step work scales with it)
template <int LSAP_CPL>
__global__ void __launch_bounds__(128)
lsap_warp_kernel(const float* __restrict__ cost_all, const int* __restrict__ tgt_off, int Q, int Tmax,
                 int64_t* __restrict__ out_q, int64_t* __restrict__ out_t, int* __restrict__ out_n,
                 int* __restrict__ status, int stage_cost, int stage_off) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int b = blockIdx.x;
  const int T = tgt_off[b + 1] - tgt_off[b];
  const bool transposed = T < Q;  // rows = targets (scipy transposes iff nr > nc)
  const int nr = transposed ? T : Q, nc = transposed ? Q : T;
  const int ncap = max(Q, Tmax);
  const float* cost = cost_all + (int64_t)b * Q * Tmax;  // [nr, nc] row-major
  float* cost_s = reinterpret_cast<float*>(smraw + stage_off);
  double* u = reinterpret_cast<double*>(smraw);
  int* path = reinterpret_cast<int*>(u + ncap);
  int* col4row = path + ncap;
  int* row4col = col4row + ncap;
  int* remaining = row4col + ncap;
  const int lane = threadIdx.x & 31;
  const int tid = threadIdx.x, nt = blockDim.x;   // warps 1.. only help to stage the cost slab
  int64_t* oq = out_q + (int64_t)b * min(Q, Tmax);
  int64_t* ot = out_t + (int64_t)b * min(Q, Tmax);
  if (nr == 0) {
    if (tid == 0) out_n[b] = 0;
    return;
  }
  for (int i = tid; i < nr; i += nt) { u[i] = 0.0; col4row[i] = -1; }
  for (int j = tid; j < nc; j += nt) row4col[j] = -1;
  if (stage_cost) {
    const int n4 = (nr * nc) >> 2;
    if ((reinterpret_cast<uintptr_t>(cost) & 15) == 0) {
      for (int i = tid; i < n4; i += nt) reinterpret_cast<float4*>(cost_s)[i] = __ldg(reinterpret_cast<const float4*>(cost) + i);
      for (int i = 4 * n4 + tid; i < nr * nc; i += nt) cost_s[i] = __ldg(cost + i);
    } else {
      for (int i = tid; i < nr * nc; i += nt) cost_s[i] = __ldg(cost + i);
    }
  }
  __syncthreads();
  if (tid >= 32) return;
  const float* cbase = stage_cost ? cost_s : cost;
  double v[LSAP_CPL], spc[LSAP_CPL];
  int pos[LSAP_CPL], r4c[LSAP_CPL];
#pragma unroll
  for (int k = 0; k < LSAP_CPL; ++k) { v[k] = 0.0; r4c[k] = -1; }
  __syncwarp();
  bool fail = false;

  for (int cur = 0; cur < nr && !fail; ++cur) {
    uint32_t open = 0;   // bit k: column lane + 32 k is still in `remaining`
#pragma unroll
    for (int k = 0; k < LSAP_CPL; ++k) {
      const int j = lane + 32 * k;
      if (j < nc) {
        open |= 1u << k;
        remaining[nc - j - 1] = j;   // reverse fill, as scipy: remaining[it] = nc - it - 1
        pos[k] = nc - j - 1;
      }
      spc[k] = INFINITY;
    }
    __syncwarp();
    int i = cur, sink = -1, nrem = nc;
    double minval = 0.0;
    while (true) {
      const double ui = u[i];
      const float* crow = cbase + (int64_t)i * nc;
      // (1) relax the open columns (branch-free) and take the lane's minimum shortest-path cost
      double m = INFINITY;
#pragma unroll
      for (int k = 0; k < LSAP_CPL; ++k) {
        const int j = lane + 32 * k;
        const bool op = (open >> k) & 1u;
        const double r = __dsub_rn(__dsub_rn(__dadd_rn(minval, (double)crow[op ? j : 0]), ui), v[k]);
        if (op && r < spc[k]) {
          spc[k] = r;
          path[j] = i;
        }
        const double val = op ? spc[k] : INFINITY;
        if (val < m) m = val;
      }
      // (2) warp minimum of a double through two 32-bit REDUX operations on its order-preserving bit pattern
      //     (+0.0 first: -0.0 and +0.0 compare equal and must map to one pattern)
      unsigned long long bits = (unsigned long long)__double_as_longlong(__dadd_rn(m, 0.0));
      bits = (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
      const unsigned hmin = __reduce_min_sync(0xffffffffu, (unsigned)(bits >> 32));
      const unsigned lmin = __reduce_min_sync(0xffffffffu, (unsigned)(bits >> 32) == hmin ? (unsigned)bits : 0xffffffffu);
      unsigned long long gb = ((unsigned long long)hmin << 32) | lmin;
      gb = (gb >> 63) ? (gb & 0x7fffffffffffffffull) : ~gb;
      const double gmin = __longlong_as_double((long long)gb);
      if (gmin == INFINITY) {  // infeasible (cannot happen for finite costs)
        fail = true;
        break;
      }
      // (3) scipy's tie rule among the columns at the minimum: unassigned columns first and of those the LAST in scan
      //     order, else the first assigned one -- one integer per column, scan positions are unique
      int tb = 0x7fffffff, bj = -1;
#pragma unroll
      for (int k = 0; k < LSAP_CPL; ++k) {
        const bool cand = ((open >> k) & 1u) && spc[k] == gmin;
        const int t = r4c[k] == -1 ? -pos[k] - (1 << 20) : pos[k];
        if (cand && t < tb) { tb = t; bj = lane + 32 * k; }
      }
      const int tbmin = __reduce_min_sync(0xffffffffu, tb);
      const int j = __reduce_max_sync(0xffffffffu, tb == tbmin ? bj : -1);
      const int index = tbmin < 0 ? -(tbmin + (1 << 20)) : tbmin;   // scan position of the chosen column
      minval = gmin;
      const int j_last = remaining[nrem - 1];
      const int r_j = row4col[j];
      __syncwarp();
      if (lane == 0) remaining[index] = j_last;
      // owner lanes: the chosen column leaves `remaining`, the former last column takes its scan position
#pragma unroll
      for (int k = 0; k < LSAP_CPL; ++k) {
        if (lane + 32 * k == j) open &= ~(1u << k);
        if (lane + 32 * k == j_last) pos[k] = index;
      }
      --nrem;
      __syncwarp();
      if (r_j == -1) { sink = j; break; }
      i = r_j;
    }
    if (fail) break;
    // dual update (scipy order of arithmetic): visited columns are exactly those that left `remaining`; the rows scanned
    // besides `cur` are the rows matched to the visited assigned columns
    if (lane == 0) u[cur] = __dadd_rn(u[cur], minval);
#pragma unroll
    for (int k = 0; k < LSAP_CPL; ++k) {
      const int j = lane + 32 * k;
      if (j < nc && !(open & (1u << k))) {
        const double d = __dsub_rn(minval, spc[k]);
        if (r4c[k] != -1) u[r4c[k]] = __dadd_rn(u[r4c[k]], d);
        v[k] = __dsub_rn(v[k], d);
      }
    }
    __syncwarp();
    if (lane == 0) {   // augment along the alternating path
      int j = sink;
      while (true) {
        const int i2 = path[j];
        row4col[j] = i2;
        const int tmp = col4row[i2];
        col4row[i2] = j;
        j = tmp;
        if (i2 == cur) break;
      }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < LSAP_CPL; ++k) {
      const int j = lane + 32 * k;
      if (j < nc) r4c[k] = row4col[j];
    }
  }
  if (fail) {
    if (lane == 0) { out_n[b] = 0; atomicExch(status, 1); }
    return;
  }
  // output: (query, target) pairs with query ascending (scipy ordering, incl. the transposed case)
  if (!transposed) {
    for (int i = lane; i < nr; i += 32) { oq[i] = i; ot[i] = col4row[i]; }
  } else {
    for (int t = lane; t < nr; t += 32) {
      const int q = col4row[t];
      int rank = 0;
      for (int t2 = 0; t2 < nr; ++t2) rank += col4row[t2] < q;
      oq[rank] = q;
      ot[rank] = t;
    }
  }
  if (lane == 0) out_n[b] = nr;
}

// The same register-resident formulation for larger problems (up to 4096 columns): a whole CTA, thread t owns columns
// t, t + blockDim, ...; the arg-min is a warp REDUX followed by a 32-entry cross-warp step that EVERY warp repeats from
// shared memory (no broadcast barrier), three CTA barriers per augmentation step.  The cost row of the next step is read
// from global memory / L2 (the matrix does not fit shared memory), which is now the longest link of the chain.
template <int CPL>
__global__ void __launch_bounds__(1024)
lsap_block_kernel(const float* __restrict__ cost_all, const int* __restrict__ tgt_off, int Q, int Tmax,
                  int64_t* __restrict__ out_q, int64_t* __restrict__ out_t, int* __restrict__ out_n,
                  int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ unsigned long long red64[32];
  __shared__ int red_tb[32];
  __shared__ int s_j, s_rj;
  const int b = blockIdx.x;
  const int T = tgt_off[b + 1] - tgt_off[b];
  const bool transposed = T < Q;
  const int nr = transposed ? T : Q, nc = transposed ? Q : T;
  const int ncap = max(Q, Tmax);
  const float* cost = cost_all + (int64_t)b * Q * Tmax;  // [nr, nc] row-major
  double* u = reinterpret_cast<double*>(smraw);
  int* path = reinterpret_cast<int*>(u + ncap);
  int* col4row = path + ncap;
  int* row4col = col4row + ncap;
  int* remaining = row4col + ncap;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  int64_t* oq = out_q + (int64_t)b * min(Q, Tmax);
  int64_t* ot = out_t + (int64_t)b * min(Q, Tmax);
  if (nr == 0) {
    if (tid == 0) out_n[b] = 0;
    return;
  }
  for (int i = tid; i < nr; i += nt) { u[i] = 0.0; col4row[i] = -1; }
  for (int j = tid; j < nc; j += nt) row4col[j] = -1;
  double v[CPL], spc[CPL];
  int pos[CPL], r4c[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) { v[k] = 0.0; r4c[k] = -1; }
  __syncthreads();
  bool fail = false;

  for (int cur = 0; cur < nr && !fail; ++cur) {
    uint32_t open = 0;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int j = tid + nt * k;
      if (j < nc) {
        open |= 1u << k;
        remaining[nc - j - 1] = j;
        pos[k] = nc - j - 1;
      }
      spc[k] = INFINITY;
    }
    __syncthreads();
    int i = cur, sink = -1, nrem = nc;
    double minval = 0.0;
    while (true) {
      const double ui = u[i];
      const float* crow = cost + (int64_t)i * nc;
      double m = INFINITY;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int j = tid + nt * k;
        const bool op = (open >> k) & 1u;
        const double r = __dsub_rn(__dsub_rn(__dadd_rn(minval, (double)__ldg(crow + (op ? j : 0))), ui), v[k]);
        if (op && r < spc[k]) {
          spc[k] = r;
          path[j] = i;
        }
        const double val = op ? spc[k] : INFINITY;
        if (val < m) m = val;
      }
      // CTA minimum: warp REDUX on the order-preserving bit pattern, one entry per warp, every warp reduces the 32 entries
      unsigned long long bits = (unsigned long long)__double_as_longlong(__dadd_rn(m, 0.0));
      bits = (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
      {
        const unsigned hmin = __reduce_min_sync(0xffffffffu, (unsigned)(bits >> 32));
        const unsigned lmin = __reduce_min_sync(0xffffffffu, (unsigned)(bits >> 32) == hmin ? (unsigned)bits : 0xffffffffu);
        if (lane == 0) red64[warp] = ((unsigned long long)hmin << 32) | lmin;
      }
      __syncthreads();
      unsigned long long wb = lane < nw ? red64[lane] : ~0ull;
      const unsigned hmin = __reduce_min_sync(0xffffffffu, (unsigned)(wb >> 32));
      const unsigned lmin = __reduce_min_sync(0xffffffffu, (unsigned)(wb >> 32) == hmin ? (unsigned)wb : 0xffffffffu);
      unsigned long long gb = ((unsigned long long)hmin << 32) | lmin;
      gb = (gb >> 63) ? (gb & 0x7fffffffffffffffull) : ~gb;
      const double gmin = __longlong_as_double((long long)gb);
      if (gmin == INFINITY) {  // infeasible (cannot happen for finite costs); uniform over the CTA
        fail = true;
        break;
      }
      int tb = 0x7fffffff, bj = -1, brj = -1;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const bool cand = ((open >> k) & 1u) && spc[k] == gmin;
        const int t = r4c[k] == -1 ? -pos[k] - (1 << 20) : pos[k];
        if (cand && t < tb) { tb = t; bj = tid + nt * k; brj = r4c[k]; }
      }
      {
        const int wtb = __reduce_min_sync(0xffffffffu, tb);
        if (lane == 0) red_tb[warp] = wtb;
      }
      __syncthreads();
      const int tbmin = __reduce_min_sync(0xffffffffu, lane < nw ? red_tb[lane] : 0x7fffffff);
      if (tb == tbmin) { s_j = bj; s_rj = brj; }     // scan positions are unique: exactly one thread
      const int j_last = remaining[nrem - 1];
      __syncthreads();
      const int j = s_j, r_j = s_rj;
      const int index = tbmin < 0 ? -(tbmin + (1 << 20)) : tbmin;
      minval = gmin;
      if (tid == 0) remaining[index] = j_last;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        if (tid + nt * k == j) open &= ~(1u << k);
        if (tid + nt * k == j_last) pos[k] = index;
      }
      --nrem;
      if (r_j == -1) { sink = j; break; }
      i = r_j;
    }
    if (fail) break;
    __syncthreads();
    if (tid == 0) u[cur] = __dadd_rn(u[cur], minval);
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int j = tid + nt * k;
      if (j < nc && !((open >> k) & 1u)) {
        const double d = __dsub_rn(minval, spc[k]);
        if (r4c[k] != -1) u[r4c[k]] = __dadd_rn(u[r4c[k]], d);
        v[k] = __dsub_rn(v[k], d);
      }
    }
    __syncthreads();
    if (tid == 0) {   // augment along the alternating path
      int j = sink;
      while (true) {
        const int i2 = path[j];
        row4col[j] = i2;
        const int tmp = col4row[i2];
        col4row[i2] = j;
        j = tmp;
        if (i2 == cur) break;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int j = tid + nt * k;
      if (j < nc) r4c[k] = row4col[j];
    }
  }
  if (fail) {
    if (tid == 0) { out_n[b] = 0; atomicExch(status, 1); }
    return;
  }
  if (!transposed) {
    for (int i = tid; i < nr; i += nt) { oq[i] = i; ot[i] = col4row[i]; }
  } else {
    for (int t = tid; t < nr; t += nt) {
      const int q = col4row[t];
      int rank = 0;
      for (int t2 = 0; t2 < nr; ++t2) rank += col4row[t2] < q;
      oq[rank] = q;
      ot[rank] = t;
    }
  }
  if (tid == 0) out_n[b] = nr;
}

size_t lsap_warp_smem(int ncap) { return (sizeof(double) + 4 * sizeof(int)) * (size_t)ncap + 16; }

size_t lsap_smem(int ncap, int nthreads) {
  size_t s = 3 * sizeof(double) * ncap + 4 * sizeof(int) * ncap + 2 * (size_t)ncap;
  s += (16 - (2 * ncap) % 16) % 16;
  s += sizeof(Key) * 32;
  (void)nthreads;
  return s + 16;
}

}  // namespace

extern "C" int cdetr_match_cost(const float* logits, int num_logits, const float* boxes, const float* tgt_boxes,
                                const int* tgt_off, int B, int Q, int Tmax, float w_class, float w_bbox,
                                float w_giou, float* cost, cdetr_stream_t s) {
  CDETR_CHECK_ARG(logits && boxes && tgt_boxes && tgt_off && cost && B > 0 && Q > 0 && Tmax >= 0,
                  "match_cost: bad args");
  if (Tmax == 0) return 0;
  dim3 grid(cdiv((int64_t)Q * Tmax, 256), B);
  if (grid.x > 64) grid.x = 64;
  match_cost_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(s)>>>(logits, num_logits, boxes, tgt_boxes,
                                                                       tgt_off, Q, Tmax, w_class, w_bbox,
                                                                       w_giou, cost);
  CDETR_CHECK_LAUNCH();
  return 0;
}

// out_q/out_t: [B, min(Q,Tmax)] int64 (first out_n[b] entries valid); status: device int set to 1 on failure.
extern "C" int cdetr_lsap(const float* cost, const int* tgt_off, int B, int Q, int Tmax, int64_t* out_q,
                          int64_t* out_t, int* out_n, int* status, cdetr_stream_t s_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(s_);
  CDETR_CHECK_ARG(cost && tgt_off && out_q && out_t && out_n && status && B > 0 && Q > 0, "lsap: bad args");
  const int ncap = Q > Tmax ? Q : Tmax;
  const size_t base = (lsap_smem(ncap, 32) + 15) / 16 * 16;
  const size_t slab = (size_t)Q * Tmax * sizeof(float);
  const int stage_cost = base + slab <= 200 * 1024 ? 1 : 0;
  static const bool legacy = getenv("CDETR_LSAP_LEGACY") != nullptr;     // A/B: shared-memory-state single-warp kernel
  if (ncap <= 32 * LSAP_CPL_MAX && !legacy) {
    const size_t wbase = (lsap_warp_smem(ncap) + 15) / 16 * 16;
    const int stage = wbase + slab <= 200 * 1024 ? 1 : 0;
    const size_t smem = stage ? wbase + slab : wbase;
    const int cpl = (ncap + 31) / 32;
#define CDETR_LSAP_LAUNCH(N)                                                                                              \
  do {                                                                                                                    \
    static DevAttrCache cfg = {};                                                                                         \
    CDETR_CHECK_CUDA(cdetr_ensure_smem(lsap_warp_kernel<N>, 200 * 1024, &cfg));                                           \
    lsap_warp_kernel<N><<<B, stage ? 128 : 32, smem, s>>>(cost, tgt_off, Q, Tmax, out_q, out_t, out_n, status, stage,     \
                                                          (int)wbase);                                                    \
  } while (0)
    if (cpl <= 2) CDETR_LSAP_LAUNCH(2);
    else if (cpl <= 4) CDETR_LSAP_LAUNCH(4);
    else if (cpl <= 7) CDETR_LSAP_LAUNCH(7);
    else if (cpl <= 10) CDETR_LSAP_LAUNCH(10);
    else CDETR_LSAP_LAUNCH(12);
#undef CDETR_LSAP_LAUNCH
  } else if (ncap <= 384) {
    const size_t smem = stage_cost ? base + slab : lsap_smem(ncap, 32);
    { static DevAttrCache cfg = {}; CDETR_CHECK_CUDA(cdetr_ensure_smem(lsap_kernel<true>, 200 * 1024, &cfg)); }
    lsap_kernel<true><<<B, 32, smem, s>>>(cost, tgt_off, Q, Tmax, out_q, out_t, out_n, status, stage_cost, (int)base);
  } else if (ncap <= 4096 && !legacy) {
    // register-resident CTA solver: one column per thread up to 1024 columns, then 2 / 4 per thread
    const int cpl = ncap <= 1024 ? 1 : (ncap <= 2048 ? 2 : 4);
    int nt = ((ncap + cpl - 1) / cpl + 31) / 32 * 32;
    if (nt < 64) nt = 64;
    const size_t smem = lsap_warp_smem(ncap);
#define CDETR_LSAP_BLOCK(N)                                                                              \
  do {                                                                                                   \
    static DevAttrCache cfg = {};                                                                        \
    CDETR_CHECK_CUDA(cdetr_ensure_smem(lsap_block_kernel<N>, 128 * 1024, &cfg));                         \
    lsap_block_kernel<N><<<B, nt, smem, s>>>(cost, tgt_off, Q, Tmax, out_q, out_t, out_n, status);       \
  } while (0)
    if (cpl == 1) CDETR_LSAP_BLOCK(1);
    else if (cpl == 2) CDETR_LSAP_BLOCK(2);
    else CDETR_LSAP_BLOCK(4);
#undef CDETR_LSAP_BLOCK
  } else {
    int nt = 256;
    if (ncap > 768) nt = 512;
    if (ncap > 2048) nt = 1024;
    const size_t smem = lsap_smem(ncap, nt);
    CDETR_CHECK_ARG(smem <= 200 * 1024, "lsap: problem too large for shared memory (n=%d)", ncap);
    { static DevAttrCache cfg = {}; CDETR_CHECK_CUDA(cdetr_ensure_smem(lsap_kernel<false>, 200 * 1024, &cfg)); }
    lsap_kernel<false><<<B, nt, smem, s>>>(cost, tgt_off, Q, Tmax, out_q, out_t, out_n, status, 0, 0);
  }
  CDETR_CHECK_LAUNCH();
  return 0;
}
