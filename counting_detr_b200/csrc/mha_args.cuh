// Argument block shared by the decoder self-attention kernels (mha.cu: CUDA cores, mha_tc.cu: tensor cores).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

struct MhaArgs {
  int B, L, E, nh;
  const float* q;  // [B,L,ldq] (projected, bias added, unscaled); q/k/v may be column slices of one buffer
  const float* k;
  const float* v;
  int64_t ldq;     // row pitch of q/k/v in floats
  __nv_bfloat16 *o_hi, *o_lo;  // [B,L,E] split
  int64_t ld_o;
  float* lse;      // [B,nh,L]
  // backward
  const float* d_o;  // [B,L,E]
  float* dsum;       // [B,nh,L]  D_i = dO_i . O_i
  __nv_bfloat16 *dq_hi, *dq_lo, *dk_hi, *dk_lo, *dv_hi, *dv_lo;  // split, row pitch ld_g
  int64_t ld_g;
};

// mha_tc.cu
int mha_tc_fits(int L);
int mha_fwd_tc_launch(const MhaArgs& a, cudaStream_t s);
int mha_bwd_tc_launch(const MhaArgs& a, cudaStream_t s);
