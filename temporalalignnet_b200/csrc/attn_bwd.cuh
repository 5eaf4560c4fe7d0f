// Argument block of the attention backward (entry point in backward.cu, kernels in attention_bwd_tc.cu).
#pragma once
#include "common.cuh"

namespace tanb {

struct AttnBwdArgs {
  const bf16* q; int64_t ldq;
  const bf16* k; int64_t ldk;
  const bf16* v; int64_t ldv;
  const bf16* o; int64_t ldo;
  const bf16* dO; int64_t lddo;
  const uint8_t* kpm;
  bf16* dq; int64_t lddq;
  bf16* dk; int64_t lddk;
  bf16* dv; int64_t lddv;
  float* lse;      // [B, H, Lq]
  float* delta;    // [B, H, Lq]
  int B, H, Lq, Lk;
};

// attention_bwd_tc.cu: tcgen05 kernels; lse is an INPUT (stored by the forward kernel, log2 domain, pitch pad64(Lq)),
// delta a workspace of the same shape
int attention_bwd_tc(const AttnBwdArgs& a, cudaStream_t stream);

}  // namespace tanb
