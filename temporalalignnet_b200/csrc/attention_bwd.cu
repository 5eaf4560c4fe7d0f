// Attention backward (head_dim 64) with register-resident score tiles: mma.sync.m16n8k16 (bf16 in, fp32
// accumulate), operands through ldmatrix, no shared-memory round trip for S / P / dS.
//
//   kernel A (CTA = 64 queries x (clip, head)):  pass 1: lse_i by an online max / sum over the key blocks;
//       pass 2: S = Q K^T, dP = dO V^T, dS = P (dP - delta) kept in the accumulator layout, repacked in registers
//       into the A operand of dQ += dS K.  Writes dq, lse, delta.
//   kernel B (CTA = 64 keys x (clip, head)): walks the query blocks with the TRANSPOSED tiles S^T = K Q^T,
//       dP^T = V dO^T so that P^T and dS^T come out of the accumulators as A operands of dV += P^T dO and
//       dK += dS^T Q.  Writes dk, dv.
// Two kernels instead of one with atomics on dq: deterministic, at the price of recomputing S and dP once more.
// 4 warps per CTA, warp w owns rows 16 w .. 16 w + 15 of the CTA's 64-row block; 64-column steps.
//
// This is the legacy warp-level tensor path (it compiles for sm_100a but does not use tcgen05 / TMEM): the
// backward pass' first optimisation step over the wmma version in backward.cu (kept as TAN_ATTN_BWD=wmma);
// the tcgen05 version (S / dP accumulators in TMEM as in attention.cu) is the follow-up.
#include "attn_bwd_mma.cuh"

namespace tanb {

namespace {

using namespace abw;

__global__ void __launch_bounds__(128) attn_bwd_dq_mma_kernel(const AttnBwdArgs a) {
  __shared__ __align__(128) Tiles sm;
  const int q0 = blockIdx.x * kBlk, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t qbase = static_cast<int64_t>(b) * a.Lq, kbase = static_cast<int64_t>(b) * a.Lk;

  load_tile(sm.q, a.q, a.ldq, q0, a.Lq, qbase, h * 64);
  load_tile(sm.dO, a.dO, a.lddo, q0, a.Lq, qbase, h * 64);
  {  // delta_i = <do_i, o_i>: two threads per row, 32 features each
    const int row = threadIdx.x >> 1, cbeg = (threadIdx.x & 1) * 32;
    float d = 0.f;
    if (q0 + row < a.Lq) {
      const bf16* po = a.o + (qbase + q0 + row) * a.ldo + h * 64 + cbeg;
      const bf16* pd = a.dO + (qbase + q0 + row) * a.lddo + h * 64 + cbeg;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 x = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(po + j));
        const float2 y = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(pd + j));
        d += x.x * y.x + x.y * y.y;
      }
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    if ((threadIdx.x & 1) == 0) sm.delta[row] = d;
  }
  __syncthreads();
  uint32_t aq[4][4], ado[4][4];
  load_a_rows(aq, sm.q, warp * 16, lane);
  load_a_rows(ado, sm.dO, warp * 16, lane);
  const int r_lo = warp * 16 + g, r_hi = r_lo + 8;
  const float delta_lo = sm.delta[r_lo], delta_hi = sm.delta[r_hi];

  // ---- pass 1: lse of the two rows this thread shares with its quad ---------------------------------
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  for (int k0 = 0; k0 < a.Lk; k0 += kBlk) {
    __syncthreads();
    load_tile(sm.k, a.k, a.ldk, k0, a.Lk, kbase, h * 64);
    fill_bias(sm.bias, a.kpm, b, a.Lk, k0);
    __syncthreads();
    float s[8][4];
    zero_acc(s);
    mm_nt(s, aq, sm.k, lane);
    float bm_lo = -INFINITY, bm_hi = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float b0 = sm.bias[8 * j + 2 * t], b1 = sm.bias[8 * j + 2 * t + 1];
      s[j][0] = s[j][0] * kScale + b0; s[j][1] = s[j][1] * kScale + b1;
      s[j][2] = s[j][2] * kScale + b0; s[j][3] = s[j][3] * kScale + b1;
      bm_lo = fmaxf(bm_lo, fmaxf(s[j][0], s[j][1]));
      bm_hi = fmaxf(bm_hi, fmaxf(s[j][2], s[j][3]));
    }
    bm_lo = quad_max(bm_lo);
    bm_hi = quad_max(bm_hi);
    const float mn_lo = fmaxf(m_lo, bm_lo), mn_hi = fmaxf(m_hi, bm_hi);
    const float ref_lo = mn_lo == -INFINITY ? 0.f : mn_lo, ref_hi = mn_hi == -INFINITY ? 0.f : mn_hi;
    float ps_lo = 0.f, ps_hi = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ps_lo += __expf(s[j][0] - ref_lo) + __expf(s[j][1] - ref_lo);
      ps_hi += __expf(s[j][2] - ref_hi) + __expf(s[j][3] - ref_hi);
    }
    ps_lo = quad_sum(ps_lo);
    ps_hi = quad_sum(ps_hi);
    l_lo = l_lo * __expf(m_lo - ref_lo) + ps_lo;
    l_hi = l_hi * __expf(m_hi - ref_hi) + ps_hi;
    m_lo = mn_lo;
    m_hi = mn_hi;
  }
  const float lse_lo = m_lo + __logf(l_lo), lse_hi = m_hi + __logf(l_hi);
  if (t == 0) {
    const int64_t base = (static_cast<int64_t>(b) * a.H + h) * a.Lq + q0;
    if (q0 + r_lo < a.Lq) { a.lse[base + r_lo] = lse_lo; a.delta[base + r_lo] = delta_lo; }
    if (q0 + r_hi < a.Lq) { a.lse[base + r_hi] = lse_hi; a.delta[base + r_hi] = delta_hi; }
  }

  // ---- pass 2: dq ------------------------------------------------------------------------------------
  float dq[8][4];
  zero_acc(dq);
  for (int k0 = 0; k0 < a.Lk; k0 += kBlk) {
    __syncthreads();
    load_tile(sm.k, a.k, a.ldk, k0, a.Lk, kbase, h * 64);
    load_tile(sm.v, a.v, a.ldv, k0, a.Lk, kbase, h * 64);
    fill_bias(sm.bias, a.kpm, b, a.Lk, k0);
    __syncthreads();
    float s[8][4], dp[8][4];
    zero_acc(s);
    mm_nt(s, aq, sm.k, lane);
    zero_acc(dp);
    mm_nt(dp, ado, sm.v, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float b0 = sm.bias[8 * j + 2 * t], b1 = sm.bias[8 * j + 2 * t + 1];
      const float p0 = __expf(s[j][0] * kScale + b0 - lse_lo), p1 = __expf(s[j][1] * kScale + b1 - lse_lo);
      const float p2 = __expf(s[j][2] * kScale + b0 - lse_hi), p3 = __expf(s[j][3] * kScale + b1 - lse_hi);
      s[j][0] = p0 * (dp[j][0] - delta_lo); s[j][1] = p1 * (dp[j][1] - delta_lo);
      s[j][2] = p2 * (dp[j][2] - delta_hi); s[j][3] = p3 * (dp[j][3] - delta_hi);
    }
    uint32_t ads[4][4];
    acc_to_a(ads, s);
    mm_nn(dq, ads, sm.k, lane);
  }
  store_acc(dq, kScale, a.dq, a.lddq, qbase, q0 + warp * 16, a.Lq, h * 64, lane);
}

__global__ void __launch_bounds__(128) attn_bwd_dkv_mma_kernel(const AttnBwdArgs a) {
  __shared__ __align__(128) Tiles sm;
  const int k0 = blockIdx.x * kBlk, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t qbase = static_cast<int64_t>(b) * a.Lq, kbase = static_cast<int64_t>(b) * a.Lk;

  load_tile(sm.k, a.k, a.ldk, k0, a.Lk, kbase, h * 64);
  load_tile(sm.v, a.v, a.ldv, k0, a.Lk, kbase, h * 64);
  fill_bias(sm.bias, a.kpm, b, a.Lk, k0);
  __syncthreads();
  uint32_t ak[4][4], av[4][4];
  load_a_rows(ak, sm.k, warp * 16, lane);
  load_a_rows(av, sm.v, warp * 16, lane);
  const float bias_lo = sm.bias[warp * 16 + g], bias_hi = sm.bias[warp * 16 + g + 8];   // this thread's two keys
  float dk[8][4], dv[8][4];
  zero_acc(dk);
  zero_acc(dv);
  for (int q0 = 0; q0 < a.Lq; q0 += kBlk) {
    __syncthreads();
    load_tile(sm.q, a.q, a.ldq, q0, a.Lq, qbase, h * 64);
    load_tile(sm.dO, a.dO, a.lddo, q0, a.Lq, qbase, h * 64);
    for (int j = threadIdx.x; j < kBlk; j += blockDim.x) {
      const bool ok = q0 + j < a.Lq;
      const int64_t idx = (static_cast<int64_t>(b) * a.H + h) * a.Lq + q0 + j;
      sm.lse[j] = ok ? a.lse[idx] : INFINITY;           // rows beyond Lq: p = exp(-inf) = 0
      sm.delta[j] = ok ? a.delta[idx] : 0.f;
    }
    __syncthreads();
    float st[8][4], dpt[8][4];                          // S^T, dP^T: rows = keys, columns = queries
    zero_acc(st);
    mm_nt(st, ak, sm.q, lane);
    zero_acc(dpt);
    mm_nt(dpt, av, sm.dO, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = 8 * j + 2 * t;
      const float l0 = sm.lse[c], l1 = sm.lse[c + 1], d0 = sm.delta[c], d1 = sm.delta[c + 1];
      const float p0 = __expf(st[j][0] * kScale + bias_lo - l0), p1 = __expf(st[j][1] * kScale + bias_lo - l1);
      const float p2 = __expf(st[j][2] * kScale + bias_hi - l0), p3 = __expf(st[j][3] * kScale + bias_hi - l1);
      st[j][0] = p0; st[j][1] = p1; st[j][2] = p2; st[j][3] = p3;
      dpt[j][0] = p0 * (dpt[j][0] - d0); dpt[j][1] = p1 * (dpt[j][1] - d1);
      dpt[j][2] = p2 * (dpt[j][2] - d0); dpt[j][3] = p3 * (dpt[j][3] - d1);
    }
    uint32_t ap[4][4];
    acc_to_a(ap, st);
    mm_nn(dv, ap, sm.dO, lane);                         // dv[key, :] += P^T do
    acc_to_a(ap, dpt);
    mm_nn(dk, ap, sm.q, lane);                          // dk[key, :] += dS^T q
  }
  store_acc(dv, 1.0f, a.dv, a.lddv, kbase, k0 + warp * 16, a.Lk, h * 64, lane);
  store_acc(dk, kScale, a.dk, a.lddk, kbase, k0 + warp * 16, a.Lk, h * 64, lane);
}

}  // namespace

int attention_bwd_mma(const AttnBwdArgs& a, cudaStream_t stream) {
  attn_bwd_dq_mma_kernel<<<dim3((a.Lq + kBlk - 1) / kBlk, a.H, a.B), 128, 0, stream>>>(a);
  TAN_CUDA(cudaGetLastError());
  attn_bwd_dkv_mma_kernel<<<dim3((a.Lk + kBlk - 1) / kBlk, a.H, a.B), 128, 0, stream>>>(a);
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}

}  // namespace tanb
