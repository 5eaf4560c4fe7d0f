// Pipelined variant of the mma.sync attention backward (attention_bwd.cu): the same arithmetic, with the K / V
// (kernel A) and Q / dO / lse / delta (kernel B) tiles of block i+1 fetched by cp.async into a second shared-memory
// buffer while block i is computed, one __syncthreads per block instead of two, and the clip's key mask staged
// once.  Motivation: the ncu capture of the synchronous version (profiles/r01g_prof_attnbwd_summary.txt) shows the
// top stalls on the STS behind the global loads and on the block barriers, tensor pipe 21-27 % active.
//
// STATUS: EXPERIMENTAL, NOT YET RUN ON A GPU (written after round 1's GPU budget was spent).  Selected only by
// TAN_ATTN_BWD=pipe; tests/test_backward_kernels_gpu.py::test_attention_bwd_pipelined_variant covers it when
// TAN_TEST_EXPERIMENTAL=1.  The default stays the validated synchronous version.
#include "attn_bwd_mma.cuh"

namespace tanb {

namespace {

using namespace abw;

constexpr int kMaxKeys = 2048;             // key-mask staging (floats); longer sequences use the synchronous version

struct TilesA {                            // kernel A: queries resident, keys stream
  bf16 q[kBlk][kPitch];
  bf16 dO[kBlk][kPitch];
  bf16 k[2][kBlk][kPitch];
  bf16 v[2][kBlk][kPitch];
  float delta[kBlk];
  float bias[kMaxKeys];                    // 0 / -inf per key of the clip
};

struct TilesB {                            // kernel B: keys resident, queries stream
  bf16 k[kBlk][kPitch];
  bf16 v[kBlk][kPitch];
  bf16 q[2][kBlk][kPitch];
  bf16 dO[2][kBlk][kPitch];
  float lse[2][kBlk];
  float delta[2][kBlk];
  float bias[kBlk];
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  const int sz = valid ? 16 : 0;           // 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// asynchronous counterpart of load_tile: rows beyond n_rows are zero-filled
__device__ __forceinline__ void prefetch_tile(bf16 (*dst)[kPitch], const bf16* src, int64_t ld, int row0, int n_rows,
                                              int64_t base_row, int col0) {
  for (int i = threadIdx.x; i < kBlk * 8; i += blockDim.x) {
    const int r = i >> 3, c8 = (i & 7) * 8;
    const bool ok = row0 + r < n_rows;
    const bf16* p = src + (base_row + (ok ? row0 + r : 0)) * ld + col0 + c8;
    cp_async16(&dst[r][c8], p, ok);
  }
}

__global__ void __launch_bounds__(128) attn_bwd_dq_pipe_kernel(const AttnBwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  TilesA& sm = *reinterpret_cast<TilesA*>(smem_raw);
  const int q0 = blockIdx.x * kBlk, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t qbase = static_cast<int64_t>(b) * a.Lq, kbase = static_cast<int64_t>(b) * a.Lk;
  const int nb = (a.Lk + kBlk - 1) / kBlk;

  load_tile(sm.q, a.q, a.ldq, q0, a.Lq, qbase, h * 64);
  load_tile(sm.dO, a.dO, a.lddo, q0, a.Lq, qbase, h * 64);
  for (int key = threadIdx.x; key < nb * kBlk; key += blockDim.x)
    sm.bias[key] = (key < a.Lk && (a.kpm == nullptr || a.kpm[static_cast<int64_t>(b) * a.Lk + key] == 0)) ? 0.f : -INFINITY;
  prefetch_tile(sm.k[0], a.k, a.ldk, 0, a.Lk, kbase, h * 64);          // pass 1, block 0
  cp_async_commit();
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

  // ---- pass 1: lse ------------------------------------------------------------------------------------
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  for (int i = 0; i < nb; ++i) {
    cp_async_wait_all();
    __syncthreads();                       // block i has landed; everybody has finished block i-1 (other buffer)
    if (i + 1 < nb) {
      prefetch_tile(sm.k[(i + 1) & 1], a.k, a.ldk, (i + 1) * kBlk, a.Lk, kbase, h * 64);
      cp_async_commit();
    }
    const float* bias = sm.bias + i * kBlk;
    float s[8][4];
    zero_acc(s);
    mm_nt(s, aq, sm.k[i & 1], lane);
    float bm_lo = -INFINITY, bm_hi = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float b0 = bias[8 * j + 2 * t], b1 = bias[8 * j + 2 * t + 1];
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

  // ---- pass 2: dq --------------------------------------------------------------------------------------
  __syncthreads();                         // every warp has left pass 1's buffers
  prefetch_tile(sm.k[0], a.k, a.ldk, 0, a.Lk, kbase, h * 64);
  prefetch_tile(sm.v[0], a.v, a.ldv, 0, a.Lk, kbase, h * 64);
  cp_async_commit();
  float dq[8][4];
  zero_acc(dq);
  for (int i = 0; i < nb; ++i) {
    cp_async_wait_all();
    __syncthreads();
    if (i + 1 < nb) {
      prefetch_tile(sm.k[(i + 1) & 1], a.k, a.ldk, (i + 1) * kBlk, a.Lk, kbase, h * 64);
      prefetch_tile(sm.v[(i + 1) & 1], a.v, a.ldv, (i + 1) * kBlk, a.Lk, kbase, h * 64);
      cp_async_commit();
    }
    const float* bias = sm.bias + i * kBlk;
    float s[8][4], dp[8][4];
    zero_acc(s);
    mm_nt(s, aq, sm.k[i & 1], lane);
    zero_acc(dp);
    mm_nt(dp, ado, sm.v[i & 1], lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float b0 = bias[8 * j + 2 * t], b1 = bias[8 * j + 2 * t + 1];
      const float p0 = __expf(s[j][0] * kScale + b0 - lse_lo), p1 = __expf(s[j][1] * kScale + b1 - lse_lo);
      const float p2 = __expf(s[j][2] * kScale + b0 - lse_hi), p3 = __expf(s[j][3] * kScale + b1 - lse_hi);
      s[j][0] = p0 * (dp[j][0] - delta_lo); s[j][1] = p1 * (dp[j][1] - delta_lo);
      s[j][2] = p2 * (dp[j][2] - delta_hi); s[j][3] = p3 * (dp[j][3] - delta_hi);
    }
    uint32_t ads[4][4];
    acc_to_a(ads, s);
    mm_nn(dq, ads, sm.k[i & 1], lane);
  }
  store_acc(dq, kScale, a.dq, a.lddq, qbase, q0 + warp * 16, a.Lq, h * 64, lane);
}

__global__ void __launch_bounds__(128) attn_bwd_dkv_pipe_kernel(const AttnBwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  TilesB& sm = *reinterpret_cast<TilesB*>(smem_raw);
  const int k0 = blockIdx.x * kBlk, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t qbase = static_cast<int64_t>(b) * a.Lq, kbase = static_cast<int64_t>(b) * a.Lk;
  const int nq = (a.Lq + kBlk - 1) / kBlk;
  const int64_t stat_base = (static_cast<int64_t>(b) * a.H + h) * a.Lq;

  // query block i -> buffer i & 1: Q, dO tiles and the 64 lse / delta values (thread j < 64: lse, else delta);
  // rows beyond Lq are zero-filled everywhere: q = do = 0 make their contributions vanish (p stays finite)
  auto prefetch_q = [&](int i) {
    const int q0 = i * kBlk, buf = i & 1;
    prefetch_tile(sm.q[buf], a.q, a.ldq, q0, a.Lq, qbase, h * 64);
    prefetch_tile(sm.dO[buf], a.dO, a.lddo, q0, a.Lq, qbase, h * 64);
    const int j = threadIdx.x & 63;
    const bool ok = q0 + j < a.Lq;
    const float* src = (threadIdx.x < 64 ? a.lse : a.delta) + stat_base + (ok ? q0 + j : 0);
    cp_async4(threadIdx.x < 64 ? &sm.lse[buf][j] : &sm.delta[buf][j], src, ok);
    cp_async_commit();
  };

  prefetch_q(0);
  load_tile(sm.k, a.k, a.ldk, k0, a.Lk, kbase, h * 64);
  load_tile(sm.v, a.v, a.ldv, k0, a.Lk, kbase, h * 64);
  fill_bias(sm.bias, a.kpm, b, a.Lk, k0);
  __syncthreads();
  uint32_t ak[4][4], av[4][4];
  load_a_rows(ak, sm.k, warp * 16, lane);
  load_a_rows(av, sm.v, warp * 16, lane);
  const float bias_lo = sm.bias[warp * 16 + g], bias_hi = sm.bias[warp * 16 + g + 8];
  float dk[8][4], dv[8][4];
  zero_acc(dk);
  zero_acc(dv);
  for (int i = 0; i < nq; ++i) {
    cp_async_wait_all();
    __syncthreads();
    if (i + 1 < nq) prefetch_q(i + 1);
    const int buf = i & 1;
    float st[8][4], dpt[8][4];                          // S^T, dP^T: rows = keys, columns = queries
    zero_acc(st);
    mm_nt(st, ak, sm.q[buf], lane);
    zero_acc(dpt);
    mm_nt(dpt, av, sm.dO[buf], lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = 8 * j + 2 * t;
      const float l0 = sm.lse[buf][c], l1 = sm.lse[buf][c + 1], d0 = sm.delta[buf][c], d1 = sm.delta[buf][c + 1];
      const float p0 = __expf(st[j][0] * kScale + bias_lo - l0), p1 = __expf(st[j][1] * kScale + bias_lo - l1);
      const float p2 = __expf(st[j][2] * kScale + bias_hi - l0), p3 = __expf(st[j][3] * kScale + bias_hi - l1);
      st[j][0] = p0; st[j][1] = p1; st[j][2] = p2; st[j][3] = p3;
      dpt[j][0] = p0 * (dpt[j][0] - d0); dpt[j][1] = p1 * (dpt[j][1] - d1);
      dpt[j][2] = p2 * (dpt[j][2] - d0); dpt[j][3] = p3 * (dpt[j][3] - d1);
    }
    uint32_t ap[4][4];
    acc_to_a(ap, st);
    mm_nn(dv, ap, sm.dO[buf], lane);
    acc_to_a(ap, dpt);
    mm_nn(dk, ap, sm.q[buf], lane);
  }
  store_acc(dv, 1.0f, a.dv, a.lddv, kbase, k0 + warp * 16, a.Lk, h * 64, lane);
  store_acc(dk, kScale, a.dk, a.lddk, kbase, k0 + warp * 16, a.Lk, h * 64, lane);
}

}  // namespace

bool attention_bwd_pipe_supported(const AttnBwdArgs& a) { return a.Lk <= kMaxKeys; }

int attention_bwd_mma_pipe(const AttnBwdArgs& a, cudaStream_t stream) {
  const int smem_a = static_cast<int>(sizeof(TilesA)), smem_b = static_cast<int>(sizeof(TilesB));
  TAN_CUDA(cudaFuncSetAttribute(attn_bwd_dq_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a));
  TAN_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
  attn_bwd_dq_pipe_kernel<<<dim3((a.Lq + kBlk - 1) / kBlk, a.H, a.B), 128, smem_a, stream>>>(a);
  TAN_CUDA(cudaGetLastError());
  attn_bwd_dkv_pipe_kernel<<<dim3((a.Lk + kBlk - 1) / kBlk, a.H, a.B), 128, smem_b, stream>>>(a);
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}

}  // namespace tanb
