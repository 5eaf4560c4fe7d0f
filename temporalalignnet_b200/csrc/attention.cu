// tan_attention_bf16: multi-head softmax attention core, head_dim 64, arbitrary key-padding mask.
//
// v1 data path: one CTA per (64-query tile, head, clip); K/V tiles of 64 keys stream through a
// double-buffered cp.async ring in XOR-swizzled shared memory; S = Q K^T and O += P V run on the
// warp-level tensor-core path (mma.sync m16n8k16, bf16 in / fp32 accumulate) with an online
// (running max / running sum) softmax in the log2 domain, so the L x L score matrix never exists
// in HBM.  The attention core is 4 L^2 d of the layer's 24 L d^2 + 4 L^2 d flops (7.7 % at L=256,
// d=512); the projections around it run on tcgen05 (gemm_linear.cu).
#include "common.cuh"

namespace tanb {

constexpr int kAttBQ = 64;
constexpr int kAttBK = 64;
constexpr int kAttHD = 64;
constexpr int kAttThreads = 128;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t n = valid ? 16u : 0u;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// tile [64 rows][64 bf16] = 8 chunks of 16 B per row; chunk index XOR (row & 7) kills ldmatrix conflicts
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ void load_tile(uint8_t* sdst, const bf16* gbase, int64_t ld, int row0, int nrows_total) {
  // 64 rows x 8 chunks = 512 chunks, 128 threads -> 4 each
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = threadIdx.x + i * kAttThreads;
    const int row = idx >> 3, chunk = idx & 7;
    const int grow = row0 + row;
    const bool ok = grow < nrows_total;
    const bf16* src = gbase + static_cast<int64_t>(ok ? grow : 0) * ld + chunk * 8;
    cp_async16(sdst + swz(row, chunk), src, ok);
  }
}

__global__ void __launch_bounds__(kAttThreads)
attention_kernel(const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k, int64_t ldk,
                 const bf16* __restrict__ v, int64_t ldv, const uint8_t* __restrict__ kpm, bf16* __restrict__ out,
                 int64_t ldo, int Lq, int Lk) {
  __shared__ __align__(128) uint8_t sQ[kAttBQ * 128];
  __shared__ __align__(128) uint8_t sK[2][kAttBK * 128];
  __shared__ __align__(128) uint8_t sV[2][kAttBK * 128];
  __shared__ uint8_t sM[2][kAttBK];

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* qb = q + (static_cast<int64_t>(b) * Lq) * ldq + h * kAttHD;
  const bf16* kb = k + (static_cast<int64_t>(b) * Lk) * ldk + h * kAttHD;
  const bf16* vb = v + (static_cast<int64_t>(b) * Lk) * ldv + h * kAttHD;
  const uint8_t* mb = kpm ? kpm + static_cast<int64_t>(b) * Lk : nullptr;
  const int q0 = qt * kAttBQ;
  const int nk = (Lk + kAttBK - 1) / kAttBK;

  auto load_mask = [&](int buf, int key0) {
    if (threadIdx.x < kAttBK) {
      const int key = key0 + threadIdx.x;
      sM[buf][threadIdx.x] = (key < Lk) ? (mb ? mb[key] : uint8_t(0)) : uint8_t(1);
    }
  };

  load_tile(sQ, qb, ldq, q0, Lq);
  load_tile(sK[0], kb, ldk, 0, Lk);
  load_tile(sV[0], vb, ldv, 0, Lk);
  load_mask(0, 0);
  cp_async_commit();

  // softmax scale folded with log2(e): head_dim 64 -> 1/8
  const float sl2 = 0.125f * 1.4426950408889634f;
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  uint32_t qf[4][4];
  bool q_loaded = false;

  for (int j = 0; j < nk; ++j) {
    const int buf = j & 1;
    if (j + 1 < nk) {
      load_tile(sK[buf ^ 1], kb, ldk, (j + 1) * kAttBK, Lk);
      load_tile(sV[buf ^ 1], vb, ldv, (j + 1) * kAttBK, Lk);
      load_mask(buf ^ 1, (j + 1) * kAttBK);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (!q_loaded) {
      // A fragments of this warp's 16 query rows, 4 k-steps of 16
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int row = warp * 16 + (lane & 15);
        const int chunk = kk * 2 + (lane >> 4);
        ldsm_x4(smem_u32(sQ + swz(row, chunk)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
      }
      q_loaded = true;
    }

    // S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {   // pairs of 8-key n-tiles
        const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int chunk = kk * 2 + ((lane >> 3) & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(sK[buf] + swz(row, chunk)), b0, b1, b2, b3);
        mma_bf16_16816(s[2 * np], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b0, b1);
        mma_bf16_16816(s[2 * np + 1], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b2, b3);
      }
    }

    // mask + online softmax (rows lane/4 and lane/4 + 8)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + (lane & 3) * 2;
      const bool k0 = sM[buf][c] != 0, k1 = sM[buf][c + 1] != 0;
      s[nt][0] = k0 ? -INFINITY : s[nt][0] * sl2;
      s[nt][1] = k1 ? -INFINITY : s[nt][1] * sl2;
      s[nt][2] = k0 ? -INFINITY : s[nt][2] * sl2;
      s[nt][3] = k1 ? -INFINITY : s[nt][3] * sl2;
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float scale[2], m_use[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      scale[r] = (m_new == -INFINITY) ? 1.f : fast_exp2(m_run[r] - m_new);
      m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
      m_run[r] = m_new;
    }
    float ls[2] = {0.f, 0.f};
    uint32_t pf[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = fast_exp2(s[nt][0] - m_use[0]);
      const float p1 = fast_exp2(s[nt][1] - m_use[0]);
      const float p2 = fast_exp2(s[nt][2] - m_use[1]);
      const float p3 = fast_exp2(s[nt][3] - m_use[1]);
      ls[0] += p0 + p1;
      ls[1] += p2 + p3;
      pf[nt][0] = pack_bf16x2(p0, p1);
      pf[nt][1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * scale[r] + ls[r];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      o[nt][0] *= scale[0]; o[nt][1] *= scale[0];
      o[nt][2] *= scale[1]; o[nt][3] *= scale[1];
    }

    // O += P V   (keys are the k dimension: 4 k-steps of 16 keys; 8 n-tiles of 8 head dims)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        const int row = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int chunk = np * 2 + (lane >> 4);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(smem_u32(sV[buf] + swz(row, chunk)), b0, b1, b2, b3);
        mma_bf16_16816(o[2 * np], pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1], b0, b1);
        mma_bf16_16816(o[2 * np + 1], pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1], b2, b3);
      }
    }
    __syncthreads();   // everyone done with buf before the next iteration's prefetch overwrites it
  }

  // finalize: full row sums across the quad, normalise, store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];   // l == 0 (all keys masked) -> NaN, as torch
  const int r0 = q0 + warp * 16 + (lane >> 2);
  const int r1 = r0 + 8;
  bf16* ob = out + (static_cast<int64_t>(b) * Lq) * ldo + h * kAttHD;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = nt * 8 + (lane & 3) * 2;
    if (r0 < Lq) *reinterpret_cast<uint32_t*>(ob + static_cast<int64_t>(r0) * ldo + c) = pack_bf16x2(o[nt][0] * inv0, o[nt][1] * inv0);
    if (r1 < Lq) *reinterpret_cast<uint32_t*>(ob + static_cast<int64_t>(r1) * ldo + c) = pack_bf16x2(o[nt][2] * inv1, o[nt][3] * inv1);
  }
}

}  // namespace tanb

using namespace tanb;

extern "C" int tan_attention_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                  int64_t ldv, const uint8_t* key_padding_mask, void* out, int64_t ldo, int B, int H,
                                  int Lq, int Lk, void* stream) {
  TAN_CHECK(tan_device_check());
  if (q == nullptr || k == nullptr || v == nullptr || out == nullptr)
    return set_error(TAN_ERR_ARG, "tan_attention_bf16: null pointer");
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0 || B > 65535 || H > 65535)
    return set_error(TAN_ERR_SHAPE, "tan_attention_bf16: bad dims B=%d H=%d Lq=%d Lk=%d", B, H, Lq, Lk);
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 2 || ldq < H * 64 || ldk < H * 64 || ldv < H * 64 || ldo < H * 64)
    return set_error(TAN_ERR_SHAPE, "tan_attention_bf16: row pitches must cover H*64 columns and be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15)
    return set_error(TAN_ERR_SHAPE, "tan_attention_bf16: q/k/v must be 16-byte aligned");
  dim3 grid((Lq + kAttBQ - 1) / kAttBQ, H, B);
  attention_kernel<<<grid, kAttThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), ldk, static_cast<const bf16*>(v), ldv,
      key_padding_mask, static_cast<bf16*>(out), ldo, Lq, Lk);
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}
