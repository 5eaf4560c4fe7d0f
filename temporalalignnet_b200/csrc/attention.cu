// tan_attention_bf16: multi-head softmax attention core on tcgen05 tensor cores, head_dim 64,
// arbitrary key-padding mask, Lq != Lk allowed (cross-attention).
//
// One CTA per (128-query tile, head, clip), 192 threads, two CTAs per SM:
//   warp 0      TMA producer (Q tile once, K/V blocks of 64 keys through a 3-stage ring) + TMEM allocator
//   warp 1      MMA issuer:  S  = Q K_j^T     (M=128, N=64, K=64; both operands K-major, 128B swizzle)
//                            O_j = P_j V_j     (M=128, N=64, K=64; V is consumed as it lies in HBM,
//                                               [keys, 64] = MN-major B operand, no transpose anywhere)
//   warps 2-5   softmax: thread = one query row (TMEM lane), so row max / row sum are thread-local
//               (no shuffles); S comes from TMEM with tcgen05.ld, P_j = exp2(S - m) goes back to shared
//               memory as the bf16 A operand of the second MMA (128B-swizzled K-major, the layout TMA
//               would have produced), O accumulates in registers: O = O * alpha_j + O_j (online softmax),
//               so nothing in TMEM ever needs rescaling and the only hand-offs are three mbarriers per block.
//               S, P and O_j are double buffered and the O_j accumulation is deferred by one block, so QK_{j+1}
//               and PV_j execute under the softmax of the neighbouring blocks.
// The L x L score matrix never exists outside TMEM.  Per 64-key block a CTA reads 64 KB out of TMEM
// (S and O_j, ~64 B/clk/SM) and issues 8192 exp2 (16/clk/SM): both ~1k cycles against 256 cycles of MMA, so
// the kernel is bound by TMEM-read / MUFU throughput, not by the tensor pipe; two resident CTAs overlap one
// CTA's softmax with the other's loads and MMAs.
#include "common.cuh"

namespace tanb {

constexpr int kAttBQ = 128;
constexpr int kAttBK = 64;
constexpr int kAttThreads = 192;
constexpr int kAttQBytes = kAttBQ * 128;      // 16 KB
constexpr int kAttKVBytes = kAttBK * 128;     // 8 KB
constexpr int kAttStages = 3;
constexpr int kAttSmem = kAttQBytes /*Q*/ + kAttStages * 2 * kAttKVBytes /*K,V ring*/ + 2 * kAttQBytes /*P x 2*/ +
                         1024 /*bars*/ + 1024 /*alignment slack*/;

__device__ __forceinline__ uint32_t att_swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__global__ void __launch_bounds__(kAttThreads, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const uint8_t* __restrict__ kpm, bf16* __restrict__ out,
                 int64_t ldo, int Lq, int Lk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAttQBytes;                 // [3][8 KB]
  uint8_t* sV = sK + kAttStages * kAttKVBytes;   // [3][8 KB]
  uint8_t* sP = sV + kAttStages * kAttKVBytes;   // [2][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kAttQBytes);
  uint64_t* q_full = bars;            // [1]
  uint64_t* kv_full = bars + 1;       // [3]
  uint64_t* kv_empty = bars + 4;      // [3]
  uint64_t* s_full = bars + 7;        // [2]
  uint64_t* p_ready = bars + 9;       // [2] count 4 (one arrive per softmax warp)
  uint64_t* o_full = bars + 11;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = qt * kAttBQ;
  const int nb = (Lk + kAttBK - 1) / kAttBK;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_init(q_full, 1);
      for (int i = 0; i < kAttStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&o_full[i], 1); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);        // S[2]: columns [0,64), [64,128); O_j[2]: [128,192), [192,256)
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kAttQBytes);
      tma_load_2d(sQ, &tmQ, q_full, h * 64, b * Lq + q0);
      for (int j = 0; j < nb; ++j) {
        const int st = j % kAttStages;
        mbar_wait(&kv_empty[st], ((j / kAttStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * kAttKVBytes);
        tma_load_2d(sK + st * kAttKVBytes, &tmK, &kv_full[st], h * 64, b * Lk + j * kAttBK);
        tma_load_2d(sV + st * kAttKVBytes, &tmV, &kv_full[st], h * 64, b * Lk + j * kAttBK);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc_qk = umma_idesc_bf16(kAttBQ, kAttBK);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(kAttBQ, 64) | (1u << 16);     // B (= V) is MN-major
    // S and O_j are double buffered: QK_{j+2} is issued as soon as S_j has been consumed, PV_j as soon as
    // P_j is staged, so the softmax warps (the bottleneck) always find their next S tile ready.
    auto issue_qk = [&](int j) {
      const int st = j % kAttStages;
      mbar_wait(&kv_full[st], (j / kAttStages) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ));
        const uint64_t dk = umma_desc_k_sw128(smem_u32(sK + st * kAttKVBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + (j & 1) * 64, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
        tc_commit(&s_full[j & 1]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    if (nb > 1) issue_qk(1);
    for (int j = 0; j < nb; ++j) {
      mbar_wait(&p_ready[j & 1], (j >> 1) & 1);  // P_j is in shared memory, S_j and O_{j-2} have been consumed
      tc_fence_after();
      if (lane == 0) {
        const int st = j % kAttStages;
        const uint64_t dp = umma_desc_k_sw128(smem_u32(sP + (j & 1) * kAttQBytes));
        const uint64_t dv = umma_desc_k_sw128(smem_u32(sV + st * kAttKVBytes));
        // 16 keys per MMA: +32 B along P's rows (K-major), +16 rows x 128 B = 2048 B in V (MN-major)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + 128 + (j & 1) * 64, dp + 2 * k, dv + 128 * k, idesc_pv, k != 0);
        tc_commit(&o_full[j & 1]);
        tc_commit(&kv_empty[st]);                // K_j (read by the earlier QK MMAs) and V_j are free
      }
      __syncwarp();
      if (j + 2 < nb) issue_qk(j + 2);
    }
  } else {
    // ===== softmax / output (warps 2..5) =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;               // query row inside the tile = TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sl2 = 0.125f * 1.4426950408889634f;    // 1/sqrt(64) folded with log2(e)
    const uint8_t* mb = kpm != nullptr ? kpm + static_cast<int64_t>(b) * Lk : nullptr;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

    // o <- o * alpha_j + O_j, deferred by one block so that PV_j runs under the softmax of block j+1
    auto accumulate = [&](int j, float alpha) {
      mbar_wait(&o_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t a0[32], a1[32];
      const uint32_t t_o = t_lane + 128 + (j & 1) * 64;
      tmem_ld_32x32(t_o, a0);
      tmem_ld_32x32(t_o + 32, a1);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        o[c] = fmaf(o[c], alpha, __uint_as_float(a0[c]));
        o[32 + c] = fmaf(o[32 + c], alpha, __uint_as_float(a1[c]));
      }
      tc_fence_before();                         // O_j reads retired before PV_{j+2} (ordered by p_ready)
    };

    for (int j = 0; j < nb; ++j) {
      // key mask of this block as two warp-uniform words (bit set = ignore key)
      const int k0 = j * kAttBK + lane, k1 = k0 + 32;
      const bool ig0 = k0 >= Lk || (mb != nullptr && mb[k0] != 0);
      const bool ig1 = k1 >= Lk || (mb != nullptr && mb[k1] != 0);
      const uint32_t w0 = __ballot_sync(0xffffffffu, ig0), w1 = __ballot_sync(0xffffffffu, ig1);

      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      const uint32_t t_s = t_lane + (j & 1) * 64;
      tmem_ld_32x32(t_s, r0);
      tmem_ld_32x32(t_s + 32, r1);
      tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float a = ((w0 >> c) & 1u) ? -INFINITY : __uint_as_float(r0[c]) * sl2;
        const float bq = ((w1 >> c) & 1u) ? -INFINITY : __uint_as_float(r1[c]) * sl2;
        r0[c] = __float_as_uint(a);
        r1[c] = __float_as_uint(bq);
        mx = fmaxf(mx, fmaxf(a, bq));
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = (m_new == -INFINITY) ? 1.f : fast_exp2(m_run - m_new);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      m_run = m_new;
      float ls = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float p0 = fast_exp2(__uint_as_float(r0[2 * c]) - m_use);
        const float p1 = fast_exp2(__uint_as_float(r0[2 * c + 1]) - m_use);
        const float p2 = fast_exp2(__uint_as_float(r1[2 * c]) - m_use);
        const float p3 = fast_exp2(__uint_as_float(r1[2 * c + 1]) - m_use);
        ls += (p0 + p1) + (p2 + p3);
        pk[c] = pack_bf16x2(p0, p1);            // keys 2c, 2c+1
        pk[16 + c] = pack_bf16x2(p2, p3);       // keys 32+2c, 32+2c+1
      }
      l_run = l_run * alpha + ls;
      // P row: 64 keys = 128 bytes = 8 chunks of 8 keys (sP[j&1] is free: o_full_{j-2} was awaited last iteration)
      uint8_t* sPj = sP + (j & 1) * kAttQBytes;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        *reinterpret_cast<uint4*>(sPj + att_swz(row, ch)) =
            make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
      fence_proxy_async_smem();                  // P visible to the tensor core (async proxy)
      tc_fence_before();                         // S reads retired before QK_{j+2} overwrites S[j&1]
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[j & 1]);

      if (j >= 1) accumulate(j - 1, alpha_prev);
      alpha_prev = alpha;
    }
    accumulate(nb - 1, alpha_prev);

    // finalize: normalise, stage this warp's 32 rows in its slice of sP (free: the last PV MMA has
    // completed), then write whole 128-byte rows: lane = (row % 4, 16-byte chunk)
    const float inv = 1.f / l_run;               // l == 0 (all keys masked) -> NaN, as torch
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      uint4 u;
      u.x = pack_bf16x2(o[8 * ch] * inv, o[8 * ch + 1] * inv);
      u.y = pack_bf16x2(o[8 * ch + 2] * inv, o[8 * ch + 3] * inv);
      u.z = pack_bf16x2(o[8 * ch + 4] * inv, o[8 * ch + 5] * inv);
      u.w = pack_bf16x2(o[8 * ch + 6] * inv, o[8 * ch + 7] * inv);
      *reinterpret_cast<uint4*>(sP + att_swz(row, ch)) = u;
    }
    __syncwarp();
    const int rr = lane >> 3, cc = lane & 7;
    bf16* ob = out + (static_cast<int64_t>(b) * Lq + q0 + quarter * 32) * ldo + h * 64 + cc * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rl = 4 * i + rr;
      if (q0 + quarter * 32 + rl < Lq)
        *reinterpret_cast<uint4*>(ob + static_cast<int64_t>(rl) * ldo) =
            *reinterpret_cast<const uint4*>(sP + att_swz(quarter * 32 + rl, cc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
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
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8 || ldq < H * 64 || ldk < H * 64 || ldv < H * 64 || ldo < H * 64)
    return set_error(TAN_ERR_SHAPE, "tan_attention_bf16: row pitches must cover H*64 columns and be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(out)) & 15)
    return set_error(TAN_ERR_SHAPE, "tan_attention_bf16: q/k/v/out must be 16-byte aligned");
  if (static_cast<int64_t>(B) * Lq > 0x7fffffffll || static_cast<int64_t>(B) * Lk > 0x7fffffffll)
    return set_error(TAN_ERR_SHAPE, "tan_attention_bf16: B*L exceeds the TMA coordinate range");
  CUtensorMap tmQ, tmK, tmV;
  TAN_CHECK(make_tmap_2d(&tmQ, q, 2, static_cast<uint64_t>(B) * Lq, static_cast<uint64_t>(H) * 64, ldq, kAttBQ));
  TAN_CHECK(make_tmap_2d(&tmK, k, 2, static_cast<uint64_t>(B) * Lk, static_cast<uint64_t>(H) * 64, ldk, kAttBK));
  TAN_CHECK(make_tmap_2d(&tmV, v, 2, static_cast<uint64_t>(B) * Lk, static_cast<uint64_t>(H) * 64, ldv, kAttBK));
  static bool attr_set = false;
  if (!attr_set) {
    TAN_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
    attr_set = true;
  }
  dim3 grid((Lq + kAttBQ - 1) / kAttBQ, H, B);
  return launch_pdl(attention_kernel, grid, dim3(kAttThreads), kAttSmem, static_cast<cudaStream_t>(stream), 1, tmQ,
                    tmK, tmV, key_padding_mask, static_cast<bf16*>(out), ldo, Lq, Lk);
}
