// tan_attention_bf16: multi-head softmax attention core on tcgen05 tensor cores, head_dim 64,
// arbitrary key-padding mask, Lq != Lk allowed (cross-attention).
//
// One CTA per (clip, head), 192 threads, two CTAs per SM.  The CTA walks ALL 128-query tiles of its (clip, head)
// and, inside each, the 64-key blocks; every pipeline (Q tiles, K ring, V ring, S / P / O buffers) keeps running
// across query-tile boundaries, so the launch / TMEM-allocation / first-load latency is paid once per (clip,
// head) instead of once per 4-5 key blocks (the per-tile CTAs of the first versions spent about half of their
// ~8 us lifetime in that prologue: profiles/r01b_prof_attn, 15 % tensor-pipe activity).
//   warp 0      TMA producer: Q tile (double buffered), K blocks (2-stage ring, freed by QK), V blocks (3-stage
//               ring, freed by PV) + TMEM allocator
//   warp 1      MMA issuer:  S_g = Q K_g^T    (M=128, N=64, K=64; both operands K-major, 128B swizzle)
//                            O  += P_g V_g     (M=128, N=64, K=64; V is consumed as it lies in HBM,
//                                               [keys, 64] = MN-major B operand, no transpose anywhere)
//   warps 2-5   softmax: thread = one query row (TMEM lane), so row max / row sum are thread-local
//               (no shuffles); S_g comes from TMEM with tcgen05.ld, P_g = exp2(S_g - m) goes back to shared
//               memory as the bf16 A operand of the second MMA (128B-swizzled K-major, the layout TMA
//               would have produced).
// O ACCUMULATES IN TMEM across the key blocks (double buffered across query tiles) and is read once per tile.
// The reference max m of a row is only raised when a block's maximum exceeds it by more than 8 (log2 domain,
// i.e. P <= 256: harmless for bf16 P and fp32 sums); only then the warp rescales its 32 rows of O in TMEM
// (tcgen05.ld / st) -- after the first block this almost never happens.  A TMEM read moves 64 B/clk/SM, so
// reading S_g (32 KB per block) costs as much as the block's 8192 exp2 on the MUFU pipe (512 clk each), both 2x
// the two MMAs: that, not the tensor pipe, is the floor of head_dim-64 attention.  S and P are double buffered
// so QK_{g+1} and PV_g execute under the softmax of the neighbouring blocks.  The L x L score matrix never
// exists outside TMEM.
#include "common.cuh"

namespace tanb {

constexpr int kAttBQ = 128;
constexpr int kAttBK = 64;
constexpr int kAttThreads = 192;
constexpr int kAttQBytes = kAttBQ * 128;      // 16 KB
constexpr int kAttKVBytes = kAttBK * 128;     // 8 KB
// Round-2 per-CTA timelines (scripts/attn_trace.py; profiles/r02i_attention_timeline_before.txt): at L = 256 a CTA lives ~26 k
// clk = 5.4 k until its first QK (prologue + ~3.7 k clk for the first Q / K boxes), 8 blocks at a ~2.0 k clk cadence
// and two tile epilogues of ~3.3 k clk (half of it waiting for the tile's last PV).  The softmax of a block takes
// ~1.35 k clk; the rest of the cadence is the MMA warp's serial path between p_ready(g) and a usable S_{g+2}: ~0.2 k
// to see the barrier, 4 tcgen05.mma ~0.23 k (~55 clk per issue), two tcgen05.commit + the next waits ~0.65 k, 4 more
// MMAs.  Three variants were measured against that picture and REJECTED (same box, kept under profiles/):
//   * L2 prefetch of every K / V / Q box at CTA start (r02j): the prefetches queue ahead of the real loads (first QK
//     at 9.6 k instead of 5.1 k clk), the cadence does not change -- the loads are not what paces the blocks;
//   * 4-deep K and V rings with a single P buffer (r02k): 178.9 vs 182.7 us per launch, within noise;
//   * releasing S_g as soon as it is in registers and issuing QK_{g+2} during the softmax of block g (r02l): PV_g
//     then waits behind issue_qk's K wait, softmax blocks lengthen to ~1.5 k clk, 188 vs 169 us per launch.
// Ring depths (a 3-deep K ring was measured slower: 1.96 vs 1.62 ms per step at the bench shape).  The dynamic
// segment is declared 1024-byte aligned instead of carrying an alignment slack; two CTAs per SM.
#ifndef TAN_ATT_KSTAGES
#define TAN_ATT_KSTAGES 2
#endif
constexpr int kAttKStages = TAN_ATT_KSTAGES;
constexpr int kAttVStages = 3;
constexpr int kAttMaskWords = 128;            // mask bits for Lk <= 4096 (longer sequences: per-block loads)
constexpr int kAttSmem = 2 * kAttQBytes /*Q x 2*/ + (kAttKStages + kAttVStages) * kAttKVBytes + 2 * kAttQBytes /*P x 2*/ +
                         256 /*bars*/ + kAttMaskWords * 4;
static_assert(2 * (kAttSmem + 1024) <= 228 * 1024, "two CTAs per SM");

__device__ __forceinline__ uint32_t att_swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// Development aid (tan_debug_set_trace): per-CTA clock64 stamps, 128 slots per CTA for the first 512 CTAs.
// 0 globaltimer, 1 start, 2 prologue done, 3 exit; block g < 10 at 8 + 10 g: +0 K issued, +1 V issued (producer),
// +2 QK issued, +3 p_ready seen, +4 PV issued (MMA warp), +5 s_full seen, +6 S in registers, +7 P staged (softmax warp 2)
extern __device__ long long* g_gemm_trace;
// (compiled in with -DTAN_ATT_TRACE only -- `make variant SRC=attention NAME=trace FLAGS="-DTAN_WAIT_HINT_ALL=0
// -DTAN_ATT_TRACE"`, then TAN_LIB_PATH=.../libtan_b200_trace.so python scripts/attn_trace.py B H L: eight predicated
// stamp sequences per key block are instructions the softmax loop does not need)
__device__ __forceinline__ void att_trace(long long* tr, int g, int k) {
#ifdef TAN_ATT_TRACE
  if (tr != nullptr && g < 10) tr[8 + 10 * g + k] = clock64();
#endif
}

__global__ void __launch_bounds__(kAttThreads, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const uint8_t* __restrict__ kpm, bf16* __restrict__ out,
                 int64_t ldo, int Lq, int Lk, int tiles_per_cta, float* __restrict__ lse, int lse_ld) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();     // the 128-byte swizzle atoms need 1024-byte aligned tiles
  uint8_t* sQ = smem;                               // [2][16 KB]
  uint8_t* sK = sQ + 2 * kAttQBytes;                // [2][8 KB]
  uint8_t* sV = sK + kAttKStages * kAttKVBytes;     // [3][8 KB]
  uint8_t* sP = sV + kAttVStages * kAttKVBytes;     // [2][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kAttQBytes);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2] the tile's last QK has completed
  uint64_t* k_full = bars + 4;        // [3]
  uint64_t* k_empty = bars + 7;       // [3]
  uint64_t* v_full = bars + 10;       // [3]
  uint64_t* v_empty = bars + 13;      // [3]
  uint64_t* s_full = bars + 16;       // [2]
  uint64_t* p_ready = bars + 18;      // [2] count 4 (one arrive per softmax warp)
  uint64_t* pv_done = bars + 20;      // [2] PV_g has completed (g & 1): P buffer free, O stable
  uint64_t* o_free = bars + 22;       // [2] count 4: the tile's O has been read out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  // key mask of the clip as bits (bit set = ignore key).  Inside the dynamic segment on purpose: a static
  // __shared__ array would shift the 1024-aligned dynamic base without growing the allocation.
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(bars + 32);

  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* tr = g_gemm_trace;
  {
    const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    tr = (tr != nullptr && cta < 512) ? tr + cta * 128 : nullptr;
    if (tr != nullptr && threadIdx.x == 0) {
      long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      tr[0] = gt;
      tr[1] = clock64();
    }
  }
  // this CTA's query tiles: [qt0, qt0 + nq) (all of them unless the launch splits long sequences over blockIdx.z)
  const int qt0 = blockIdx.z * tiles_per_cta;
  const int nq = min(tiles_per_cta, (Lq + kAttBQ - 1) / kAttBQ - qt0);
  const int nb = (Lk + kAttBK - 1) / kAttBK;
  const int G = nq * nb;                            // (query tile, key block) pairs, g = qt * nb + j

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
        mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&pv_done[i], 1); mbar_init(&o_free[i], 4);
      }
      for (int i = 0; i < kAttKStages; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
      for (int i = 0; i < kAttVStages; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);        // S[2]: columns [0,64), [64,128); O[2]: [128,192), [192,256)
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  if (tr != nullptr && threadIdx.x == 0) tr[2] = clock64();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int qt = 0; qt < nq; ++qt) {
        const int qb = qt & 1;
        if (qt >= 2) mbar_wait_relaxed(&q_empty[qb], ((qt >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&q_full[qb], kAttQBytes);
        tma_load_2d(sQ + qb * kAttQBytes, &tmQ, &q_full[qb], h * 64, b * Lq + (qt0 + qt) * kAttBQ);
        for (int j = 0; j < nb; ++j) {
          const int g = qt * nb + j;
          const int ks = g % kAttKStages, vs = g % kAttVStages;
          mbar_wait_relaxed(&k_empty[ks], ((g / kAttKStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[ks], kAttKVBytes);
          tma_load_2d(sK + ks * kAttKVBytes, &tmK, &k_full[ks], h * 64, b * Lk + j * kAttBK);
          att_trace(tr, g, 0);
          mbar_wait_relaxed(&v_empty[vs], ((g / kAttVStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[vs], kAttKVBytes);
          tma_load_2d(sV + vs * kAttKVBytes, &tmV, &v_full[vs], h * 64, b * Lk + j * kAttBK);
          att_trace(tr, g, 1);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc_qk = umma_idesc_bf16(kAttBQ, kAttBK);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(kAttBQ, 64) | (1u << 16);     // B (= V) is MN-major
    // S is double buffered: QK_{g+2} is issued as soon as S_g has been consumed, PV_g as soon as P_g is staged,
    // so the softmax warps (the bottleneck) always find their next S tile ready -- also across query tiles.
    auto issue_qk = [&](int g) {
      const int qt = g / nb, j = g - qt * nb;
      const int qb = qt & 1, ks = g % kAttKStages;
      if (j == 0) mbar_wait(&q_full[qb], (qt >> 1) & 1);
      mbar_wait(&k_full[ks], (g / kAttKStages) & 1);
      tc_fence_after();
      if (TAN_MMA_LEADER()) {
        const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ + qb * kAttQBytes));
        const uint64_t dk = umma_desc_k_sw128(smem_u32(sK + ks * kAttKVBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + (g & 1) * 64, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
        att_trace(tr, g, 2);
        tc_commit(&s_full[g & 1]);
        tc_commit(&k_empty[ks]);
        if (j == nb - 1) tc_commit(&q_empty[qb]);
      }
      __syncwarp();
    };
    issue_qk(0);
    if (G > 1) issue_qk(1);
    for (int g = 0; g < G; ++g) {
      const int qt = g / nb, j = g - qt * nb;
      const int ob = qt & 1, vs = g % kAttVStages;
      mbar_wait(&p_ready[g & 1], (g >> 1) & 1);  // P_g is in shared memory, S_g consumed, O rescaled if needed
      if (lane == 0) att_trace(tr, g, 3);
      if (j == 0 && qt >= 2) mbar_wait(&o_free[ob], ((qt >> 1) - 1) & 1);   // tile qt-2 has left this O buffer
      mbar_wait(&v_full[vs], (g / kAttVStages) & 1);
      tc_fence_after();
      if (TAN_MMA_LEADER()) {
        const uint64_t dp = umma_desc_k_sw128(smem_u32(sP + (g & 1) * kAttQBytes));
        const uint64_t dv = umma_desc_k_sw128(smem_u32(sV + vs * kAttKVBytes));
        // 16 keys per MMA: +32 B along P's rows (K-major), +16 rows x 128 B = 2048 B in V (MN-major)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + 128 + ob * 64, dp + 2 * k, dv + 128 * k, idesc_pv, (j | k) != 0);
        att_trace(tr, g, 4);
        tc_commit(&pv_done[g & 1]);
        tc_commit(&v_empty[vs]);
      }
      __syncwarp();
      if (g + 2 < G) issue_qk(g + 2);
    }
  } else {
    // ===== softmax / output (warps 2..5) =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;               // query row inside the tile = TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sl2 = 0.125f * 1.4426950408889634f;    // 1/sqrt(64) folded with log2(e)
    const uint8_t* mb = kpm != nullptr ? kpm + static_cast<int64_t>(b) * Lk : nullptr;
    // the clip's key mask as bits, once (a per-block byte load would put a global-memory round trip on the
    // critical path of every block: 5 % of the samples in ncu r01e)
    const bool mask_in_smem = nb * 2 <= kAttMaskWords;
    if (mask_in_smem) {
      for (int wd = quarter; wd < nb * 2; wd += 4) {
        const int key = wd * 32 + lane;
        const bool ig = key >= Lk || (mb != nullptr && mb[key] != 0);
        const uint32_t bits = __ballot_sync(0xffffffffu, ig);
        if (lane == 0) s_mask[wd] = bits;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }

    for (int qt = 0; qt < nq; ++qt) {
      const int q0 = (qt0 + qt) * kAttBQ;
      const uint32_t t_o = t_lane + 128 + (qt & 1) * 64;
      const bool live = q0 + quarter * 32 < Lq;        // warp-uniform: this warp owns at least one real query row
      float m_ref = -INFINITY, l_run = 0.f;

      for (int j = 0; j < nb; ++j) {
        const int g = qt * nb + j;
        mbar_wait(&s_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        long long* trw = (warp == 2 && lane == 0) ? tr : nullptr;
        att_trace(trw, g, 5);
        if (live) {
          // key mask of this block as two warp-uniform words (bit set = ignore key)
          uint32_t w0, w1;
          if (mask_in_smem) {
            w0 = s_mask[2 * j];
            w1 = s_mask[2 * j + 1];
          } else {
            const int k0 = j * kAttBK + lane, k1 = k0 + 32;
            const bool ig0 = k0 >= Lk || (mb != nullptr && mb[k0] != 0);
            const bool ig1 = k1 >= Lk || (mb != nullptr && mb[k1] != 0);
            w0 = __ballot_sync(0xffffffffu, ig0);
            w1 = __ballot_sync(0xffffffffu, ig1);
          }
          uint32_t r0[32], r1[32];
          const uint32_t t_s = t_lane + (g & 1) * 64;
          tmem_ld_32x32(t_s, r0);
          tmem_ld_32x32(t_s + 32, r1);
          tmem_ld_wait();
          att_trace(trw, g, 6);
          float mx = -INFINITY;
          if ((w0 | w1) != 0u) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              if ((w0 >> c) & 1u) r0[c] = 0xff800000u;  // -inf
              if ((w1 >> c) & 1u) r1[c] = 0xff800000u;
            }
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[c]), __uint_as_float(r1[c])));
          mx *= sl2;                                  // sl2 > 0: -inf stays -inf
          // lazy reference max: raise it only when this block exceeds it by more than 2^8
          const bool need = (j == 0) ? (mx > m_ref) : (mx > m_ref + 8.0f);
          if (j > 0 && __any_sync(0xffffffffu, need)) {
            // rescale this warp's rows of O (and l) to the new reference; PV_{g-1} must have completed
            mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
            tc_fence_after();
            const float m_new = need ? mx : m_ref;
            const float alpha = need ? fast_exp2(m_ref - m_new) : 1.f;    // m_ref = -inf -> 0 (O and l are 0 then)
            uint32_t a0[32];
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              tmem_ld_32x32(t_o + hlf * 32, a0);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) a0[c] = __float_as_uint(__uint_as_float(a0[c]) * alpha);
              tmem_st_32x32(t_o + hlf * 32, a0);
            }
            tmem_st_wait();
            l_run *= alpha;
            m_ref = m_new;
          } else if (need) {
            m_ref = mx;                               // first block of the tile
          }
          const float nm = (m_ref == -INFINITY) ? 0.f : -m_ref;
          float ls = 0.f;
          uint32_t pk[32];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(r0[2 * c]), sl2, nm));
            const float p1 = fast_exp2(fmaf(__uint_as_float(r0[2 * c + 1]), sl2, nm));
            const float p2 = fast_exp2(fmaf(__uint_as_float(r1[2 * c]), sl2, nm));
            const float p3 = fast_exp2(fmaf(__uint_as_float(r1[2 * c + 1]), sl2, nm));
            ls += (p0 + p1) + (p2 + p3);
            pk[c] = pack_bf16x2(p0, p1);            // keys 2c, 2c+1
            pk[16 + c] = pack_bf16x2(p2, p3);       // keys 32+2c, 32+2c+1
          }
          l_run += ls;
          // sP[g&1] was the A operand of PV_{g-2}
          if (g >= 2) mbar_wait(&pv_done[g & 1], ((g >> 1) - 1) & 1);
          // P row: 64 keys = 128 bytes = 8 chunks of 8 keys
          uint8_t* sPg = sP + (g & 1) * kAttQBytes;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<uint4*>(sPg + att_swz(row, ch)) =
                make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
          fence_proxy_async_smem();                  // P visible to the tensor core (async proxy)
          att_trace(trw, g, 7);
        }
        tc_fence_before();                           // S reads / O writes retired before the MMAs that follow p_ready
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[g & 1]);
      }
      // all MMAs of the tile have completed when its last PV has
      const int gl = qt * nb + nb - 1;
      mbar_wait(&pv_done[gl & 1], (gl >> 1) & 1);
      tc_fence_after();
      // log2-domain log-sum-exp of the row's scaled scores for the backward pass (rows Lq .. lse_ld get +inf, so
      // that exp2(s - lse) of a padding row is 0 there without a predicate)
      if (lse != nullptr && q0 + row < lse_ld)
        lse[(static_cast<int64_t>(b) * gridDim.x + h) * lse_ld + q0 + row] =
            (live && q0 + row < Lq) ? m_ref + __log2f(l_run) : INFINITY;
      if (live) {
        float o[64];
        {
          uint32_t a0[32], a1[32];
          tmem_ld_32x32(t_o, a0);
          tmem_ld_32x32(t_o + 32, a1);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) { o[c] = __uint_as_float(a0[c]); o[32 + c] = __uint_as_float(a1[c]); }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[qt & 1]);   // the next-but-one tile may overwrite this O buffer
        // normalise, stage this warp's 32 rows in ITS slice of the P buffer the next block will use (its last
        // reader, PV_{gl-1}, has completed and only this warp writes these rows), then write whole 128-byte
        // rows: lane = (row % 4, 16-byte chunk)
        uint8_t* stg = sP + ((gl + 1) & 1) * kAttQBytes;
        const float inv = 1.f / l_run;               // l == 0 (all keys masked) -> NaN, as torch
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint4 u;
          u.x = pack_bf16x2(o[8 * ch] * inv, o[8 * ch + 1] * inv);
          u.y = pack_bf16x2(o[8 * ch + 2] * inv, o[8 * ch + 3] * inv);
          u.z = pack_bf16x2(o[8 * ch + 4] * inv, o[8 * ch + 5] * inv);
          u.w = pack_bf16x2(o[8 * ch + 6] * inv, o[8 * ch + 7] * inv);
          *reinterpret_cast<uint4*>(stg + att_swz(row, ch)) = u;
        }
        __syncwarp();
        const int rr = lane >> 3, cc = lane & 7;
        bf16* ob = out + (static_cast<int64_t>(b) * Lq + q0 + quarter * 32) * ldo + h * 64 + cc * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = 4 * i + rr;
          if (q0 + quarter * 32 + rl < Lq)
            *reinterpret_cast<uint4*>(ob + static_cast<int64_t>(rl) * ldo) =
                *reinterpret_cast<const uint4*>(stg + att_swz(quarter * 32 + rl, cc));
        }
        __syncwarp();                                // the staged rows are read before the next block's P overwrites them
      } else {
        if (lane == 0) mbar_arrive(&o_free[qt & 1]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tr != nullptr && threadIdx.x == 0) tr[3] = clock64();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace tanb

using namespace tanb;

extern "C" int tan_attention_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                  int64_t ldv, const uint8_t* key_padding_mask, void* out, int64_t ldo, int B, int H,
                                  int Lq, int Lk, float* lse, void* stream) {
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
  TAN_CHECK(set_max_dyn_smem(reinterpret_cast<const void*>(attention_kernel), kAttSmem));
  // one CTA per (clip, head) when that fills the GPU (2 CTAs per SM); otherwise split the query tiles over
  // blockIdx.z so that short batches of long sequences still occupy every SM
  const int nq_total = (Lq + kAttBQ - 1) / kAttBQ;
  const int64_t want = 2ll * num_sms();           // two CTAs per SM; rounding DOWN: a full pipeline per CTA beats
  int split = static_cast<int>(want / (static_cast<int64_t>(B) * H));   // twice as many short CTAs (measured)
  if (split > nq_total) split = nq_total;
  if (split < 1) split = 1;
  const int tiles_per_cta = (nq_total + split - 1) / split;
  dim3 grid(H, B, (nq_total + tiles_per_cta - 1) / tiles_per_cta);
  return launch_pdl(attention_kernel, grid, dim3(kAttThreads), kAttSmem, static_cast<cudaStream_t>(stream), 1, tmQ,
                    tmK, tmV, key_padding_mask, static_cast<bf16*>(out), ldo, Lq, Lk, tiles_per_cta, lse,
                    (Lq + 63) / 64 * 64);
}
