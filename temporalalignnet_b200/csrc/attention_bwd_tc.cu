// Attention backward (head_dim 64) on tcgen05 tensor cores: score tiles in TMEM, operands fed by TMA, the same
// warp roles as the forward kernel (attention.cu).  Autograd of F.scaled_dot_product_attention reached from
// nn.MultiheadAttention at model/tfm_model.py:32.
//
// With P = softmax(Q K^T / 8 + mask), delta_i = <dO_i, O_i>:
//     dV = P^T dO            dS = P o (dO V^T - delta)           dQ = dS K / 8            dK = dS^T Q / 8
// Two kernels, no atomics (deterministic), both instances of ONE template.  A CTA keeps a 128-row tile of two
// operands resident in shared memory and streams 64-row blocks of the other two through a 2-stage TMA ring:
//
//   kernel dKV  resident K_t, V_t (128 keys), streams Q_g, dO_g (64 queries) + lse_g, delta_g (cp.async.bulk):
//               S^T = K_t Q_g^T, dP^T = V_t dO_g^T  (M=128 keys, N=64 queries)        -> TMEM
//               thread = key row:  P^T = exp2(S^T/8 log2e - lse_q),  dS^T = P^T (dP^T - delta_q)  -> smem, bf16
//               dV_t += P^T dO_g,  dK_t += dS^T Q_g   (dO_g / Q_g are the SAME shared-memory tiles, consumed
//               MN-major as B operands -- like V in the forward kernel; no transposed copies anywhere)
//   kernel dQ   resident Q_t, dO_t (128 queries), streams K_g, V_g (64 keys):
//               S = Q_t K_g^T, dP = dO_t V_g^T;  thread = query row (lse, delta thread-local):  dS -> smem
//               dQ_t += dS K_g  (K_g MN-major)
//
// The forward kernel stores the log2-domain log-sum-exp of every row ([B, H, pad64(Lq)], +inf in the padding), so
// nothing is recomputed here except the score tiles themselves; delta comes from a row-dot kernel.  Per 64-row
// block a CTA reads 2 x 32 KB of fp32 scores from TMEM (1024 clk at 64 B/clk) against 512 clk of MMAs and 512 clk
// of MUFU exp2: like the forward, TMEM read bandwidth is the floor; two CTAs per SM keep it busy.
//
//   warp 0      TMA producer (+ TMEM allocation)
//   warp 1      MMA issuer: first(g+1) as soon as the compute warps have read block g out of TMEM, second(g) as
//               soon as they have staged its bf16 tiles
//   warps 2-5   compute: thread = one row of the resident tile = TMEM lane
#include <algorithm>

#include "attn_bwd.cuh"

namespace tanb {

namespace {

constexpr int kBtRows = 128;                   // resident tile rows
constexpr int kBtBlk = 64;                     // streamed block rows
constexpr int kBtThreads = 192;
constexpr int kBtResBytes = kBtRows * 128;     // 16 KB per resident operand
constexpr int kBtBlkBytes = kBtBlk * 128;      // 8 KB per streamed operand
constexpr int kBtStatBytes = kBtBlk * 4;       // 256 B of lse / delta per block (kernel dKV)
constexpr int kBtStageBytes = 2 * kBtBlkBytes + 2 * kBtStatBytes;   // 16.5 KB, keeps 1024-byte alignment of the tiles
constexpr int kBtStages = 2;
constexpr int kBtMaskWords = 128;
// [resident x2][stages][P, dS tiles][barriers][mask bits]
constexpr int kBtSmem = 2 * kBtResBytes + kBtStages * 17408 + 2 * kBtResBytes + 256 + kBtMaskWords * 4;
static_assert(kBtStageBytes <= 17408, "stage slot");
static_assert(2 * (kBtSmem + 1024) <= 228 * 1024, "two CTAs per SM");

__device__ __forceinline__ uint32_t bt_swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// 1-D bulk copy global -> shared, completion bytes on an mbarrier (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct BtArgs {
  const uint8_t* kpm;      // [B, Lk] or null
  const float* lse;        // [B, H, Lp] log2 domain, +inf beyond Lq
  const float* delta;      // [B, H, Lp], 0 beyond Lq
  bf16* out0;              // kDKV: dv   | dQ kernel: dq
  int64_t ld0;
  bf16* out1;              // kDKV: dk   | unused
  int64_t ld1;
  int Lq, Lk, Lp, H;
};

// kDKV = true : resident X = K, Y = V (rows = keys);   streamed U = Q, W = dO (blocks of queries)
// kDKV = false: resident X = Q, Y = dO (rows = queries); streamed U = K, W = V (blocks of keys)
template <bool kDKV>
__global__ void __launch_bounds__(kBtThreads, 2)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                   const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmW, const BtArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* sX = smem;                               // 16 KB
  uint8_t* sY = sX + kBtResBytes;                   // 16 KB
  uint8_t* sStage = sY + kBtResBytes;               // [2][U 8 KB | W 8 KB | lse 256 B | delta 256 B | pad]
  uint8_t* sP = sStage + kBtStages * 17408;         // P^T (kDKV) 16 KB
  uint8_t* sDS = sP + kBtResBytes;                  // dS^T / dS 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + kBtResBytes);
  uint64_t* res_full = bars;          // [1]
  uint64_t* u_full = bars + 1;        // [2]
  uint64_t* u_empty = bars + 3;       // [2]
  uint64_t* s_full = bars + 5;        // [1] both score tiles of the block are in TMEM
  uint64_t* s_free = bars + 6;        // [1] count 4: the compute warps have read them
  uint64_t* p_ready = bars + 7;       // [1] count 4: the bf16 tiles are staged
  uint64_t* pv_done = bars + 8;       // [1] the block's accumulating MMAs have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(bars + 32);

  const int h = blockIdx.x, b = blockIdx.y;
  const int r0 = blockIdx.z * kBtRows;              // first resident row (key or query) inside the clip
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L_res = kDKV ? a.Lk : a.Lq;             // rows of the resident operands per clip
  const int L_str = kDKV ? a.Lq : a.Lk;             // rows of the streamed operands per clip
  const int nb = (L_str + kBtBlk - 1) / kBtBlk;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmX);
      tma_prefetch_desc(&tmY);
      tma_prefetch_desc(&tmU);
      tma_prefetch_desc(&tmW);
      mbar_init(res_full, 1);
      for (int i = 0; i < kBtStages; ++i) { mbar_init(&u_full[i], 1); mbar_init(&u_empty[i], 1); }
      mbar_init(s_full, 1);
      mbar_init(s_free, 4);
      mbar_init(p_ready, 4);
      mbar_init(pv_done, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);       // S' [0,64)  dP' [64,128)  acc0 [128,192)  acc1 [192,256)
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
      mbar_arrive_expect_tx(res_full, 2 * kBtResBytes);
      tma_load_2d(sX, &tmX, res_full, h * 64, b * L_res + r0);
      tma_load_2d(sY, &tmY, res_full, h * 64, b * L_res + r0);
      const float* lse_row = a.lse + (static_cast<int64_t>(b) * a.H + h) * a.Lp;
      const float* delta_row = a.delta + (static_cast<int64_t>(b) * a.H + h) * a.Lp;
      for (int g = 0; g < nb; ++g) {
        const int st = g % kBtStages;
        uint8_t* slot = sStage + st * 17408;
        mbar_wait_relaxed(&u_empty[st], ((g / kBtStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&u_full[st], kDKV ? kBtStageBytes : 2 * kBtBlkBytes);
        tma_load_2d(slot, &tmU, &u_full[st], h * 64, b * L_str + g * kBtBlk);
        tma_load_2d(slot + kBtBlkBytes, &tmW, &u_full[st], h * 64, b * L_str + g * kBtBlk);
        if (kDKV) {
          bulk_load_1d(slot + 2 * kBtBlkBytes, lse_row + g * kBtBlk, kBtStatBytes, &u_full[st]);
          bulk_load_1d(slot + 2 * kBtBlkBytes + kBtStatBytes, delta_row + g * kBtBlk, kBtStatBytes, &u_full[st]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc_s = umma_idesc_bf16(kBtRows, kBtBlk);                  // both operands K-major
    constexpr uint32_t idesc_acc = umma_idesc_bf16(kBtRows, 64) | (1u << 16);       // B operand MN-major
    auto issue_first = [&](int g) {
      const int st = g % kBtStages;
      mbar_wait(&u_full[st], (g / kBtStages) & 1);
      tc_fence_after();
      if (TAN_MMA_LEADER()) {
        const uint8_t* slot = sStage + st * 17408;
        const uint64_t dx = umma_desc_k_sw128(smem_u32(sX));
        const uint64_t dy = umma_desc_k_sw128(smem_u32(sY));
        const uint64_t du = umma_desc_k_sw128(smem_u32(slot));
        const uint64_t dw = umma_desc_k_sw128(smem_u32(slot + kBtBlkBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, dx + 2 * k, du + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + 64, dy + 2 * k, dw + 2 * k, idesc_s, k != 0);
        tc_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(res_full, 0);
    issue_first(0);
    for (int g = 0; g < nb; ++g) {
      const int st = g % kBtStages;
      if (g + 1 < nb) {
        mbar_wait(s_free, g & 1);                // block g has left TMEM
        issue_first(g + 1);
      }
      mbar_wait(p_ready, g & 1);                 // the bf16 tiles of block g are staged
      tc_fence_after();
      if (TAN_MMA_LEADER()) {
        const uint8_t* slot = sStage + st * 17408;
        const uint64_t du = umma_desc_k_sw128(smem_u32(slot));                 // as MN-major B: 16 rows = 2048 B per step
        const uint64_t dw = umma_desc_k_sw128(smem_u32(slot + kBtBlkBytes));
        const uint64_t dds = umma_desc_k_sw128(smem_u32(sDS));
        if (kDKV) {
          const uint64_t dp = umma_desc_k_sw128(smem_u32(sP));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + 128, dp + 2 * k, dw + 128 * k, idesc_acc, (g | k) != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + 192, dds + 2 * k, du + 128 * k, idesc_acc, (g | k) != 0);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + 128, dds + 2 * k, du + 128 * k, idesc_acc, (g | k) != 0);
        }
        tc_commit(pv_done);
        tc_commit(&u_empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ===== compute warps: thread = row of the resident tile = TMEM lane =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sl2 = 0.125f * 1.4426950408889634f;
    const uint8_t* mb = a.kpm != nullptr ? a.kpm + static_cast<int64_t>(b) * a.Lk : nullptr;
    float row_bias, row_delta = 0.f;     // exponent offset of the row; kDKV: 0 / -inf (masked key); dQ: -lse of the query
    if (kDKV) {
      const int key = r0 + row;
      const bool ok = key < a.Lk && !(mb != nullptr && mb[key] != 0);
      row_bias = ok ? 0.f : -INFINITY;
    } else {
      const int64_t idx = (static_cast<int64_t>(b) * a.H + h) * a.Lp + r0 + row;
      const bool ok = r0 + row < a.Lq;
      row_bias = ok ? -a.lse[idx] : -INFINITY;
      row_delta = ok ? a.delta[idx] : 0.f;
      // key mask of the clip as bits (bit set = ignore key)
      for (int wd = quarter; wd < nb * 2 && wd < kBtMaskWords; wd += 4) {
        const int key = wd * 32 + lane;
        const bool ig = key >= a.Lk || (mb != nullptr && mb[key] != 0);
        const uint32_t bits = __ballot_sync(0xffffffffu, ig);
        if (lane == 0) s_mask[wd] = bits;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    const bool mask_in_smem = nb * 2 <= kBtMaskWords;

    for (int g = 0; g < nb; ++g) {
      const int st = g % kBtStages;
      const uint8_t* slot = sStage + st * 17408;
      const float4* c_lse = reinterpret_cast<const float4*>(slot + 2 * kBtBlkBytes);
      const float4* c_delta = reinterpret_cast<const float4*>(slot + 2 * kBtBlkBytes + kBtStatBytes);
      uint32_t w[2] = {0u, 0u};
      if (!kDKV) {
        if (mask_in_smem) {
          w[0] = s_mask[2 * g];
          w[1] = s_mask[2 * g + 1];
        } else {
          const int k0 = g * kBtBlk + lane, k1 = k0 + 32;
          w[0] = __ballot_sync(0xffffffffu, k0 >= a.Lk || (mb != nullptr && mb[k0] != 0));
          w[1] = __ballot_sync(0xffffffffu, k1 >= a.Lk || (mb != nullptr && mb[k1] != 0));
        }
      }
      mbar_wait(s_full, g & 1);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t rs[32], rd[32];
        tmem_ld_32x32(t_lane + half * 32, rs);
        tmem_ld_32x32(t_lane + 64 + half * 32, rd);
        tmem_ld_wait();
        if (half == 1) {                       // both score tiles are in registers: the next block may overwrite them
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);
        }
        uint32_t pk_p[16], pk_ds[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {          // 4 columns per step
          float e[4], dl[4];
          if (kDKV) {
            const float4 l4 = c_lse[half * 8 + j], d4 = c_delta[half * 8 + j];
            e[0] = row_bias - l4.x; e[1] = row_bias - l4.y; e[2] = row_bias - l4.z; e[3] = row_bias - l4.w;
            dl[0] = d4.x; dl[1] = d4.y; dl[2] = d4.z; dl[3] = d4.w;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              e[i] = ((w[half] >> (4 * j + i)) & 1u) ? -INFINITY : row_bias;
              dl[i] = row_delta;
            }
          }
          float p[4], ds[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            p[i] = fast_exp2(fmaf(__uint_as_float(rs[4 * j + i]), sl2, e[i]));
            ds[i] = p[i] * (__uint_as_float(rd[4 * j + i]) - dl[i]);
          }
          pk_p[2 * j] = pack_bf16x2(p[0], p[1]);
          pk_p[2 * j + 1] = pack_bf16x2(p[2], p[3]);
          pk_ds[2 * j] = pack_bf16x2(ds[0], ds[1]);
          pk_ds[2 * j + 1] = pack_bf16x2(ds[2], ds[3]);
        }
        // the staging tiles were the A operands of the previous block's accumulating MMAs
        if (half == 0 && g >= 1) mbar_wait(pv_done, (g - 1) & 1);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          if (kDKV)
            *reinterpret_cast<uint4*>(sP + bt_swz(row, half * 4 + ch)) =
                make_uint4(pk_p[4 * ch], pk_p[4 * ch + 1], pk_p[4 * ch + 2], pk_p[4 * ch + 3]);
          *reinterpret_cast<uint4*>(sDS + bt_swz(row, half * 4 + ch)) =
              make_uint4(pk_ds[4 * ch], pk_ds[4 * ch + 1], pk_ds[4 * ch + 2], pk_ds[4 * ch + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }

    // ----- epilogue: accumulators -> bf16 -> this warp's 32 rows of the staging tile -> coalesced 128-byte rows -----
    mbar_wait(pv_done, (nb - 1) & 1);
    tc_fence_after();
    const int64_t grow0 = static_cast<int64_t>(b) * L_res + r0 + quarter * 32;     // first global row of this warp
    const int rows_left = L_res - (r0 + quarter * 32);                              // rows of this warp inside the clip
#pragma unroll
    for (int t = 0; t < (kDKV ? 2 : 1); ++t) {
      uint32_t a0[32], a1[32];
      tmem_ld_32x32(t_lane + 128 + t * 64, a0);
      tmem_ld_32x32(t_lane + 128 + t * 64 + 32, a1);
      tmem_ld_wait();
      const float sc = (kDKV && t == 0) ? 1.0f : 0.125f;       // dV unscaled; dK, dQ carry the softmax scale
      uint8_t* stg = (t == 0) ? sP : sDS;                       // both free: every MMA of the CTA has completed
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint32_t* src = ch < 4 ? a0 + 8 * ch : a1 + 8 * (ch - 4);
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(src[0]) * sc, __uint_as_float(src[1]) * sc);
        u.y = pack_bf16x2(__uint_as_float(src[2]) * sc, __uint_as_float(src[3]) * sc);
        u.z = pack_bf16x2(__uint_as_float(src[4]) * sc, __uint_as_float(src[5]) * sc);
        u.w = pack_bf16x2(__uint_as_float(src[6]) * sc, __uint_as_float(src[7]) * sc);
        *reinterpret_cast<uint4*>(stg + bt_swz(row, ch)) = u;
      }
      __syncwarp();
      bf16* outp = (t == 0) ? a.out0 : a.out1;
      const int64_t ldo = (t == 0) ? a.ld0 : a.ld1;
      const int rr = lane >> 3, cc = lane & 7;
      bf16* ob = outp + grow0 * ldo + h * 64 + cc * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = 4 * i + rr;
        if (rl < rows_left)
          *reinterpret_cast<uint4*>(ob + static_cast<int64_t>(rl) * ldo) =
              *reinterpret_cast<const uint4*>(stg + bt_swz(quarter * 32 + rl, cc));
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// delta[b, h, q] = <dO[q, h], O[q, h]> for q < Lq, 0 for Lq <= q < Lp.  One warp per (clip, query) row.
__global__ void attn_delta_kernel(const bf16* __restrict__ o, int64_t ldo, const bf16* __restrict__ dO, int64_t lddo,
                                  float* __restrict__ delta, int B, int H, int Lq, int Lp) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t wid = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int chunks = H * 8;                               // 16-byte chunks per row; 8 chunks per head
  for (int64_t r = wid; r < static_cast<int64_t>(B) * Lp; r += nw) {
    const int b = static_cast<int>(r / Lp), q = static_cast<int>(r - static_cast<int64_t>(b) * Lp);
    for (int c0 = 0; c0 < chunks; c0 += 32) {
      const int c = c0 + lane;
      float s = 0.f;
      if (q < Lq && c < chunks) {
        const int64_t row = static_cast<int64_t>(b) * Lq + q;
        const uint4 x = *reinterpret_cast<const uint4*>(o + row * ldo + c * 8);
        const uint4 y = *reinterpret_cast<const uint4*>(dO + row * lddo + c * 8);
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 p = unpack_bf16x2(xs[i]), t = unpack_bf16x2(ys[i]);
          s += p.x * t.x + p.y * t.y;
        }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if ((lane & 7) == 0 && c < chunks) delta[(static_cast<int64_t>(b) * H + (c >> 3)) * Lp + q] = s;
    }
  }
}

}  // namespace

int attention_bwd_tc(const AttnBwdArgs& a, cudaStream_t stream) {
  const int Lp = (a.Lq + 63) / 64 * 64;
  if (static_cast<int64_t>(a.B) * a.Lq > 0x7fffffffll || static_cast<int64_t>(a.B) * a.Lk > 0x7fffffffll)
    return set_error(TAN_ERR_SHAPE, "tan_attention_bwd_bf16: B*L exceeds the TMA coordinate range");
  if ((reinterpret_cast<uintptr_t>(a.lse) & 15) || (reinterpret_cast<uintptr_t>(a.delta) & 15))
    return set_error(TAN_ERR_SHAPE, "tan_attention_bwd_bf16: lse / delta must be 16-byte aligned");
  {
    const int64_t rows = static_cast<int64_t>(a.B) * Lp;
    const int blocks = static_cast<int>(std::min<int64_t>((rows + 7) / 8, 8ll * num_sms()));
    TAN_CHECK(launch_pdl(attn_delta_kernel, dim3(blocks), dim3(256), 0, stream, 1, a.o, a.ldo, a.dO, a.lddo, a.delta,
                         a.B, a.H, a.Lq, Lp));
  }
  const uint64_t rq = static_cast<uint64_t>(a.B) * a.Lq, rk = static_cast<uint64_t>(a.B) * a.Lk;
  const uint64_t cols = static_cast<uint64_t>(a.H) * 64;
  BtArgs g;
  g.kpm = a.kpm; g.lse = a.lse; g.delta = a.delta; g.Lq = a.Lq; g.Lk = a.Lk; g.Lp = Lp; g.H = a.H;
  TAN_CHECK(set_max_dyn_smem(reinterpret_cast<const void*>(attn_bwd_tc_kernel<true>), kBtSmem));
  TAN_CHECK(set_max_dyn_smem(reinterpret_cast<const void*>(attn_bwd_tc_kernel<false>), kBtSmem));
  {   // dK, dV: resident K, V tiles; streamed Q, dO blocks
    CUtensorMap tmX, tmY, tmU, tmW;
    TAN_CHECK(make_tmap_2d(&tmX, a.k, 2, rk, cols, a.ldk, kBtRows));
    TAN_CHECK(make_tmap_2d(&tmY, a.v, 2, rk, cols, a.ldv, kBtRows));
    TAN_CHECK(make_tmap_2d(&tmU, a.q, 2, rq, cols, a.ldq, kBtBlk));
    TAN_CHECK(make_tmap_2d(&tmW, a.dO, 2, rq, cols, a.lddo, kBtBlk));
    g.out0 = a.dv; g.ld0 = a.lddv; g.out1 = a.dk; g.ld1 = a.lddk;
    TAN_CHECK(launch_pdl(attn_bwd_tc_kernel<true>, dim3(a.H, a.B, (a.Lk + kBtRows - 1) / kBtRows), dim3(kBtThreads),
                         kBtSmem, stream, 1, tmX, tmY, tmU, tmW, g));
  }
  {   // dQ: resident Q, dO tiles; streamed K, V blocks
    CUtensorMap tmX, tmY, tmU, tmW;
    TAN_CHECK(make_tmap_2d(&tmX, a.q, 2, rq, cols, a.ldq, kBtRows));
    TAN_CHECK(make_tmap_2d(&tmY, a.dO, 2, rq, cols, a.lddo, kBtRows));
    TAN_CHECK(make_tmap_2d(&tmU, a.k, 2, rk, cols, a.ldk, kBtBlk));
    TAN_CHECK(make_tmap_2d(&tmW, a.v, 2, rk, cols, a.ldv, kBtBlk));
    g.out0 = a.dq; g.ld0 = a.lddq; g.out1 = nullptr; g.ld1 = 0;
    TAN_CHECK(launch_pdl(attn_bwd_tc_kernel<false>, dim3(a.H, a.B, (a.Lq + kBtRows - 1) / kBtRows), dim3(kBtThreads),
                         kBtSmem, stream, 1, tmX, tmY, tmU, tmW, g));
  }
  return TAN_OK;
}

}  // namespace tanb
