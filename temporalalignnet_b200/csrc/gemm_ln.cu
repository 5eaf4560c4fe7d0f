// tan_linear_res_ln_bf16:  x <- x + A @ W^T + bias   (fp32 residual stream, in place)
//                          xn <- LayerNorm(x) * gamma + beta   (bf16)
// for a layer whose output width is 512 = the model width (attention out-projection + ln_2,
// model/tfm_model.py:32-37).  The unfused pair costs a GEMM that is HBM-bound on the fp32 residual stream (A 67 MB
// + x read 134 MB + x write 134 MB at 65 536 tokens: 83 us) plus a LayerNorm kernel that reads x again
// (134 MB) to write xn: 38 us.  Fused, x is read and written once: 402 MB instead of 536 MB and one launch.
//
// LayerNorm needs whole rows, so a CTA pair owns a WIDE tile: 256 tokens x all 512 features, i.e. 512 fp32
// accumulator columns per CTA = all of TMEM; the tile's MMAs therefore do not overlap the previous tile's
// epilogue, which is fine for a kernel bound by HBM.  Per 64-wide K block a CTA stages its 128 token rows and its
// halves (128 rows) of BOTH 256-feature weight tiles: 48 KB, 3 stages.
//   warp 0 TMA producer, warp 1 MMA issuer (two M=256 N=256 K=16 MMAs per K step), warp 2 TMEM allocator,
//   warps 4-11 epilogue: warp (quarter, half) owns 32 tokens x 256 features.
// Epilogue pass 1 (per 32-feature chunk): TMEM -> registers (+bias) -> 128-byte-swizzled smem box -> read back
// with lane = (row % 4, 16-byte column), add the fp32 residual with fully coalesced loads, store x, put the sum
// back into the box -> thread = row again: accumulate sum / sum of squares thread-locally (no shuffles) and
// write the updated row back to TMEM (tcgen05.st).  The two halves of a row exchange their partial sums through
// shared memory.  Pass 2: TMEM -> normalise -> bf16 box -> TMA store of xn.
#include "umma_gemm2.cuh"

namespace tanb {

constexpr int kLnThreads = 384;
constexpr int kLnN = 512;                                   // output width = LayerNorm width
constexpr int kLnStages = 3;
constexpr int kLnStageBytes = kG2ABytes + 2 * kG2BBytes;    // 48 KB
constexpr int kLnScratch = 4096;                            // per epilogue warp
constexpr int kLnVecBytes = 3 * kLnN * 4;                   // bias, gamma, beta
constexpr int kLnStatBytes = 2 * 128 * 2 * 4;               // [half][row][sum, sumsq]
constexpr int kLnSmem = kLnStages * kLnStageBytes + kG2EpiWarps * kLnScratch + kLnVecBytes + kLnStatBytes +
                        1024 /*barriers*/ + 1024 /*alignment*/;
static_assert(kLnSmem <= 232448, "shared memory budget exceeded");

struct ResLnArgs {
  int M, num_kb, n_tiles;
  const float* bias;
  const float* gamma;
  const float* beta;
  float* x;          // [M, 512] fp32, in/out
  int64_t ldx;
  int has_out;       // xn output requested
  int emit;          // L2-normalised stage features requested (tmNA / tmNB)
  int L, l_split;    // tokens per clip, first token of the "B" part (text tokens of the joint sequence)
  int64_t strideA, strideB;   // destination row = b * strideA + l  (l < l_split)  or  b * strideB + (l - l_split)
};

__device__ __forceinline__ uint32_t ln_swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__global__ void __launch_bounds__(kLnThreads, 1)
gemm_res_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmNA,
                   const __grid_constant__ CUtensorMap tmNB, const ResLnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = smem;
  uint8_t* scratch = smem + kLnStages * kLnStageBytes;
  float* vecs = reinterpret_cast<float*>(scratch + kG2EpiWarps * kLnScratch);       // bias | gamma | beta
  float* stats = vecs + 3 * kLnN;                                                    // [2][128][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stats) + kLnStatBytes);
  uint64_t* full_bar = bars;                    // [stages] (leader's are used)
  uint64_t* empty_bar = bars + kLnStages;       // [stages]
  uint64_t* tmem_full = bars + 2 * kLnStages;   // [1]
  uint64_t* tmem_empty = tmem_full + 1;         // [1] (leader's is used)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    tma_prefetch_desc(&tmNA);
    tma_prefetch_desc(&tmNB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kLnStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 2 * kG2EpiWarps);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_base_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair_id; t < p.n_tiles; t += num_pairs) {
        const int a_row = t * 2 * kG2BM + static_cast<int>(rank) * kG2BM;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kLnStageBytes);
          uint8_t* sa = stages + stage * kLnStageBytes;
          tma_load_2d_pair(sa, &tmA, full_leader, kb * kG2BK, a_row);
          tma_load_2d_pair(sa + kG2ABytes, &tmB, full_leader, kb * kG2BK, static_cast<int>(rank) * 128);
          tma_load_2d_pair(sa + kG2ABytes + kG2BBytes, &tmB, full_leader, kb * kG2BK, 256 + static_cast<int>(rank) * 128);
          if (++stage == kLnStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kG2BM, 256);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = pair_id; t < p.n_tiles; t += num_pairs) {
        mbar_wait(tmem_empty, acc_phase ^ 1);            // both CTAs' epilogues have drained the accumulators
        tc_fence_after();
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (TAN_MMA_LEADER()) {
            const uint32_t sa = smem_u32(stages + stage * kLnStageBytes);
            const uint64_t da = umma_desc_k_sw128(sa);
            const uint64_t db0 = umma_desc_k_sw128(sa + kG2ABytes);
            const uint64_t db1 = umma_desc_k_sw128(sa + kG2ABytes + kG2BBytes);
#pragma unroll
            for (int k = 0; k < kG2BK / 16; ++k) {
              umma_bf16_ss_pair(tmem_base, da + 2 * k, db0 + 2 * k, idesc, (kb | k) != 0);
              umma_bf16_ss_pair(tmem_base + 256, da + 2 * k, db1 + 2 * k, idesc, (kb | k) != 0);
            }
            tc_commit_pair(&empty_bar[stage], 0x3);
            if (kb == p.num_kb - 1) tc_commit_pair(tmem_full, 0x3);
          }
          __syncwarp();
          if (++stage == kLnStages) { stage = 0; phase ^= 1; }
        }
        acc_phase ^= 1;
      }
    }
  } else if (warp >= kG2EpiWarp0) {
    // ===== epilogue (both CTAs) =====
    const int ew = warp - kG2EpiWarp0;
    const int quarter = ew & 3, half = ew >> 2;
    const int tid = ew * 32 + lane;
    uint8_t* ws = scratch + ew * kLnScratch;
    // bias / gamma / beta are the same for every tile
    for (int i = tid; i < kLnN; i += 256) {
      vecs[i] = p.bias != nullptr ? __ldg(p.bias + i) : 0.f;
      vecs[kLnN + i] = __ldg(p.gamma + i);
      vecs[2 * kLnN + i] = __ldg(p.beta + i);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float4* bvec = reinterpret_cast<const float4*>(vecs + half * 256);
    const float4* gvec = reinterpret_cast<const float4*>(vecs + kLnN + half * 256);
    const float4* evec = reinterpret_cast<const float4*>(vecs + 2 * kLnN + half * 256);
    uint32_t acc_phase = 0;
    const int rr = lane >> 3, cc = lane & 7;
    // The residual rows of a tile are pulled into L2 while its MMAs run (the epilogue warps idle then; this
    // epilogue does not overlap the MMAs, so an HBM round trip per 32-feature chunk would be fully exposed;
    // prefetching a whole tile earlier was measured to be evicted again: 38 % L2 hits, ncu r01f): lane = row,
    // 8 lines of 128 bytes.
    auto prefetch_residual = [&](int t) {
      const int row = t * 2 * kG2BM + static_cast<int>(rank) * kG2BM + quarter * 32 + lane;
      if (t < p.n_tiles && row < p.M) {
        const char* px = reinterpret_cast<const char*>(p.x + static_cast<int64_t>(row) * p.ldx + half * 256);
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + i * 128));
      }
    };
    for (int t = pair_id; t < p.n_tiles; t += num_pairs) {
      prefetch_residual(t);
      const int row0 = t * 2 * kG2BM + static_cast<int>(rank) * kG2BM + quarter * 32;
      const int col0 = half * 256;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + col0;
      // residual of chunk 0 in the read-back mapping (coalesced: a warp instruction covers 4 rows x 128 bytes)
      float4 res[8];
      auto load_res = [&](int col) {
        const float* px = p.x + static_cast<int64_t>(row0 + rr) * p.ldx + col + 4 * cc;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          res[i] = (row0 + 4 * i + rr < p.M) ? *reinterpret_cast<const float4*>(px + 4 * i * p.ldx)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      load_res(col0);
      mbar_wait(tmem_full, acc_phase);
      tc_fence_after();

      // ---- pass 1: x <- x + acc + bias; row statistics; updated rows back to TMEM
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        // the residual of chunk c + 2 into L1 (lane = row, one 128-byte line): the register prefetch below is issued
        // one chunk ahead, ~0.6 of an iteration before its use, which does not cover an L2 round trip (ncu r02ad:
        // long_sb on the FADDs that consume the residual)
        if (c + 2 < 8 && row0 + lane < p.M)
          asm volatile("prefetch.global.L1 [%0];" ::"l"(p.x + static_cast<int64_t>(row0 + lane) * p.ldx + col0 + 32 * (c + 2)));
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = bvec[c * 8 + j];
          *reinterpret_cast<float4*>(ws + ln_swz(lane, j)) =
              make_float4(__uint_as_float(r[4 * j]) + b.x, __uint_as_float(r[4 * j + 1]) + b.y,
                          __uint_as_float(r[4 * j + 2]) + b.z, __uint_as_float(r[4 * j + 3]) + b.w);
        }
        __syncwarp();
        const int col = col0 + 32 * c;
        float* po = p.x + static_cast<int64_t>(row0 + rr) * p.ldx + col + 4 * cc;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 v = *reinterpret_cast<const float4*>(ws + ln_swz(4 * i + rr, cc));
          v.x += res[i].x; v.y += res[i].y; v.z += res[i].z; v.w += res[i].w;
          if (row0 + 4 * i + rr < p.M) *reinterpret_cast<float4*>(po + 4 * i * p.ldx) = v;
          *reinterpret_cast<float4*>(ws + ln_swz(4 * i + rr, cc)) = v;
        }
        if (c + 1 < 8) load_res(col + 32);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(ws + ln_swz(lane, j));
          s1 += (v.x + v.y) + (v.z + v.w);
          s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
          r[4 * j] = __float_as_uint(v.x); r[4 * j + 1] = __float_as_uint(v.y);
          r[4 * j + 2] = __float_as_uint(v.z); r[4 * j + 3] = __float_as_uint(v.w);
        }
        tmem_st_32x32(taddr + c * 32, r);
        __syncwarp();                 // the box is free for the next chunk
      }
      tmem_st_wait();
      // the two halves of a row combine their sums (fixed order)
      stats[(half * 128 + quarter * 32 + lane) * 2] = s1;
      stats[(half * 128 + quarter * 32 + lane) * 2 + 1] = s2;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int rloc = quarter * 32 + lane;
      const float t1 = stats[rloc * 2] + stats[(128 + rloc) * 2];
      const float t2 = stats[rloc * 2 + 1] + stats[(128 + rloc) * 2 + 1];
      const float mean = t1 * (1.0f / kLnN);
      const float var = fmaxf(t2 * (1.0f / kLnN) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);

      // ---- pass 2: y = (x - mean) * rstd * gamma + beta -> bf16 boxes -> TMA store of xn; with stage emission
      // also sum of squares of y (thread-local) and y back to TMEM for pass 3
      float q2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (c == 7 && !p.emit) {                       // the accumulators are in registers: hand TMEM back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(tmem_empty), 0));
        }
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = gvec[c * 8 + j], e = evec[c * 8 + j];
          const float a0 = (__uint_as_float(r[4 * j]) - mean) * rstd * g.x + e.x;
          const float a1 = (__uint_as_float(r[4 * j + 1]) - mean) * rstd * g.y + e.y;
          const float a2 = (__uint_as_float(r[4 * j + 2]) - mean) * rstd * g.z + e.z;
          const float a3 = (__uint_as_float(r[4 * j + 3]) - mean) * rstd * g.w + e.w;
          packed[2 * j] = pack_bf16x2(a0, a1);
          packed[2 * j + 1] = pack_bf16x2(a2, a3);
          if (p.emit) {
            q2 += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
            r[4 * j] = __float_as_uint(a0); r[4 * j + 1] = __float_as_uint(a1);
            r[4 * j + 2] = __float_as_uint(a2); r[4 * j + 3] = __float_as_uint(a3);
          }
        }
        if (p.emit) tmem_st_32x32(taddr + c * 32, r);
        if (p.has_out) {
          if ((c & 1) == 0) {                           // the box's previous TMA store has read it out
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(ws + ln_swz(lane, 4 * (c & 1) + j)) =
                make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          if (c & 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && row0 < p.M) {
              tma_store_2d(&tmOut, ws, col0 + 32 * (c - 1), row0);
              tma_store_commit();
            }
          }
        }
      }
      if (p.emit) {
        // ---- pass 3: stage feature = y / ||y||_2 (model/tan_model.py:116-117,:136-137) -> bf16, scattered by clip
        tmem_st_wait();
        asm volatile("bar.sync 1, 256;" ::: "memory");   // everyone has read the pass-1 sums
        stats[(half * 128 + rloc) * 2] = q2;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float inv = 1.0f / sqrtf(stats[rloc * 2] + stats[(128 + rloc) * 2]);
        const int b_clip = row0 / p.L, l0 = row0 - b_clip * p.L;          // the warp's 32 rows share clip and part
        const bool part_a = l0 < p.l_split;
        const int64_t dst_row = part_a ? b_clip * p.strideA + l0 : b_clip * p.strideB + (l0 - p.l_split);
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (c == 7) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(tmem_empty), 0));
          }
          if ((c & 1) == 0) {
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(r[8 * j]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
            *reinterpret_cast<uint4*>(ws + ln_swz(lane, 4 * (c & 1) + j)) = u;
          }
          if (c & 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && row0 < p.M) {
              tma_store_2d(part_a ? &tmNA : &tmNB, ws, col0 + 32 * (c - 1), static_cast<int>(dst_row));
              tma_store_commit();
            }
          }
        }
      }
      if (lane == 0) tma_store_wait_read<0>();          // pass 1 of the next tile reuses the box with plain stores
      __syncwarp();
      asm volatile("bar.sync 1, 256;" ::: "memory");     // stats are reused by the next tile
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace tanb

using namespace tanb;

static int launch_res_ln(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, float* x,
                         int64_t ldx, const float* gamma, const float* beta, void* out_bf16, int64_t ldo, int M,
                         int N, int K, int L, int l_split, void* nrmA, int64_t strideA, void* nrmB, int64_t strideB,
                         void* stream) {
  TAN_CHECK(tan_device_check());
  if (A == nullptr || W == nullptr || x == nullptr || gamma == nullptr || beta == nullptr)
    return set_error(TAN_ERR_ARG, "tan_linear_res_ln: null pointer");
  const bool emit = nrmA != nullptr || nrmB != nullptr;
  if (out_bf16 == nullptr && !emit) return set_error(TAN_ERR_ARG, "tan_linear_res_ln: no output requested");
  if (N != kLnN) return set_error(TAN_ERR_SHAPE, "tan_linear_res_ln: N must be 512 (N=%d)", N);
  if (M <= 0 || K <= 0 || K % kG2BK != 0)
    return set_error(TAN_ERR_SHAPE, "tan_linear_res_ln: need M>0, K%%64==0 (M=%d K=%d)", M, K);
  if (lda % 8 != 0 || ldw % 8 != 0 || lda < K || ldw < K || ldx % 4 != 0 || ldx < N ||
      (reinterpret_cast<uintptr_t>(x) & 15) ||
      (out_bf16 != nullptr && (ldo % 8 != 0 || ldo < N || (reinterpret_cast<uintptr_t>(out_bf16) & 15))))
    return set_error(TAN_ERR_SHAPE, "tan_linear_res_ln: row pitches / alignment");
  CUtensorMap tmA, tmB, tmOut, tmNA, tmNB;
  TAN_CHECK(make_tmap_2d(&tmA, A, 2, M, K, lda, kG2BM));
  TAN_CHECK(make_tmap_2d(&tmB, W, 2, N, K, ldw, 128));
  tmOut = tmA;
  if (out_bf16 != nullptr) TAN_CHECK(make_tmap_2d(&tmOut, out_bf16, 2, M, N, ldo, 32));
  tmNA = tmOut;
  tmNB = tmOut;
  ResLnArgs p;
  p.M = M;
  p.num_kb = K / kG2BK;
  p.n_tiles = (M + 2 * kG2BM - 1) / (2 * kG2BM);
  p.bias = bias;
  p.gamma = gamma;
  p.beta = beta;
  p.x = x;
  p.ldx = ldx;
  p.has_out = out_bf16 != nullptr;
  p.emit = emit;
  p.L = L > 0 ? L : M;
  p.l_split = l_split > 0 ? l_split : p.L;
  p.strideA = strideA;
  p.strideB = strideB;
  if (emit) {
    // a warp stores 32 consecutive tokens as one box: they must share their clip and their part
    if (M % p.L != 0 || p.L % 32 != 0 || p.l_split % 32 != 0 || p.l_split > p.L)
      return set_error(TAN_ERR_SHAPE, "tan_linear_res_ln: stage emission needs M %% L == 0 and L, l_split %% 32 == 0");
    if ((nrmA == nullptr) != (p.l_split == 0) && nrmA == nullptr)
      return set_error(TAN_ERR_ARG, "tan_linear_res_ln: nrmA missing");
    if (p.l_split < p.L && nrmB == nullptr) return set_error(TAN_ERR_ARG, "tan_linear_res_ln: nrmB missing");
    const int64_t clips = M / p.L;
    if (nrmA != nullptr) {
      if ((reinterpret_cast<uintptr_t>(nrmA) & 15) || strideA < p.l_split)
        return set_error(TAN_ERR_SHAPE, "tan_linear_res_ln: nrmA alignment / stride");
      TAN_CHECK(make_tmap_2d(&tmNA, nrmA, 2, (clips - 1) * strideA + p.l_split, N, N, 32));
    }
    if (nrmB != nullptr && p.l_split < p.L) {
      if ((reinterpret_cast<uintptr_t>(nrmB) & 15) || strideB < p.L - p.l_split)
        return set_error(TAN_ERR_SHAPE, "tan_linear_res_ln: nrmB alignment / stride");
      TAN_CHECK(make_tmap_2d(&tmNB, nrmB, 2, (clips - 1) * strideB + (p.L - p.l_split), N, N, 32));
    }
  }
  TAN_CHECK(set_max_dyn_smem(reinterpret_cast<const void*>(gemm_res_ln_kernel), kLnSmem));
  const int max_pairs = num_sms() / 2;
  const int pairs = p.n_tiles < max_pairs ? p.n_tiles : max_pairs;
  return launch_pdl(gemm_res_ln_kernel, dim3(2 * pairs), dim3(kLnThreads), kLnSmem, static_cast<cudaStream_t>(stream), 2,
                    tmA, tmB, tmOut, tmNA, tmNB, p);
}

extern "C" int tan_linear_res_ln_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                      float* x, int64_t ldx, const float* gamma, const float* beta, void* out_bf16,
                                      int64_t ldo, int M, int N, int K, void* stream) {
  if (out_bf16 == nullptr) return set_error(TAN_ERR_ARG, "tan_linear_res_ln_bf16: null output");
  return launch_res_ln(A, lda, W, ldw, bias, x, ldx, gamma, beta, out_bf16, ldo, M, N, K, 0, 0, nullptr, 0, nullptr, 0,
                       stream);
}

extern "C" int tan_linear_res_ln_stage_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                            float* x, int64_t ldx, const float* gamma, const float* beta,
                                            void* out_bf16, int64_t ldo, int M, int N, int K, int L, int l_split,
                                            void* nrmA_bf16, int64_t strideA, void* nrmB_bf16, int64_t strideB,
                                            void* stream) {
  return launch_res_ln(A, lda, W, ldw, bias, x, ldx, gamma, beta, out_bf16, ldo, M, N, K, L, l_split, nrmA_bf16, strideA,
                       nrmB_bf16, strideB, stream);
}
