// tan_gemm_tn_bf16: out[P, Q] (+)= A[R, P]^T @ B[R, Q]  (bf16 operands, fp32 accumulate / output), the contraction
// running over the ROWS of both row-major operands -- the shape of every weight gradient (dW = dY^T X, autograd of
// F.linear in model/tfm_model.py:21,35-37) and of the text-side similarity gradient (dB = G^T V, autograd of the
// einsum at model/tan_model.py:119,:139).  Round 1 fed these products to the K-major pair GEMM through explicit
// transposes of both operands (352 tan_transpose_bf16 launches, 53 GB and 12 ms per training step at the bench
// shape); here both operands are consumed AS THEY LIE in HBM: a TMA box of [64 rows x 64 columns] lands in shared
// memory as 64 contraction rows of 128 bytes, which is exactly the canonical MN-major 128-byte-swizzle UMMA layout
// (cute: Sw<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units), so the instruction descriptor only sets the
// a_major / b_major bits and the shared-memory descriptors carry LBO = 8 KB (next 64-wide column box), SBO = 1 KB.
//
// Same CTA-pair organisation as umma_gemm2.cuh (tcgen05.mma.cta_group::2, M = 256, N = 256, K = 16; 2 accumulator
// stages in TMEM; TMA producer / MMA issuer / 8 epilogue warps; persistent over tasks).  A weight gradient has few
// output tiles (4..16 of 256 x 256) and a very long contraction (all tokens), so the contraction is SPLIT inside
// the one launch: task = (row chunk, output tile); each task writes its fp32 tile to a partial buffer and a second
// kernel sums the partials in a fixed order (deterministic; no atomics).  Round 1 did this with up to 6 concurrent
// launches on side streams per weight gradient.
#include <algorithm>

#include "linear_epi.cuh"

namespace tanb {

namespace {

constexpr int kTnStages = 6;
constexpr int kTnBoxBytes = 64 * 128;                  // [64 contraction rows x 64 columns] bf16

// MN-major operand, 128-byte swizzle: 64-column boxes `lbo_bytes` apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

using TnEpi = LinearEpi2<kModeF32, TAN_ACT_NONE>;

struct TnGeom {
  int R;                  // contraction length (rows of A and B)
  int rows_per_split;     // multiple of 64
  int splits;
  int64_t split_stride;   // elements between the partial tiles of consecutive splits (0 when splits == 1)
};

__global__ void __launch_bounds__(kG2Threads, 1)
umma_gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TnEpi epi,
                    const TnGeom geo) {
  constexpr int STAGES = kTnStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                  // [STAGES][2 boxes]
  uint8_t* smem_b = smem + STAGES * kG2ABytes;             // [STAGES][2 boxes]
  uint8_t* scratch = smem + STAGES * kG2StageBytes;
  uint8_t* misc = scratch + kG2EpiWarps * TnEpi::kWarpScratch;
  float* colvec = reinterpret_cast<float*>(misc);
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc + 1024);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint64_t* rbar = bars + 2 * STAGES + 4;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(rbar + kG2EpiWarps * 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int tiles = epi.num_tiles();
  const int num_tasks = tiles * geo.splits;               // task = split * tiles + tile: concurrent tasks share rows

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kG2EpiWarps);
    }
    for (int i = 0; i < kG2EpiWarps * 4; ++i) mbar_init(&rbar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_base_slot, 2 * kG2BN);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_launch_dependents();
  pdl_wait();

  auto task_rows = [&](int task, int& r0, int& nkb) {
    const int s = task / tiles;
    r0 = s * geo.rows_per_split;
    const int r1 = min(geo.R, r0 + geo.rows_per_split);
    nkb = (r1 - r0 + kG2BK - 1) / kG2BK;
  };

  if (warp == 0) {
    // ===== TMA producer (both CTAs): 2 column boxes of A and of this CTA's half of B per stage =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair_id; t < num_tasks; t += num_pairs) {
        const PairTile pt = epi.coord(t % tiles);
        int r0, nkb;
        task_rows(t, r0, nkb);
        const int a_col = pt.a_row + static_cast<int>(rank) * kG2BM;         // output row = column of A
        const int b_col = pt.b_row + static_cast<int>(rank) * (kG2BN / 2);   // output column = column of B
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kG2StageBytes);
          const int row = r0 + kb * kG2BK;
          uint8_t* sa = smem_a + stage * kG2ABytes;
          uint8_t* sb = smem_b + stage * kG2BBytes;
          tma_load_2d_pair(sa, &tmA, full_leader, a_col, row);
          tma_load_2d_pair(sa + kTnBoxBytes, &tmA, full_leader, a_col + 64, row);
          tma_load_2d_pair(sb, &tmB, full_leader, b_col, row);
          tma_load_2d_pair(sb + kTnBoxBytes, &tmB, full_leader, b_col + 64, row);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kG2BM, kG2BN) | (1u << 15) | (1u << 16);   // A and B MN-major
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = pair_id; t < num_tasks; t += num_pairs) {
        int r0, nkb;
        task_rows(t, r0, nkb);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kG2BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (TAN_MMA_LEADER()) {
            const uint64_t da = umma_desc_mn_sw128(smem_u32(smem_a + stage * kG2ABytes), kTnBoxBytes);
            const uint64_t db = umma_desc_mn_sw128(smem_u32(smem_b + stage * kG2BBytes), kTnBoxBytes);
            // 16 contraction rows per MMA = 16 x 128 B = 2048 B further into each box
#pragma unroll
            for (int k = 0; k < kG2BK / 16; ++k)
              umma_bf16_ss_pair(tmem_d, da + 128 * k, db + 128 * k, idesc, (kb | k) != 0);
            tc_commit_pair(&empty_bar[stage], 0x3);
            if (kb == nkb - 1) tc_commit_pair(&tmem_full[acc], 0x3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= kG2EpiWarp0) {
    // ===== epilogue (both CTAs): fp32 tile (+ existing contents when accumulating in place) =====
    const int ew = warp - kG2EpiWarp0;
    uint8_t* wscratch = scratch + ew * TnEpi::kWarpScratch;
    uint64_t* wrbar = rbar + ew * 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t tile_phase = 0;
    for (int t = pair_id; t < num_tasks; t += num_pairs) {
      const int s = t / tiles, tile = t - s * tiles;
      TnEpi e = epi;
      e.out_f32 = epi.out_f32 + static_cast<int64_t>(s) * geo.split_stride;
      typename TnEpi::State st;
      e.pre(tile, rank, ew, lane, wscratch, colvec, wrbar, tile_phase, &tmA, &tmA, st);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      e.run(tile, rank, tmem_base + acc * kG2BN, ew, lane, wscratch, colvec, wrbar, tile_phase, &tmA, &tmA, st);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      tile_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 2 * kG2BN);
  }
}

// out[i] = (accumulate ? out[i] : 0) + sum_s partial[s][i], fixed order; row pitch of `out` may exceed Q.
__global__ void tn_reduce_kernel(const float4* __restrict__ partial, int splits, int64_t split_stride4, float* out,
                                 int64_t ldo, int P, int Q4, int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n4 = static_cast<int64_t>(P) * Q4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t p = i / Q4, q4 = i - p * Q4;
    float4* po = reinterpret_cast<float4*>(out + p * ldo) + q4;
    float4 a = accumulate ? *po : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 v = __ldg(partial + s * split_stride4 + i);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    *po = a;
  }
}

// Split heuristic: minimise waves(tiles * splits over the CTA pairs) x (rows per split + a per-task overhead of
// ~512 rows: accumulator drain + pipeline refill).  Deterministic function of the shape and the SM count.
struct TnPlan { int splits, rows_per_split; };
TnPlan tn_plan(int R, int P, int Q) {
  const int tiles = ((P + 255) / 256) * ((Q + 255) / 256);
  const int pairs = num_sms() / 2;
  const int kb_total = (R + 63) / 64;
  int best = 1;
  int64_t best_cost = -1;
  for (int s = 1; s <= 64 && s <= kb_total; ++s) {
    const int kb = (kb_total + s - 1) / s;
    if ((kb_total + kb - 1) / kb != s) continue;             // not a distinct partition
    if (s > 1 && kb < 16) break;                              // chunks shorter than 1024 rows: not worth a task
    const int64_t waves = (static_cast<int64_t>(tiles) * s + pairs - 1) / pairs;
    const int64_t cost = waves * (kb * 64 + 512);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  TnPlan p;
  p.rows_per_split = ((kb_total + best - 1) / best) * 64;
  p.splits = (R + p.rows_per_split - 1) / p.rows_per_split;
  return p;
}

}  // namespace

}  // namespace tanb

using namespace tanb;

extern "C" size_t tan_gemm_tn_workspace_bytes(int R, int P, int Q) {
  if (R <= 0 || P <= 0 || Q <= 0) return 0;
  const TnPlan p = tn_plan(R, P, Q);
  return p.splits > 1 ? static_cast<size_t>(p.splits) * P * Q * sizeof(float) : 0;
}

extern "C" int tan_gemm_tn_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int R, int P, int Q, float* out,
                                int64_t ldo, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  TAN_CHECK(tan_device_check());
  if (A == nullptr || B == nullptr || out == nullptr) return set_error(TAN_ERR_ARG, "tan_gemm_tn_bf16: null pointer");
  if (R <= 0 || P <= 0 || Q <= 0 || Q % 32 != 0 || P % 8 != 0)
    return set_error(TAN_ERR_SHAPE, "tan_gemm_tn_bf16: need R>0, P%%8==0, Q%%32==0 (R=%d P=%d Q=%d)", R, P, Q);
  if (lda % 8 != 0 || ldb % 8 != 0 || lda < P || ldb < Q || ldo % 4 != 0 || ldo < Q ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(TAN_ERR_SHAPE, "tan_gemm_tn_bf16: row pitches must cover the columns (lda, ldb multiples of 8, ldo "
                                    "of 4) and out must be 16-byte aligned");
  const TnPlan plan = tn_plan(R, P, Q);
  const size_t need = plan.splits > 1 ? static_cast<size_t>(plan.splits) * P * Q * sizeof(float) : 0;
  if (need > 0 && (workspace == nullptr || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15)))
    return set_error(TAN_ERR_WORKSPACE, "tan_gemm_tn_bf16: workspace of %zu bytes needed (16-byte aligned)", need);
  CUtensorMap tmA, tmB;
  TAN_CHECK(make_tmap_2d(&tmA, A, 2, R, P, lda, 64));
  TAN_CHECK(make_tmap_2d(&tmB, B, 2, R, Q, ldb, 64));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TnEpi e;
  e.M = P; e.N = Q; e.f_tiles = (Q + kG2BN - 1) / kG2BN; e.n_tiles = e.f_tiles * ((P + 2 * kG2BM - 1) / (2 * kG2BM));
  e.bias = nullptr; e.act = TAN_ACT_NONE; e.extra_bf16 = nullptr; e.ld_extra = 0;
  TnGeom geo;
  geo.R = R; geo.rows_per_split = plan.rows_per_split; geo.splits = plan.splits;
  if (plan.splits > 1) {
    e.residual = nullptr; e.ldr = 0; e.out_f32 = static_cast<float*>(workspace); e.ldo = Q;
    geo.split_stride = static_cast<int64_t>(P) * Q;
  } else {
    e.residual = accumulate ? out : nullptr; e.ldr = ldo; e.out_f32 = out; e.ldo = ldo;
    geo.split_stride = 0;
  }
  constexpr int smem = kTnStages * kG2StageBytes + kG2EpiWarps * TnEpi::kWarpScratch + kG2MiscBytes + 1024;
  static_assert(smem <= 232448, "shared memory budget exceeded");
  TAN_CHECK(set_max_dyn_smem(reinterpret_cast<const void*>(umma_gemm_tn_kernel), smem));
  const int max_pairs = num_sms() / 2;
  const int tasks = e.n_tiles * plan.splits;
  const int pairs = tasks < max_pairs ? tasks : max_pairs;
  TAN_CHECK(launch_pdl(umma_gemm_tn_kernel, dim3(2 * pairs, 1, 1), dim3(kG2Threads, 1, 1), smem, st, 2, tmA, tmB, e, geo));
  if (plan.splits > 1) {
    const int Q4 = Q / 4;
    const int64_t n4 = static_cast<int64_t>(P) * Q4;
    const int blocks = static_cast<int>(std::min<int64_t>((n4 + 255) / 256, 4 * num_sms()));
    TAN_CHECK(launch_pdl(tn_reduce_kernel, dim3(blocks), dim3(256), 0, st, 1, static_cast<const float4*>(workspace),
                         plan.splits, geo.split_stride / 4, out, ldo, P, Q4, accumulate));
  }
  return TAN_OK;
}
