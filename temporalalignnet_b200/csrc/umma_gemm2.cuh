// CTA-pair (tcgen05 cta_group::2) persistent GEMM main loop with a shared-memory-staged, TMA-stored epilogue.
//
//   pair tile  D[256 x 256] (fp32, TMEM of both CTAs) = A[256 x K] (bf16, K-major) * B[256 x K]^T (bf16, K-major)
//
// Why pairs: these GEMMs (K = 512 .. 2048 at 8-9 k tokens) are bound by L2->SM operand bytes, not by the
// tensor pipe (measured ~6-7 TB/s chip-wide).  With cta_group::2 each CTA fetches its own 128 rows of A but
// only HALF of the B tile (128 of 256 rows); the pair's tensor cores read both halves, so a CTA moves
// 32 KB per 64-wide K block for a 128 x 256 x 64 MMA share instead of 48 KB (arithmetic intensity 128 vs 85
// flop/B).  TMA multicast inside larger clusters was measured neutral on this part; the pair is not.
//
// Roles (384 threads, one CTA per SM, clusters of 2 along x, static round-robin pair tiles):
//   warp 0      TMA producer: A box [128 x 64] + B-half box [128 x 64] per stage; completion bytes of BOTH
//               CTAs are credited to the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 1      MMA issuer (leader CTA only): tcgen05.mma.cta_group::2.kind::f16, M=256 N=256 K=16, 4 per stage;
//               tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs
//   warp 2      TMEM allocator: 2 accumulator stages x 256 fp32 columns (all 512 columns)
//   warp 3      idle
//   warps 4-11  epilogue: warp w owns TMEM lane quarter w%4 (32 token rows) and column half (w-4)/4
//               (128 features); tcgen05.ld -> registers -> Epi -> 128-byte-swizzled smem -> TMA store.
//               The next tile's MMAs run meanwhile on the other accumulator stage.
//
// Every global access happens after griddepcontrol.wait (PDL): the prologue overlaps the previous kernel.
#pragma once

#include "common.cuh"

namespace tanb {

constexpr int kG2Threads = 384;
constexpr int kG2EpiWarp0 = 4;
constexpr int kG2EpiWarps = 8;
constexpr int kG2BM = 128;            // rows of A per CTA (256 per pair)
constexpr int kG2BN = 256;            // columns per pair tile = accumulator columns per CTA
constexpr int kG2BK = 64;
constexpr int kG2ABytes = kG2BM * kG2BK * 2;          // 16 KB
constexpr int kG2BBytes = (kG2BN / 2) * kG2BK * 2;    // 16 KB (this CTA's half of the B tile)
constexpr int kG2StageBytes = kG2ABytes + kG2BBytes;  // 32 KB
constexpr int kG2MiscBytes = 2048;                    // [0,1024) per-tile fp32 column vector; [1024,2048) barriers

extern __device__ long long* g_gemm_trace;
__device__ __forceinline__ void trace_evt2(long long* tr, int slot) {
  if (tr != nullptr && slot < 64 && (slot >= 32 || slot < 28)) tr[blockIdx.x * 64 + slot] = clock64();
}

struct PairTile {
  int a_row;   // first A row of the PAIR tile (this CTA adds rank * 128)
  int b_row;   // first B row of the pair tile (this CTA adds rank * 128)
};

// Epi concept (see LinearEpi2 in gemm_linear.cu):
//   static constexpr int kStages;          smem ring depth
//   static constexpr int kWarpScratch;     bytes of 1024-aligned staging per epilogue warp
//   __device__ int num_tiles() const;
//   __device__ PairTile coord(int tile) const;
//   struct State;                          per-thread registers carried from pre() to run()
//   __device__ void pre(tile, rank, ew, lane, warp_scratch, colvec, rbar, phase, tmOut, tmAux, State&)
//        before the accumulator is awaited: free the staging buffers, start auxiliary loads, fill colvec
//   __device__ void run(tile, rank, tmem_acc, ew, lane, warp_scratch, colvec, rbar, phase, tmOut, tmAux, State&)
//        read the accumulator with tcgen05.ld (finish with tmem_ld_wait()), stage and TMA-store the result
template <class Epi>
__global__ void __launch_bounds__(kG2Threads, 1)
umma_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux,
                  const Epi epi, const int num_kb) {
  constexpr int STAGES = Epi::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * kG2ABytes;
  uint8_t* scratch = smem + STAGES * kG2StageBytes;
  uint8_t* misc = scratch + kG2EpiWarps * Epi::kWarpScratch;
  float* colvec = reinterpret_cast<float*>(misc);
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc + 1024);
  uint64_t* full_bar = bars;                      // [STAGES]   (leader's are used)
  uint64_t* empty_bar = bars + STAGES;            // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;        // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;   // [2]        (leader's are used)
  uint64_t* rbar = bars + 2 * STAGES + 4;         // [8 warps][4] auxiliary-load barriers of the epilogue
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(rbar + kG2EpiWarps * 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();        // 0 = leader
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = epi.num_tiles();
  long long* const tr = g_gemm_trace;
  if (tr != nullptr && threadIdx.x == 0) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tr[blockIdx.x * 64 + 0] = gt;
    trace_evt2(tr, 1);
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    tma_prefetch_desc(&tmAux);
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
  cluster_sync_all();                  // the peer's barriers exist before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_launch_dependents();             // the next kernel may start its own prologue
  pdl_wait();                          // ... and this one may now touch global memory
  if (threadIdx.x == 0) trace_evt2(tr, 2);

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair_id; t < num_tiles; t += num_pairs, ++it) {
        const PairTile pt = epi.coord(t);
        const int a_row = pt.a_row + static_cast<int>(rank) * kG2BM;
        const int b_row = pt.b_row + static_cast<int>(rank) * (kG2BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
          if (kb == 0) trace_evt2(it < 3 ? tr : nullptr, 4 + it * 8 + 0);
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kG2StageBytes);
          tma_load_2d_pair(smem_a + stage * kG2ABytes, &tmA, full_leader, kb * kG2BK, a_row);
          tma_load_2d_pair(smem_b + stage * kG2BBytes, &tmB, full_leader, kb * kG2BK, b_row);
          if (kb == num_kb - 1) trace_evt2(tr, 4 + it * 8 + 1);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kG2BM, kG2BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int it = 0;
      for (int t = pair_id; t < num_tiles; t += num_pairs, ++it) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);      // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kG2BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);            // A and B-half of BOTH CTAs have landed
          tc_fence_after();
          if (TAN_MMA_LEADER()) {
            if (kb == 0) trace_evt2(tr, 4 + it * 8 + 2);
            const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + stage * kG2ABytes));
            const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * kG2BBytes));
#pragma unroll
            for (int k = 0; k < kG2BK / 16; ++k)
              umma_bf16_ss_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            tc_commit_pair(&empty_bar[stage], 0x3);      // slot reusable in both CTAs once these MMAs retire
            if (kb == num_kb - 1) {
              tc_commit_pair(&tmem_full[acc], 0x3);      // accumulator complete, both CTAs
              trace_evt2(tr, 4 + it * 8 + 3);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= kG2EpiWarp0) {
    // ===== epilogue (both CTAs) =====
    const int ew = warp - kG2EpiWarp0;
    uint8_t* wscratch = scratch + ew * Epi::kWarpScratch;
    uint64_t* wrbar = rbar + ew * 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t tile_phase = 0;
    int it = 0;
    for (int t = pair_id; t < num_tiles; t += num_pairs, ++it) {
      long long* const trw = (threadIdx.x == kG2EpiWarp0 * 32) ? tr : nullptr;   // warp 4 lane 0 records
      trace_evt2(trw, 4 + it * 8 + 4);
      typename Epi::State st;
      epi.pre(t, rank, ew, lane, wscratch, colvec, wrbar, tile_phase, &tmOut, &tmAux, st);
      trace_evt2(trw, 4 + it * 8 + 5);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      trace_evt2(trw, 4 + it * 8 + 6);
      epi.run(t, rank, tmem_base + acc * kG2BN, ew, lane, wscratch, colvec, wrbar, tile_phase, &tmOut, &tmAux, st);
      tc_fence_before();
      __syncwarp();
      trace_evt2(trw, 4 + it * 8 + 7);
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      tile_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_read<0>();      // staging smem has been read; the writes complete with the grid
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // nobody exits while the peer may still signal its barriers / read its smem
  if (threadIdx.x == 0) trace_evt2(tr, 3);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 2 * kG2BN);
  }
}

template <class Epi>
constexpr int gemm2_smem_bytes() {
  return Epi::kStages * kG2StageBytes + kG2EpiWarps * Epi::kWarpScratch + kG2MiscBytes + 1024;
}

// grid = 2 * min(num_tiles, #SM / 2) CTAs in clusters of 2.
template <class Epi>
int launch_umma_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                      const CUtensorMap& tmAux, const Epi& epi, int num_tiles, int num_kb, cudaStream_t stream) {
  auto kern = umma_gemm2_kernel<Epi>;
  constexpr int smem = gemm2_smem_bytes<Epi>();
  static_assert(smem <= 232448, "shared memory budget exceeded");
  TAN_CHECK(set_max_dyn_smem(reinterpret_cast<const void*>(kern), smem));
  const int max_pairs = num_sms() / 2;
  const int pairs = num_tiles < max_pairs ? num_tiles : max_pairs;
  return launch_pdl(kern, dim3(2 * pairs, 1, 1), dim3(kG2Threads, 1, 1), smem, stream, 2, tmA, tmB, tmOut, tmAux,
                    epi, num_kb);
}

// Measured alternatives that did NOT pay off for the linear layers (same box, M = 65536, K = 512, N = 1536 / 2048):
// keeping the pair's weight tile resident in shared memory (halves the TMA fill traffic; 1.08 PFLOP/s, equal to
// this kernel with a 6-stage ring: what the fills save is lost to one-wave quantisation) and keeping the token
// rows resident (reload stalls at every task boundary).  The fused similarity kernel (sim_nce.cu), with 32
// column tiles per resident row block, is where residency pays (0.93 of the measured bf16 peak).

}  // namespace tanb
