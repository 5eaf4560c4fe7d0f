// Persistent, warp-specialised tcgen05 GEMM main loop shared by the linear-layer kernel
// (gemm_linear.cu) and the similarity/NCE kernel (sim_nce.cu).
//
//   D[128 x BN] (fp32, TMEM) = A[128 x K] (bf16, K-major) * B[BN x K]^T (bf16, K-major)
//
// Roles (256 threads, one CTA per SM, grid = min(#tiles, #SMs), static round-robin tiles):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes [128 x 64] (A) and [BN x 64] (B),
//               128-byte swizzle, STAGES-deep ring guarded by full/empty mbarriers
//   warp 1      MMA issuer: one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//               4 per 64-wide K block; tcgen05.commit releases the smem slot / publishes the tile
//   warp 2      TMEM allocator (2 accumulator stages x BN fp32 columns) and deallocator
//   warps 4-7   epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> Epi functor; the next
//               tile's MMAs run meanwhile on the other accumulator stage
//
// HBM/L2 layout: A rows and B rows are both K-contiguous (nn.Linear weight layout needs no
// transpose).  K % 64 == 0.  M/N tails are zero-filled by TMA on load and masked by the Epi.
#pragma once

#include "common.cuh"

namespace tanb {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 256;
constexpr int kEpiWarp0 = 4;      // first epilogue warp
constexpr int kEpiThreads = 128;

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;            // 16 KB
  static constexpr int kBBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (192 * 1024) / kStageBytes;       // 4 (BN=256), 6 (128), 8 (64)
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kBarBytes = 1024;
  // + 1024 alignment slack for the 128B-swizzle atoms
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;
};

struct TileCoord {
  int a_row;   // TMA row coordinate of the A box
  int b_row;   // TMA row coordinate of the B box
};

// Epi concept:
//   static constexpr int kExtraSmem;                         bytes of epilogue scratch
//   __device__ int  num_tiles() const;
//   __device__ TileCoord coord(int tile) const;
//   __device__ void run(int tile, uint32_t tmem_acc, int quarter, int lane, uint8_t* scratch);
//        called by all 128 epilogue threads; must read its accumulator via tmem_ld_32x32 at
//        tmem_acc + (quarter*32 << 16) + column and finish with tmem_ld_wait() before returning.
template <int BN, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const Epi epi, const int num_kb) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                      // [STAGES]
  uint64_t* empty_bar = bars + STAGES;            // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;        // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;   // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint8_t* epi_scratch = smem + STAGES * Cfg::kStageBytes + Cfg::kBarBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = epi.num_tiles();

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
      mbar_init(&tmem_empty[i], kEpiThreads / 32);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_base_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord tc = epi.coord(tile);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * kGemmBK, tc.a_row);
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * kGemmBK, tc.b_row);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);      // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);            // TMA bytes have landed
        tc_fence_after();
        if (lane == 0) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the >>4 address field
            umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[stage]);                // smem slot reusable once these MMAs retire
          if (kb == num_kb - 1) tc_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue =====
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may address
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epi.run(tile, tmem_base + acc * BN, quarter, lane, epi_scratch);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace tanb
