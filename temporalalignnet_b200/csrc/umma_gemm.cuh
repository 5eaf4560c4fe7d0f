// Persistent, warp-specialised tcgen05 GEMM main loop shared by the linear-layer kernel
// (gemm_linear.cu) and the similarity/NCE kernel (sim_nce.cu).
//
//   D[128 x BN] (fp32, TMEM) = A[128 x K] (bf16, K-major) * B[BN x K]^T (bf16, K-major)
//
// Roles (256 threads, one CTA per SM, static round-robin tiles):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes [128 x 64] (A) and [BN/CS x 64] (B),
//               128-byte swizzle, STAGES-deep ring guarded by full/empty mbarriers
//   warp 1      MMA issuer: one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//               4 per 64-wide K block; tcgen05.commit releases the smem slot / publishes the tile
//   warp 2      TMEM allocator (2 accumulator stages x BN fp32 columns) and deallocator
//   warps 4-7   epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> Epi functor; the next
//               tile's MMAs run meanwhile on the other accumulator stage
//
// Thread-block clusters (CS = 1, 2 or 4 CTAs along M).  The CS CTAs of a cluster work on CS different
// 128-row blocks that share the same B tile.  Each CTA fetches only BN/CS rows of B and TMA-multicasts
// them into the same smem slot of every CTA in the cluster, so the B operand crosses the L2->SM fabric
// once per cluster instead of once per CTA (measured: these GEMMs are bound by L2->SM bytes, ~6-7 TB/s
// chip-wide, not by the tensor pipe; bytes per CTA per K block drop from 16+32 KB to 16+32/CS KB).
// A slot may be overwritten only when every CTA of the cluster has consumed it: tcgen05.commit
// multicasts its arrival to the empty barrier of all CS CTAs (count CS).
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

// Optional per-CTA event trace (development aid): tan_debug_set_trace() points this at a device buffer of
// 64 int64 slots per CTA; slot 0 = globaltimer at start, others = clock64 stamps (see TRACE() sites).
extern __device__ long long* g_gemm_trace;
__device__ __forceinline__ void trace_evt(long long* tr, int slot) {
  if (tr != nullptr && slot < 64) tr[blockIdx.x * 64 + slot] = clock64();
}

struct TileCoord {
  int a_row;   // TMA row coordinate of the A box
  int b_row;   // TMA row coordinate of the (full, BN-row) B tile; identical for all CTAs of a cluster
};

__device__ __forceinline__ void tma_load_2d_mcast(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int crd0,
                                                  int crd1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(crd0), "r"(crd1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// Epi concept:
//   static constexpr int kExtraSmem;                         bytes of epilogue scratch
//   struct State;                                            per-thread registers carried from pre() to run()
//   __device__ int  num_ctiles() const;                      number of cluster tiles (CS row blocks x one B tile)
//   __device__ int  tile_id(int ctile, int cta_rank) const;  this CTA's tile inside cluster tile `ctile`
//   __device__ TileCoord coord(int tile) const;
//   __device__ void init(uint8_t* scratch) const;            all 256 threads, before the role split
//   __device__ void pre(int tile, int quarter, int lane, uint8_t* scratch, State&) const;
//        epilogue threads, BEFORE waiting for the accumulator: issue loads that do not depend on it
//   __device__ void run(int tile, uint32_t tmem_acc, int quarter, int lane, uint8_t* scratch, State&) const;
//        called by all 128 epilogue threads; must read its accumulator via tmem_ld_32x32 at
//        tmem_acc + (quarter*32 << 16) + column and finish with tmem_ld_wait() before returning.
template <int BN, int CS, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const Epi epi, const int num_kb) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CS) - 1);
  constexpr int kBSliceRows = BN / CS;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B atoms (identical offset in every CTA of the cluster)
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
  const int cta_rank = (CS > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / CS;
  const int num_clusters = gridDim.x / CS;
  const int num_ctiles = epi.num_ctiles();
  long long* const tr = g_gemm_trace;
  if (tr != nullptr && threadIdx.x == 0) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tr[blockIdx.x * 64 + 0] = gt;
    trace_evt(tr, 1);
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], CS);
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
  epi.init(epi_scratch);
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();                 // peers' barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) trace_evt(tr, 2);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters, ++it) {
        const TileCoord tc = epi.coord(epi.tile_id(ct, cta_rank));
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);           // every CTA of the cluster has consumed this slot
          if (kb == 0) trace_evt(tr, 4 + it * 6 + 0);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * kGemmBK, tc.a_row);
          if (CS == 1) {
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * kGemmBK, tc.b_row);
          } else {
            tma_load_2d_mcast(smem_b + stage * Cfg::kBBytes + cta_rank * (kBSliceRows * 128), &tmB, &full_bar[stage],
                              kb * kGemmBK, tc.b_row + cta_rank * kBSliceRows, kMask);
          }
          if (kb == num_kb - 1) trace_evt(tr, 4 + it * 6 + 1);
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
    int it = 0;
    for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters, ++it) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);      // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);            // own A + all CS slices of B have landed
        tc_fence_after();
        if (lane == 0) {
          if (kb == 0) trace_evt(tr, 4 + it * 6 + 2);
          const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the >>4 address field
            umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          // slot reusable once these MMAs retire -- tell every CTA of the cluster
          if (CS == 1) tc_commit(&empty_bar[stage]);
          else tc_commit_mcast(&empty_bar[stage], kMask);
          if (kb == num_kb - 1) {
            tc_commit(&tmem_full[acc]);
            trace_evt(tr, 4 + it * 6 + 3);
          }
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
    int it = 0;
    for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters, ++it) {
      const int tile = epi.tile_id(ct, cta_rank);
      typename Epi::State st;
      epi.pre(tile, quarter, lane, epi_scratch, st);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (threadIdx.x == kEpiWarp0 * 32) trace_evt(tr, 4 + it * 6 + 4);
      epi.run(tile, tmem_base + acc * BN, quarter, lane, epi_scratch, st);
      tc_fence_before();
      __syncwarp();
      if (threadIdx.x == kEpiWarp0 * 32) trace_evt(tr, 4 + it * 6 + 5);
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();                 // no CTA exits while peers may still signal its barriers
  if (threadIdx.x == 0) trace_evt(tr, 3);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// Launch helper: cluster dimension CS along x; grid = clusters * CS.
template <int BN, int CS, class Epi>
int launch_umma_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const Epi& epi, int num_ctiles, int num_kb,
                     cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = umma_gemm_kernel<BN, CS, Epi>;
  constexpr int smem = Cfg::kSmemBytes + Epi::kExtraSmem;
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    TAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kGemmThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CS;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  // persistent grid = the number of clusters that can be co-resident (GPC sizes may strand a few SMs)
  static int max_clusters = 0;
  if (max_clusters == 0) {
    int n = 0;
    cfg.gridDim = dim3(num_sms() / CS * CS, 1, 1);
    if (CS == 1 || cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = num_sms() / CS;
    (void)cudaGetLastError();
    max_clusters = n < num_sms() / CS ? n : num_sms() / CS;
  }
  const int clusters = num_ctiles < max_clusters ? num_ctiles : max_clusters;
  cfg.gridDim = dim3(clusters * CS, 1, 1);
  TAN_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, epi, num_kb));
  return TAN_OK;
}

}  // namespace tanb
