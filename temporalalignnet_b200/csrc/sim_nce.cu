// Cosine-similarity matrix + MIL-NCE statistics.
//
//  * tan_sim_nce_fwd      fused mode (logits_out == NULL): sim_fused_kernel, a CTA-pair tcgen05 GEMM that keeps its
//                         256 video rows RESIDENT in shared memory (K = d <= 512) while it sweeps the text columns,
//                         so only the text operand streams from L2; the epilogue turns each accumulator tile into
//                         exp-sums per row (registers, carried across the sweep) and per column.  The
//                         [B*T x B*N] x S matrix never reaches HBM.
//                         store mode (logits_out != NULL): the generic pair GEMM (umma_gemm2.cuh) with SimEpi2,
//                         which also writes the bf16 logits once (TMA stores).
//  * tan_nce_from_logits  the same statistics from a materialised logits tensor: one coalesced streaming pass
//                         (HBM-bound), register column accumulators, butterfly row reductions.
//  * tan_pos_from_time    packed target bits from sentence start / end times (train/loss.py:26-41).
//  * tan_nce_reduce       rows/columns -> the four scalars of train/loss.py:248-256 (optional row / column
//                         selection masks for the thresholded loss, :277-304).
//
// Targets are a packed bitmask posbits[b][t][w] (bit n%32 of word n/32 = "sentence n of clip b is positive at
// frame t"), only for the LOCAL clips: positives live in the own-clip block of the matrix.  This covers the
// reference's interval targets and its self-labelled, de-duplicated targets (train/loss.py:88-229) alike and
// replaces the 2.1 GB [B,T,B,N] float target tensor.
//
// All sums use the fixed shift 1/0.07 (cosines are bounded by 1), so no running max is needed and
// partial sums from different tiles / ranks add directly:  e = exp((cos - 1)/0.07) in (0, 1].
#include <cstdlib>

#include "umma_gemm2.cuh"

namespace tanb {

constexpr float kInvTemp = 1.0f / 0.07f;                       // train/loss.py:66
constexpr float kExpScale = kInvTemp * 1.4426950408889634f;    // to the exp2 domain

struct SimCommon {
  tan_sim_geom g;
  const uint32_t* posbits;   // [B_loc][T][W]
  int W;                     // ceil(N / 32)
  const uint8_t* col_valid;  // [C]
  const uint8_t* row_kill;   // [B_loc * T] or NULL: 1 = this frame's own-clip entries count as exp(-inf)
  int seg_tiles;             // 256-row pair tiles per (clip, stage) segment = ceil(T / 256)
  int P;                     // column partials per stage
  int n_tiles;               // ceil(C / 256)
  int64_t R;                 // B_loc * S * T
};

// Columns [c0, c1) of local clip b_loc's own sentences (sentence n = column c0 + n).
__device__ __forceinline__ void own_columns(const tan_sim_geom& g, int b_loc, int& c0, int& c1) {
  if (g.col_off != nullptr) {
    c0 = __ldg(g.col_off + g.b_off + b_loc);
    c1 = __ldg(g.col_off + g.b_off + b_loc + 1);
  } else {
    c0 = (g.b_off + b_loc) * g.N;
    c1 = c0 + g.N;
  }
}

// 32 target bits of one frame for sentences n0 .. n0+31 of its clip (n0 in (-32, N); sentences < 0 give 0).
__device__ __forceinline__ uint32_t pos_bits32(const uint32_t* __restrict__ pw, int W, int n0) {
  if (n0 >= 0) {
    const int k = n0 >> 5;
    const uint32_t lo = pw[k];
    const uint32_t hi = (k + 1 < W) ? pw[k + 1] : 0u;
    return __funnelshift_r(lo, hi, n0 & 31);
  }
  return pw[0] << (-n0);
}

// Transposing butterfly: on entry lane l holds v[j] = value(row l, column j); on exit lane l holds
// the sum over the warp's 32 rows of column `l`.  31 shuffles.
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float keep = upper ? v[j + half] : v[j];
      const float send = upper ? v[j] : v[j + half];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

// One 32-column chunk of one accumulator row -> e[j] = exp((cos - 1)/0.07) with invalid columns / rows zeroed,
// own-clip entries of killed rows zeroed; returns the positive part in pe[] (caller checks has_pos).
struct ChunkCtx {
  bool row_ok, all_rows, kill, any_kill;
  int pos_c0, pos_c1, N, W;
  const uint32_t* pw;
};

__device__ __forceinline__ void chunk_exp(const uint32_t (&rc)[32], uint32_t okmask, const ChunkCtx& x, int col0,
                                          bool has_pos, float (&e)[32]) {
  if (x.all_rows && okmask == 0xffffffffu) {                 // common case: no masking at all
#pragma unroll
    for (int j = 0; j < 32; ++j) e[j] = fast_exp2(fmaf(__uint_as_float(rc[j]), kExpScale, -kExpScale));
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float v = fast_exp2(fmaf(__uint_as_float(rc[j]), kExpScale, -kExpScale));
      e[j] = (x.row_ok && ((okmask >> j) & 1u)) ? v : 0.f;
    }
  }
  if (has_pos && x.any_kill) {
    const int n0 = col0 - x.pos_c0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (x.kill && static_cast<unsigned>(j + n0) < static_cast<unsigned>(x.N)) e[j] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// Fused mode: A-resident CTA-pair GEMM.
//
// Task = one 256-frame row block of one (clip, stage) segment x one chunk of column tiles.  Each CTA of the
// pair keeps its 128 frames x d (<= 512) bf16 = up to 128 KB in shared memory for the whole task and streams
// only its half (128 rows) of every 256-column text tile through a 4-stage ring: 16 KB per 64-wide K block
// for a 128 x 256 x 64 MMA share, i.e. half the L2->SM bytes of the streaming kernel (which was measured
// operand-feed bound at ~7.3 TB/s chip-wide, profiles/r01b_prof_sim).
//
// Roles (384 threads, one CTA per SM, clusters of 2):
//   warp 0      TMA producer of the text ring (both CTAs; bytes credited to the leader's barriers)
//   warp 1      MMA issuer (leader CTA): tcgen05.mma.cta_group::2 M=256 N=256 K=16; commits free ring slots,
//               the A blocks (during a task's last tile, so the next task's rows reload under the MMAs) and
//               signal full accumulators
//   warp 2      TMEM allocator (2 accumulator stages x 256 columns)
//   warp 3      TMA producer of the resident video rows
//   warps 4-11  epilogue: warp (quarter, half) owns 32 frames x 128 columns of each tile
// ---------------------------------------------------------------------------------------------
constexpr int kSfThreads = 384;
constexpr int kSfStages = 4;
constexpr int kSfMaxKB = 8;
constexpr int kSfABytes = kSfMaxKB * kG2ABytes;             // 128 KB
constexpr int kSfRingBytes = kSfStages * kG2BBytes;         // 80 KB
constexpr int kSfColBytes = 2 * 4 * 2 * 256 * 4;            // [buffer][quarter][all,pos][256] fp32 = 16 KB
constexpr int kSfBiasBytes = 2 * 256 * 4;                   // [buffer][256] exponent bias of the tile's columns
constexpr int kSfSmem = kSfABytes + kSfRingBytes + kSfColBytes + kSfBiasBytes + 1024 /*barriers*/ + 1024 /*alignment*/;
static_assert(kSfSmem <= 232448, "shared memory budget exceeded");

struct SimFused {
  SimCommon c;
  int pair_m_tiles;        // B_loc * S * seg_tiles
  int64_t b_stage_rows;    // rows of the text operand per stage in its tensor map (0: shared by all stages)
  int col_chunks;          // tasks per row block
  int tiles_per_chunk;
  int num_kb;              // d / 64
  float* row_part;         // [2][2 * col_chunks][R]
  float* col_part;         // [2][S][P][C]
};

__global__ void __launch_bounds__(kSfThreads, 1)
sim_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SimFused p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kSfABytes;
  float* scol = reinterpret_cast<float*>(smem + kSfABytes + kSfRingBytes);
  float* sbias = reinterpret_cast<float*>(smem + kSfABytes + kSfRingBytes + kSfColBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSfABytes + kSfRingBytes + kSfColBytes + kSfBiasBytes);
  uint64_t* a_full = bars;                           // [8]  (leader's are used)
  uint64_t* a_empty = bars + kSfMaxKB;               // [8]
  uint64_t* b_full = bars + 2 * kSfMaxKB;            // [stages] (leader's are used)
  uint64_t* b_empty = b_full + kSfStages;            // [stages]
  uint64_t* tmem_full = b_empty + kSfStages;         // [2]
  uint64_t* tmem_empty = tmem_full + 2;              // [2]  (leader's are used)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tasks = p.pair_m_tiles * p.col_chunks;
  const int num_kb = p.num_kb;
  const SimCommon& c = p.c;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kSfMaxKB; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kSfStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * kG2EpiWarps); }
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

  if (warp == 3) {
    // ===== resident video rows (both CTAs) =====
    if (lane == 0) {
      uint32_t phase = 0;
      for (int task = pair_id; task < num_tasks; task += num_pairs) {
        const int pm = task / p.col_chunks;
        const int seg = pm / c.seg_tiles, i = pm % c.seg_tiles;
        const int a_row = seg * c.g.T + i * 256 + static_cast<int>(rank) * kG2BM;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_relaxed(&a_empty[kb], phase ^ 1);          // the previous task's MMAs have read this block
          const uint32_t full_leader = mapa_u32(smem_u32(&a_full[kb]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&a_full[kb], 2 * kG2ABytes);
          tma_load_2d_pair(smem_a + kb * kG2ABytes, &tmA, full_leader, kb * kG2BK, a_row);
        }
        phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    // ===== text ring (both CTAs) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int task = pair_id; task < num_tasks; task += num_pairs) {
        const int pm = task / p.col_chunks, ck = task % p.col_chunks;
        const int seg = pm / c.seg_tiles;
        const int tn0 = ck * p.tiles_per_chunk;
        const int tn1 = min(tn0 + p.tiles_per_chunk, c.n_tiles);
        const int b_base = static_cast<int>((seg % c.g.S) * p.b_stage_rows) + static_cast<int>(rank) * (kG2BN / 2);
        for (int tn = tn0; tn < tn1; ++tn) {
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait_relaxed(&b_empty[stage], phase ^ 1);
            const uint32_t full_leader = mapa_u32(smem_u32(&b_full[stage]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&b_full[stage], 2 * kG2BBytes);
            tma_load_2d_pair(smem_b + stage * kG2BBytes, &tmB, full_leader, kb * kG2BK, b_base + tn * kG2BN);
            if (++stage == kSfStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kG2BM, kG2BN);
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int task = pair_id; task < num_tasks; task += num_pairs) {
        const int ck = task % p.col_chunks;
        const int tn0 = ck * p.tiles_per_chunk;
        const int tn1 = min(tn0 + p.tiles_per_chunk, c.n_tiles);
        for (int tn = tn0; tn < tn1; ++tn) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kG2BN;
          for (int kb = 0; kb < num_kb; ++kb) {
            if (tn == tn0) mbar_wait(&a_full[kb], a_phase);
            mbar_wait(&b_full[stage], phase);
            tc_fence_after();
            if (TAN_MMA_LEADER()) {
              const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + kb * kG2ABytes));
              const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * kG2BBytes));
#pragma unroll
              for (int k = 0; k < kG2BK / 16; ++k)
                umma_bf16_ss_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
              tc_commit_pair(&b_empty[stage], 0x3);
              if (tn == tn1 - 1) tc_commit_pair(&a_empty[kb], 0x3);
              if (kb == num_kb - 1) tc_commit_pair(&tmem_full[acc], 0x3);
            }
            __syncwarp();
            if (++stage == kSfStages) { stage = 0; phase ^= 1; }
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        a_phase ^= 1;
      }
    }
  } else if (warp >= kG2EpiWarp0) {
    // ===== epilogue (both CTAs) =====
    const int ew = warp - kG2EpiWarp0;
    const int quarter = ew & 3, half = ew >> 2;
    const int tid = ew * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int buf = 0;
    auto sc = [&](int b, int q, int which, int col) -> float& { return scol[((b * 4 + q) * 2 + which) * 256 + col]; };
    // Exponent bias of a tile's 256 columns: -1/0.07 (log2 domain) for real sentences, -inf for padded ones and
    // the zero padding beyond C, so e = exp2(cos * k + bias) needs no per-element predicate.  Double buffered
    // like the column staging: the next tile's vector is written before the end-of-tile barrier.
    auto fill_bias = [&](int b, int tn) {
      const int col = tn * kG2BN + tid;
      sbias[b * 256 + tid] = (col < c.g.C && c.col_valid[col] != 0) ? -kExpScale : -INFINITY;
    };
    if (pair_id < num_tasks) fill_bias(0, (pair_id % p.col_chunks) * p.tiles_per_chunk);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int task = pair_id; task < num_tasks; task += num_pairs) {
      const int pm = task / p.col_chunks, ck = task % p.col_chunks;
      const int seg = pm / c.seg_tiles, i = pm % c.seg_tiles;
      const int tn0 = ck * p.tiles_per_chunk;
      const int tn1 = min(tn0 + p.tiles_per_chunk, c.n_tiles);
      const int b_loc = seg / c.g.S, s_idx = seg % c.g.S;
      const int t = i * 256 + static_cast<int>(rank) * 128 + quarter * 32 + lane;   // this thread's frame
      ChunkCtx x;
      x.row_ok = t < c.g.T;
      x.all_rows = __all_sync(0xffffffffu, x.row_ok);
      x.kill = c.row_kill != nullptr && x.row_ok && c.row_kill[b_loc * c.g.T + t] != 0;
      x.any_kill = __any_sync(0xffffffffu, x.kill);
      own_columns(c.g, b_loc, x.pos_c0, x.pos_c1);
      x.N = x.pos_c1 - x.pos_c0;
      x.W = c.W;
      x.pw = c.posbits + (static_cast<int64_t>(b_loc) * c.g.T + (x.row_ok ? t : 0)) * c.W;
      const int64_t r = static_cast<int64_t>(seg) * c.g.T + t;
      const int p_idx = (b_loc * c.seg_tiles + i) * 2 + static_cast<int>(rank);     // column partial of this CTA tile
      float row_all = 0.f, row_pos = 0.f;

      for (int tn = tn0; tn < tn1; ++tn) {
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const int n0 = tn * kG2BN + half * 128;
        const uint32_t taddr = tmem_base + acc * kG2BN + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
        uint32_t raw[2][32];
        tmem_ld_32x32(taddr, raw[0]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int col0 = n0 + ch * 32;
          tmem_ld_wait();
          if (ch + 1 < 4) {
            tmem_ld_32x32(taddr + (ch + 1) * 32, raw[(ch + 1) & 1]);
          } else {
            tc_fence_before();                       // the accumulator is in registers: hand it back to the MMA warp
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
          }
          const uint32_t(&rc)[32] = raw[ch & 1];
          const int cj = half * 128 + ch * 32 + lane;
          if (col0 >= c.g.C) {                       // warp-uniform: nothing but zero padding left
            sc(buf, quarter, 0, cj) = 0.f;
            sc(buf, quarter, 1, cj) = 0.f;
            continue;
          }
          const bool has_pos = col0 < x.pos_c1 && col0 + 32 > x.pos_c0;       // warp-uniform
          float e[32];
          if (x.all_rows && !(has_pos && x.any_kill)) {
            const float4* b4 = reinterpret_cast<const float4*>(sbias + buf * 256 + half * 128 + ch * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = b4[j];                 // same address in every lane: one broadcast read per 4 columns
              e[4 * j + 0] = fast_exp2(fmaf(__uint_as_float(rc[4 * j + 0]), kExpScale, bb.x));
              e[4 * j + 1] = fast_exp2(fmaf(__uint_as_float(rc[4 * j + 1]), kExpScale, bb.y));
              e[4 * j + 2] = fast_exp2(fmaf(__uint_as_float(rc[4 * j + 2]), kExpScale, bb.z));
              e[4 * j + 3] = fast_exp2(fmaf(__uint_as_float(rc[4 * j + 3]), kExpScale, bb.w));
            }
          } else {                                   // ragged last row block / killed frames: predicated path
            const int mycol = col0 + lane;
            const bool my_ok = mycol < c.g.C && c.col_valid[mycol] != 0;
            const uint32_t okmask = __ballot_sync(0xffffffffu, my_ok);
            chunk_exp(rc, okmask, x, col0, has_pos, e);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) row_all += e[j];
          float cpos = 0.f;
          if (has_pos) {
            const uint32_t bits = x.row_ok ? pos_bits32(x.pw, x.W, col0 - x.pos_c0) : 0u;
            float pe[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              pe[j] = ((bits >> j) & 1u) ? e[j] : 0.f;
              row_pos += pe[j];
            }
            cpos = warp_column_sums(pe, lane);
          }
          const float call = warp_column_sums(e, lane);
          sc(buf, quarter, 0, cj) = call;
          sc(buf, quarter, 1, cj) = cpos;
        }
        // combine the four quarters in a fixed order (deterministic) and publish this CTA tile's column partials;
        // the staging buffers alternate, so one barrier per tile orders writes against the previous reads
        {
          int tn_next = tn + 1;
          if (tn_next >= tn1) {
            const int task_next = task + num_pairs;
            tn_next = task_next < num_tasks ? (task_next % p.col_chunks) * p.tiles_per_chunk : -1;
          }
          if (tn_next >= 0) fill_bias(buf ^ 1, tn_next);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int j = tid; j < 2 * 256; j += 256) {
          const int which = j >> 8, cj = j & 255;
          const int col = tn * kG2BN + cj;
          if (col < c.g.C) {
            const float sum = ((sc(buf, 0, which, cj) + sc(buf, 1, which, cj)) + sc(buf, 2, which, cj)) + sc(buf, 3, which, cj);
            p.col_part[((static_cast<int64_t>(which) * c.g.S + s_idx) * c.P + p_idx) * c.g.C + col] = sum;
          }
        }
        buf ^= 1;
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (x.row_ok) {
        const int64_t part = static_cast<int64_t>(ck) * 2 + half;
        p.row_part[(static_cast<int64_t>(0) * 2 * p.col_chunks + part) * c.R + r] = row_all;
        p.row_part[(static_cast<int64_t>(1) * 2 * p.col_chunks + part) * c.R + r] = row_pos;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // nobody exits while the peer may still signal its barriers / read its smem
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 2 * kG2BN);
  }
}

// ---------------------------------------------------------------------------------------------
// Store mode: epilogue of the generic CTA-pair GEMM (umma_gemm2.cuh).  Pair tile = 256 frames of one
// (clip, stage) segment x 256 text columns; statistics as above plus the bf16 logits, staged as [32 x 64]
// boxes (128-byte swizzle) and written with TMA stores.
// ---------------------------------------------------------------------------------------------
struct SimEpi2 {
  static constexpr int kStages = 4;
  static constexpr int kWarpScratch = 2 * 4096 + 1024;    // two logits boxes + this warp's share of the column partials
  struct State {};
  SimCommon c;
  int pair_m_tiles;        // B_loc * S * seg_tiles
  int64_t b_stage_rows;    // rows of B per stage in the B tensor map (0 for the dual encoder)
  int store_logits;        // 1 TMA boxes through tmOut (3-D map [segment][T][C], needs C % 8 == 0), 2 direct
  bf16* logits;            // mode 2
  float* row_part;         // [2][2 * n_tiles][R]
  float* col_part;         // [2][S][P][C]

  __device__ __forceinline__ int num_tiles() const { return pair_m_tiles * c.n_tiles; }
  // column tiles fastest: the pair tiles in flight share their A rows and sweep B
  __device__ __forceinline__ PairTile coord(int tile) const {
    const int pm = tile / c.n_tiles, tn = tile % c.n_tiles;
    const int seg = pm / c.seg_tiles, i = pm % c.seg_tiles;
    PairTile pt;
    pt.a_row = seg * c.g.T + i * 256;
    pt.b_row = static_cast<int>((seg % c.g.S) * b_stage_rows) + tn * kG2BN;
    return pt;
  }

  __device__ __forceinline__ void pre(int, uint32_t, int, int, uint8_t*, float*, uint64_t*, uint32_t,
                                      const CUtensorMap*, const CUtensorMap*, State&) const {}

  __device__ __forceinline__ void run(int tile, uint32_t rank, uint32_t tmem_acc, int ew, int lane, uint8_t* ws,
                                      float*, uint64_t*, uint32_t, const CUtensorMap* tmOut, const CUtensorMap*,
                                      State&) const {
    const int quarter = ew & 3, half = ew >> 2;
    const int pm = tile / c.n_tiles, tn = tile % c.n_tiles;
    const int seg = pm / c.seg_tiles, i = pm % c.seg_tiles;
    const int b_loc = seg / c.g.S, s_idx = seg % c.g.S;
    const int t = i * 256 + static_cast<int>(rank) * 128 + quarter * 32 + lane;   // frame index of this thread's row
    ChunkCtx x;
    x.row_ok = t < c.g.T;
    x.all_rows = __all_sync(0xffffffffu, x.row_ok);
    x.kill = c.row_kill != nullptr && x.row_ok && c.row_kill[b_loc * c.g.T + t] != 0;
    x.any_kill = __any_sync(0xffffffffu, x.kill);
    own_columns(c.g, b_loc, x.pos_c0, x.pos_c1);
    x.N = x.pos_c1 - x.pos_c0;
    x.W = c.W;
    x.pw = c.posbits + (static_cast<int64_t>(b_loc) * c.g.T + (x.row_ok ? t : 0)) * c.W;
    const int64_t r = static_cast<int64_t>(seg) * c.g.T + t;
    const int n0 = tn * kG2BN + half * 128;                   // first column of this warp
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
    // column partials: [quarter][2 (all, pos)][256] floats spread over the eight 1 KB warp tails
    uint8_t* sbase = ws - ew * kWarpScratch;
    auto scol = [&](int q, int which, int col) -> float& {
      const int idx = (q * 2 + which) * 256 + col;            // 0 .. 2047
      return *reinterpret_cast<float*>(sbase + (idx >> 8) * kWarpScratch + 2 * 4096 + (idx & 255) * 4);
    };
    float row_all = 0.f, row_pos = 0.f;

    if (store_logits == 1 && lane == 0) tma_store_wait_read<0>();  // previous tile's boxes have been read out
    uint32_t raw[2][32];
    tmem_ld_32x32(taddr, raw[0]);
    if (store_logits == 1) __syncwarp();
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const int col0 = n0 + ch * 32;
      tmem_ld_wait();
      if (ch + 1 < 4) tmem_ld_32x32(taddr + (ch + 1) * 32, raw[(ch + 1) & 1]);
      const uint32_t(&rc)[32] = raw[ch & 1];
      if (col0 >= c.g.C) {                                    // warp-uniform: nothing but zero padding left
        scol(quarter, 0, half * 128 + ch * 32 + lane) = 0.f;
        scol(quarter, 1, half * 128 + ch * 32 + lane) = 0.f;
        continue;
      }
      const int mycol = col0 + lane;
      const bool my_ok = mycol < c.g.C && c.col_valid[mycol] != 0;
      const uint32_t okmask = __ballot_sync(0xffffffffu, my_ok);
      const bool has_pos = col0 < x.pos_c1 && col0 + 32 > x.pos_c0;   // warp-uniform
      if (store_logits == 2 && x.row_ok) {                      // generic tail path (C % 8 != 0): per-thread stores
        bf16* dst = logits + r * c.g.C + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < c.g.C) dst[j] = __float2bfloat16_rn(__uint_as_float(rc[j]));
      }
      if (store_logits == 1) {
        // 32 columns = 64 bytes = chunks [4 * (ch & 1), +4) of this row of the [32 x 64] bf16 box
        uint8_t* buf = ws + (ch >> 1) * 4096;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(rc[8 * j + 0]), __uint_as_float(rc[8 * j + 1]));
          u.y = pack_bf16x2(__uint_as_float(rc[8 * j + 2]), __uint_as_float(rc[8 * j + 3]));
          u.z = pack_bf16x2(__uint_as_float(rc[8 * j + 4]), __uint_as_float(rc[8 * j + 5]));
          u.w = pack_bf16x2(__uint_as_float(rc[8 * j + 6]), __uint_as_float(rc[8 * j + 7]));
          *reinterpret_cast<uint4*>(buf + lane * 128 + (((4 * (ch & 1) + j) ^ (lane & 7)) << 4)) = u;
        }
      }
      float e[32];
      chunk_exp(rc, okmask, x, col0, has_pos, e);
#pragma unroll
      for (int j = 0; j < 32; ++j) row_all += e[j];
      float cpos = 0.f;
      if (has_pos) {
        const uint32_t bits = x.row_ok ? pos_bits32(x.pw, x.W, col0 - x.pos_c0) : 0u;
        float pe[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          pe[j] = ((bits >> j) & 1u) ? e[j] : 0.f;
          row_pos += pe[j];
        }
        cpos = warp_column_sums(pe, lane);
      }
      const float call = warp_column_sums(e, lane);
      scol(quarter, 0, half * 128 + ch * 32 + lane) = call;
      scol(quarter, 1, half * 128 + ch * 32 + lane) = cpos;
    }
    if (store_logits == 1) {
      fence_proxy_async_smem();
      __syncwarp();
      const int t0 = i * 256 + static_cast<int>(rank) * 128 + quarter * 32;    // the box is clipped at T and C
      if (lane == 0 && t0 < c.g.T) {
#pragma unroll
        for (int pbox = 0; pbox < 2; ++pbox)
          if (n0 + 64 * pbox < c.g.C) tma_store_3d(tmOut, ws + pbox * 4096, n0 + 64 * pbox, t0, seg);
        tma_store_commit();
      }
    }
    if (x.row_ok) {
      const int64_t part = static_cast<int64_t>(tn) * 2 + half;
      row_part[(static_cast<int64_t>(0) * 2 * c.n_tiles + part) * c.R + r] = row_all;
      row_part[(static_cast<int64_t>(1) * 2 * c.n_tiles + part) * c.R + r] = row_pos;
    }
    // combine the four quarters in a fixed order (deterministic) and publish this CTA tile's column partials
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int p_idx = (b_loc * c.seg_tiles + i) * 2 + static_cast<int>(rank);
    const int tid = ew * 32 + lane;
    for (int j = tid; j < 2 * 256; j += 256) {
      const int which = j >> 8, cj = j & 255;
      const int col = tn * kG2BN + cj;
      if (col < c.g.C) {
        const float sum = ((scol(0, which, cj) + scol(1, which, cj)) + scol(2, which, cj)) + scol(3, which, cj);
        col_part[((static_cast<int64_t>(which) * c.g.S + s_idx) * c.P + p_idx) * c.g.C + col] = sum;
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // scratch is reused by the next tile
  }
};

// ---------------------------------------------------------------------------------------------
// Partial reduction (fixed order, deterministic).
//   blocks [0, row_blocks):  row_sums[w][r] = sum_p row_part[w][p][r]                 (p < row_parts)
//   the rest:                col_sums[w][s][c] = sum_p col_part[w][s][p][c]           (p < P)
// Column blocks are 32 columns x 8 slices of the partial range (128-byte coalesced loads, 8 loads in flight
// per thread), combined through shared memory in slice order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sim_reduce_partials_kernel(const float* __restrict__ row_part, const float* __restrict__ col_part, SimCommon c,
                           int row_parts, int row_blocks, float* __restrict__ row_sums, float* __restrict__ col_sums) {
  __shared__ float sm[8][32];
  pdl_launch_dependents();
  pdl_wait();
  if (static_cast<int>(blockIdx.x) < row_blocks) {
    const int64_t nrow = 2 * c.R;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < nrow; i += row_blocks * 256ll) {
      const int64_t w = i / c.R, r = i % c.R;
      const float* src = row_part + (w * row_parts) * c.R + r;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      int q = 0;
      for (; q + 4 <= row_parts; q += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += src[(q + u) * c.R];
      }
      for (; q < row_parts; ++q) acc[0] += src[q * c.R];
      row_sums[i] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    }
    return;
  }
  const int cblocks = (c.g.C + 31) / 32;
  const int cb = blockIdx.x - row_blocks;                 // (w * S + s) * cblocks + column block
  const int ws = cb / cblocks, col = (cb % cblocks) * 32 + (threadIdx.x & 31);
  const int slice = threadIdx.x >> 5;
  const int per = (c.P + 7) / 8;
  const int p0 = slice * per, p1 = min(p0 + per, c.P);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < c.g.C) {
    const float* src = col_part + (static_cast<int64_t>(ws) * c.P) * c.g.C + col;
    int q = p0;
    for (; q + 8 <= p1; q += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += src[static_cast<int64_t>(q + u) * c.g.C];
    }
    for (; q < p1; ++q) acc[0] += src[static_cast<int64_t>(q) * c.g.C];
  }
  sm[slice][threadIdx.x & 31] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  __syncthreads();
  if (slice == 0 && col < c.g.C) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sm[k][threadIdx.x];
    col_sums[static_cast<int64_t>(ws) * c.g.C + col] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Statistics from materialised logits (HBM-bound).
// CTA = 8 warps; a CTA owns (stage, 256-column slab, chunk of clips) and walks the frames of its clips.
// Each lane owns 8 consecutive columns (one 16-byte load of bf16, two of fp32), so a warp reads 512
// contiguous bytes per row; it takes 8 rows per iteration and issues all their loads before touching any of
// them (8 x 512 B in flight per warp).  Column sums live in registers for the whole walk (one partial per
// CTA); the 8 row sums of an iteration are reduced together by one transposing butterfly.  The common case
// (all 256 columns real sentences, 8 full rows, slab without own-clip columns) runs a lean loop of ~6
// instructions per element: the first version spent 19 (ncu r01b: issue-bound at 38 % of HBM peak).
// ---------------------------------------------------------------------------------------------
constexpr int kNceCols = 256;
constexpr int kNceWarps = 8;
constexpr int kNceRows = 8;

// v[i] (i < 8) per lane -> lane l returns sum over the warp of v[l & 7] (valid in every lane).
__device__ __forceinline__ float warp_sum8(float (&v)[8], int lane) {
#pragma unroll
  for (int half = 4; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float keep = upper ? v[j + half] : v[j];
      const float send = upper ? v[j] : v[j + half];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  float r = v[0];
  r += __shfl_xor_sync(0xffffffffu, r, 8);
  r += __shfl_xor_sync(0xffffffffu, r, 16);
  return r;
}

template <bool F32>
__device__ __forceinline__ void nce_unpack(const uint4 (&raw)[F32 ? 2 : 1], float (&x)[8]) {
  if (F32) {
    x[0] = __uint_as_float(raw[0].x); x[1] = __uint_as_float(raw[0].y);
    x[2] = __uint_as_float(raw[0].z); x[3] = __uint_as_float(raw[0].w);
    x[4] = __uint_as_float(raw[F32 ? 1 : 0].x); x[5] = __uint_as_float(raw[F32 ? 1 : 0].y);
    x[6] = __uint_as_float(raw[F32 ? 1 : 0].z); x[7] = __uint_as_float(raw[F32 ? 1 : 0].w);
  } else {
    // bf16 -> fp32 is a 16-bit shift (low half) or a mask (high half): one integer op per element
    x[0] = __uint_as_float(raw[0].x << 16); x[1] = __uint_as_float(raw[0].x & 0xffff0000u);
    x[2] = __uint_as_float(raw[0].y << 16); x[3] = __uint_as_float(raw[0].y & 0xffff0000u);
    x[4] = __uint_as_float(raw[0].z << 16); x[5] = __uint_as_float(raw[0].z & 0xffff0000u);
    x[6] = __uint_as_float(raw[0].w << 16); x[7] = __uint_as_float(raw[0].w & 0xffff0000u);
  }
}

template <bool F32>
__global__ void __launch_bounds__(kNceWarps * 32, 2)
nce_from_logits_kernel(const void* __restrict__ logits, SimCommon c, int clips_per_cta, float* __restrict__ row_part,
                       float* __restrict__ col_part) {
  __shared__ float scol[kNceWarps][2][kNceCols + 8];
  const int slab = blockIdx.x;                 // column slab
  const int s_idx = blockIdx.y;                // stage
  const int chunk = blockIdx.z;                // chunk of clips
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = slab * kNceCols + lane * 8;
  const bool aligned = (c.g.C % 8) == 0 && col0 + 8 <= c.g.C;
  pdl_launch_dependents();
  pdl_wait();

  uint32_t okbits = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = col0 + j;
    okbits |= (col < c.g.C && c.col_valid[col] != 0) ? (1u << j) : 0u;
  }
  // exponent bias per owned column: -1/0.07 (log2 domain) or -inf for padded sentences -> no predicates in the lean loop
  float bias[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bias[j] = ((okbits >> j) & 1u) ? -kExpScale : -INFINITY;
  const bool slab_aligned = __all_sync(0xffffffffu, aligned);
  float call[8], cpos[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { call[j] = 0.f; cpos[j] = 0.f; }

  const int b_begin = chunk * clips_per_cta, b_end = min(b_begin + clips_per_cta, c.g.B_loc);
  for (int b = b_begin; b < b_end; ++b) {
    const int seg = b * c.g.S + s_idx;
    int pos_c0, pos_c1;
    own_columns(c.g, b, pos_c0, pos_c1);
    const bool slab_has_pos = slab * kNceCols < pos_c1 && (slab + 1) * kNceCols > pos_c0;   // CTA-uniform
    const bool lane_has_pos = col0 < pos_c1 && col0 + 8 > pos_c0;
    const bool has_kill = c.row_kill != nullptr;
    const char* base = static_cast<const char*>(logits) +
                       (static_cast<int64_t>(seg) * c.g.T * c.g.C + col0) * (F32 ? 4 : 2);
    const int64_t pitch = static_cast<int64_t>(c.g.C) * (F32 ? 4 : 2);
    for (int t0 = warp * kNceRows; t0 < c.g.T; t0 += kNceWarps * kNceRows) {
      float ra[kNceRows], rp[kNceRows];
      if (slab_aligned && !slab_has_pos && t0 + kNceRows <= c.g.T) {
        // ---- lean path: full rows, no positives in this slab
        uint4 raw[kNceRows][F32 ? 2 : 1];
#pragma unroll
        for (int i = 0; i < kNceRows; ++i) {
          const uint4* q = reinterpret_cast<const uint4*>(base + (t0 + i) * pitch);
          raw[i][0] = __ldcs(q);
          if (F32) raw[i][F32 ? 1 : 0] = __ldcs(q + 1);
        }
#pragma unroll
        for (int i = 0; i < kNceRows; ++i) {
          float x[8];
          nce_unpack<F32>(raw[i], x);
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float e = fast_exp2(fmaf(x[j], kExpScale, bias[j]));
            call[j] += e;
            sum += e;
          }
          ra[i] = sum;
        }
        const float ra_sum = warp_sum8(ra, lane);
        if (lane < kNceRows) {
          const int64_t r = static_cast<int64_t>(seg) * c.g.T + t0 + lane;
          row_part[(static_cast<int64_t>(0) * gridDim.x + slab) * c.R + r] = ra_sum;
          row_part[(static_cast<int64_t>(1) * gridDim.x + slab) * c.R + r] = 0.f;
        }
        continue;
      }
      // ---- general path: column masks, ragged rows / columns, positives, killed rows
      uint4 raw[kNceRows][F32 ? 2 : 1];
#pragma unroll
      for (int i = 0; i < kNceRows; ++i) {
        const int t = t0 + i < c.g.T ? t0 + i : c.g.T - 1;
        if (aligned) {
          const uint4* q = reinterpret_cast<const uint4*>(base + t * pitch);
          raw[i][0] = __ldcs(q);
          if (F32) raw[i][F32 ? 1 : 0] = __ldcs(q + 1);
        } else {
          float xs[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            xs[j] = 0.f;
            if (col0 + j < c.g.C)
              xs[j] = F32 ? reinterpret_cast<const float*>(base + t * pitch)[j]
                          : __bfloat162float(reinterpret_cast<const bf16*>(base + t * pitch)[j]);
          }
          if (F32) {
            raw[i][0] = make_uint4(__float_as_uint(xs[0]), __float_as_uint(xs[1]), __float_as_uint(xs[2]), __float_as_uint(xs[3]));
            raw[i][F32 ? 1 : 0] = make_uint4(__float_as_uint(xs[4]), __float_as_uint(xs[5]), __float_as_uint(xs[6]), __float_as_uint(xs[7]));
          } else {
            raw[i][0] = make_uint4(pack_bf16x2(xs[0], xs[1]), pack_bf16x2(xs[2], xs[3]), pack_bf16x2(xs[4], xs[5]),
                                   pack_bf16x2(xs[6], xs[7]));
          }
        }
      }
#pragma unroll
      for (int i = 0; i < kNceRows; ++i) {
        float x[8];
        nce_unpack<F32>(raw[i], x);
        const int t = t0 + i;
        const bool row_ok = t < c.g.T;
        uint32_t bits = 0, killmask = 0;
        if (lane_has_pos && row_ok) {
          const int n0 = col0 - pos_c0;
          bits = pos_bits32(c.posbits + (static_cast<int64_t>(b) * c.g.T + t) * c.W, c.W, n0) & 0xffu;
          if (has_kill && c.row_kill[b * c.g.T + t] != 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              killmask |= (static_cast<unsigned>(j + n0) < static_cast<unsigned>(pos_c1 - pos_c0)) ? (1u << j) : 0u;
          }
        }
        const uint32_t live = row_ok ? (okbits & ~killmask) : 0u;
        ra[i] = 0.f;
        rp[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float e = ((live >> j) & 1u) ? fast_exp2(fmaf(x[j], kExpScale, -kExpScale)) : 0.f;
          const float pe = ((bits >> j) & 1u) ? e : 0.f;
          call[j] += e;
          ra[i] += e;
          cpos[j] += pe;
          rp[i] += pe;
        }
      }
      const float ra_sum = warp_sum8(ra, lane);
      float rp_sum = 0.f;
      if (slab_has_pos) rp_sum = warp_sum8(rp, lane);
      if (lane < kNceRows && t0 + lane < c.g.T) {
        const int64_t r = static_cast<int64_t>(seg) * c.g.T + t0 + lane;
        row_part[(static_cast<int64_t>(0) * gridDim.x + slab) * c.R + r] = ra_sum;
        row_part[(static_cast<int64_t>(1) * gridDim.x + slab) * c.R + r] = rp_sum;
      }
    }
  }
  // lane l holds columns 8l .. 8l+7: write them as two float4 per kind (row pitch padded by 8 floats)
  *reinterpret_cast<float4*>(&scol[warp][0][lane * 8]) = make_float4(call[0], call[1], call[2], call[3]);
  *reinterpret_cast<float4*>(&scol[warp][0][lane * 8 + 4]) = make_float4(call[4], call[5], call[6], call[7]);
  *reinterpret_cast<float4*>(&scol[warp][1][lane * 8]) = make_float4(cpos[0], cpos[1], cpos[2], cpos[3]);
  *reinterpret_cast<float4*>(&scol[warp][1][lane * 8 + 4]) = make_float4(cpos[4], cpos[5], cpos[6], cpos[7]);
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * kNceCols; j += blockDim.x) {
    const int which = j / kNceCols, cj = j % kNceCols;
    const int col = slab * kNceCols + cj;
    if (col < c.g.C) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kNceWarps; ++w) s += scol[w][which][cj];
      col_part[((static_cast<int64_t>(which) * c.g.S + s_idx) * c.P + chunk) * c.g.C + col] = s;
    }
  }
}

// posbits[b][t][w] from sentence times (train/loss.py:26-41): bit n = valid[b][n] && start[b][n] <= t < end[b][n]
__global__ void pos_from_time_kernel(const float* __restrict__ start, const float* __restrict__ end,
                                     const uint8_t* __restrict__ valid, int B, int T, int N, int W,
                                     uint32_t* __restrict__ posbits) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = static_cast<int64_t>(B) * T * W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(i % W);
    const int t = static_cast<int>((i / W) % T);
    const int b = static_cast<int>(i / (static_cast<int64_t>(W) * T));
    const float tf = static_cast<float>(t);
    uint32_t bits = 0;
    for (int k = 0; k < 32; ++k) {
      const int n = w * 32 + k;
      if (n < N && (valid == nullptr || valid[b * N + n] != 0) && start[b * N + n] <= tf && tf < end[b * N + n])
        bits |= 1u << k;
    }
    posbits[i] = bits;
  }
}

__global__ void nce_reduce_kernel(const float* __restrict__ row_sums, int64_t R, int S, int T,
                                  const uint8_t* __restrict__ row_sel, const float* __restrict__ col_sums, int64_t SC,
                                  int C, const uint8_t* __restrict__ col_sel, double* __restrict__ out) {
  // Deterministic although the blocks combine through atomics: every block's fp64 partial is rounded to a multiple
  // of 2^-20 before it is added, so all additions are EXACT in fp64 (|sum| < 2^22, <= 2^10 blocks: 52 bits) and
  // their order cannot change the result (the rounding moves the loss by < 1e-9 relative)
  double acc[4] = {0., 0., 0., 0.};
  pdl_launch_dependents();
  pdl_wait();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (row_sums != nullptr) {
    for (int64_t i = i0; i < R; i += stride) {
      const float all = row_sums[i], pos = row_sums[R + i];
      bool use = pos > 0.f;
      if (row_sel != nullptr) {                        // row i = (b * S + s) * T + t  ->  selection index b * T + t
        const int64_t t = i % T, b = i / (static_cast<int64_t>(S) * T);
        use = use && row_sel[b * T + t] != 0;
      }
      if (use) { acc[0] += static_cast<double>(logf(all) - logf(pos)); acc[1] += 1.; }
    }
  }
  if (col_sums != nullptr) {
    for (int64_t i = i0; i < SC; i += stride) {
      const float all = col_sums[i], pos = col_sums[SC + i];
      bool use = pos > 0.f;
      if (col_sel != nullptr) use = use && col_sel[i % C] != 0;
      if (use) { acc[2] += static_cast<double>(logf(all) - logf(pos)); acc[3] += 1.; }
    }
  }
  __shared__ double sred[4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sred[k][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double v = lane < (blockDim.x >> 5) ? sred[k][lane] : 0.;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((k & 1) == 0) v = rint(v * 1048576.0) * (1.0 / 1048576.0);
      if (lane == 0 && v != 0.) atomicAdd(out + k, v);
    }
  }
}

static int fill_common(SimCommon* c, const tan_sim_geom* g, const uint32_t* posbits, const uint8_t* col_valid,
                       const uint8_t* row_kill) {
  if (g == nullptr || posbits == nullptr || col_valid == nullptr)
    return set_error(TAN_ERR_ARG, "sim/nce: null geometry, target-bit or column-mask pointer");
  if (g->B_loc <= 0 || g->S <= 0 || g->T <= 0 || g->C <= 0 || g->N <= 0 || g->d <= 0 || g->b_off < 0 ||
      (g->col_off == nullptr && (g->C % g->N != 0 || (g->b_off + g->B_loc) * static_cast<int64_t>(g->N) > g->C)))
    return set_error(TAN_ERR_SHAPE, "sim/nce: bad geometry B_loc=%d S=%d T=%d C=%d N=%d d=%d b_off=%d", g->B_loc, g->S,
                     g->T, g->C, g->N, g->d, g->b_off);
  c->g = *g;
  c->posbits = posbits;
  c->W = (g->N + 31) / 32;
  c->col_valid = col_valid;
  c->row_kill = row_kill;
  c->seg_tiles = (g->T + 255) / 256;                       // pair tiles (256 frames) per segment
  c->P = g->B_loc * c->seg_tiles * 2;                      // one column partial per 128-row CTA tile of a stage
  c->n_tiles = (g->C + 255) / 256;                         // 256-column tiles / slabs
  c->R = static_cast<int64_t>(g->B_loc) * g->S * g->T;
  return TAN_OK;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

static size_t row_part_bytes(const tan_sim_geom* g) {
  const size_t R = static_cast<size_t>(g->B_loc) * g->S * g->T;
  const size_t n_tiles = (g->C + 255) / 256;
  return align256(2 * 2 * n_tiles * R * 4);
}

static int launch_reduce(const SimCommon& c, const float* row_part, const float* col_part, int row_parts,
                         float* row_sums, float* col_sums, cudaStream_t st) {
  int row_blocks = static_cast<int>((2 * c.R + 255) / 256);
  if (row_blocks > num_sms() * 8) row_blocks = num_sms() * 8;
  const int col_blocks = 2 * c.g.S * ((c.g.C + 31) / 32);
  return launch_pdl(sim_reduce_partials_kernel, dim3(row_blocks + col_blocks), dim3(256), 0, st, 1, row_part, col_part,
                    c, row_parts, row_blocks, row_sums, col_sums);
}

}  // namespace tanb

using namespace tanb;

extern "C" size_t tan_sim_nce_workspace_bytes(const tan_sim_geom* g) {
  if (g == nullptr || g->B_loc <= 0 || g->S <= 0 || g->T <= 0 || g->C <= 0) return 0;
  // one query covers every producer: row partials [2][2 * ceil(C/256)][R] (the fused and streaming kernels use
  // fewer), column partials [2][S][B_loc * 2 * ceil(T/256)][C] (the streaming kernel uses at most B_loc per stage)
  const size_t P = static_cast<size_t>(g->B_loc) * ((g->T + 255) / 256) * 2;
  return row_part_bytes(g) + align256(2 * static_cast<size_t>(g->S) * P * g->C * 4);
}

extern "C" int tan_pos_from_time(const float* start, const float* end, const uint8_t* valid, int B, int T, int N,
                                 uint32_t* posbits, void* stream) {
  TAN_CHECK(tan_device_check());
  if (start == nullptr || end == nullptr || posbits == nullptr)
    return set_error(TAN_ERR_ARG, "tan_pos_from_time: null pointer");
  if (B <= 0 || T <= 0 || N <= 0) return set_error(TAN_ERR_SHAPE, "tan_pos_from_time: bad dims B=%d T=%d N=%d", B, T, N);
  const int W = (N + 31) / 32;
  const int64_t total = static_cast<int64_t>(B) * T * W;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  return launch_pdl(pos_from_time_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, start, end,
                    valid, B, T, N, W, posbits);
}

extern "C" int tan_sim_nce_fwd(const void* vfeat, const void* tfeat, int64_t tfeat_stage_stride,
                               const tan_sim_geom* g, const uint32_t* posbits, const uint8_t* col_valid,
                               const uint8_t* row_kill, void* logits_out, float* row_sums, float* col_sums,
                               void* workspace, size_t workspace_bytes, void* stream) {
  TAN_CHECK(tan_device_check());
  if (vfeat == nullptr || tfeat == nullptr || row_sums == nullptr || col_sums == nullptr)
    return set_error(TAN_ERR_ARG, "tan_sim_nce_fwd: null pointer");
  if (g == nullptr) return set_error(TAN_ERR_ARG, "tan_sim_nce_fwd: null geometry");
  SimCommon c;
  TAN_CHECK(fill_common(&c, g, posbits, col_valid, row_kill));
  if (g->d % kG2BK != 0) return set_error(TAN_ERR_SHAPE, "tan_sim_nce_fwd: d %% 64 != 0 (d=%d)", g->d);
  if (tfeat_stage_stride != 0 && tfeat_stage_stride != static_cast<int64_t>(g->C) * g->d)
    return set_error(TAN_ERR_SHAPE, "tan_sim_nce_fwd: tfeat_stage_stride must be 0 or C*d");
  if (workspace == nullptr || workspace_bytes < tan_sim_nce_workspace_bytes(g))
    return set_error(TAN_ERR_WORKSPACE, "tan_sim_nce_fwd: workspace too small (%zu < %zu)", workspace_bytes,
                     tan_sim_nce_workspace_bytes(g));
  float* row_part = static_cast<float*>(workspace);
  float* col_part = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + row_part_bytes(g));
  const int64_t b_rows = tfeat_stage_stride == 0 ? g->C : static_cast<int64_t>(g->S) * g->C;
  const int pair_m_tiles = g->B_loc * g->S * c.seg_tiles;
  CUtensorMap tmA, tmB;
  TAN_CHECK(make_tmap_2d(&tmA, vfeat, 2, c.R, g->d, g->d, kG2BM));
  TAN_CHECK(make_tmap_2d(&tmB, tfeat, 2, b_rows, g->d, g->d, kG2BN / 2));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static const bool force_stream = getenv("TAN_SIM_STREAMING") != nullptr;   // debugging aid
  if (logits_out == nullptr && g->d <= kSfMaxKB * kG2BK && !force_stream) {
    // ---- fused mode: resident video rows.  Column chunks per row block: fewest "waves x (tiles + reload)", the
    // reload of the resident rows weighted as a fraction of a tile (it overlaps the previous task's last tile)
    // (0.25 measured best at 32 local clips x 8192 global columns: 0.396 -> 0.359 ms against a weight of 1)
    static const double reload_cost = getenv("TAN_SIM_RELOAD_COST") ? atof(getenv("TAN_SIM_RELOAD_COST")) : 0.25;
    const int max_pairs = num_sms() / 2;
    int best_k = 1;
    double best_cost = 1e300;
    for (int k = 1; k <= c.n_tiles && k <= 16; k *= 2) {
      const int tpc = (c.n_tiles + k - 1) / k;
      const int kk = (c.n_tiles + tpc - 1) / tpc;                 // chunks actually used
      const int64_t tasks = static_cast<int64_t>(pair_m_tiles) * kk;
      const int64_t pairs = tasks < max_pairs ? tasks : max_pairs;
      const double cost = static_cast<double>((tasks + pairs - 1) / pairs) * (tpc + reload_cost);
      if (cost < best_cost - 1e-9) { best_cost = cost; best_k = kk; }
    }
    SimFused f;
    f.c = c;
    f.pair_m_tiles = pair_m_tiles;
    f.b_stage_rows = tfeat_stage_stride == 0 ? 0 : g->C;
    f.tiles_per_chunk = (c.n_tiles + best_k - 1) / best_k;
    f.col_chunks = (c.n_tiles + f.tiles_per_chunk - 1) / f.tiles_per_chunk;
    f.num_kb = g->d / kG2BK;
    f.row_part = row_part;
    f.col_part = col_part;
    TAN_CHECK(set_max_dyn_smem(reinterpret_cast<const void*>(sim_fused_kernel), kSfSmem));
    const int64_t tasks = static_cast<int64_t>(pair_m_tiles) * f.col_chunks;
    const int pairs = static_cast<int>(tasks < max_pairs ? tasks : max_pairs);
    TAN_CHECK(launch_pdl(sim_fused_kernel, dim3(2 * pairs), dim3(kSfThreads), kSfSmem, st, 2, tmA, tmB, f));
    return launch_reduce(c, row_part, col_part, 2 * f.col_chunks, row_sums, col_sums, st);
  }
  SimEpi2 e;
  e.c = c;
  e.pair_m_tiles = pair_m_tiles;
  e.b_stage_rows = tfeat_stage_stride == 0 ? 0 : g->C;
  e.logits = static_cast<bf16*>(logits_out);
  e.store_logits = 0;
  CUtensorMap tmOut = tmA;
  if (logits_out != nullptr) {
    if (g->C % 8 == 0 && (reinterpret_cast<uintptr_t>(logits_out) & 15) == 0) {
      e.store_logits = 1;
      TAN_CHECK(make_tmap_3d_bf16(&tmOut, logits_out, static_cast<uint64_t>(g->B_loc) * g->S, g->T, g->C, 32));
    } else {
      e.store_logits = 2;
    }
  }
  e.row_part = row_part;
  e.col_part = col_part;
  TAN_CHECK(launch_umma_gemm2<SimEpi2>(tmA, tmB, tmOut, tmOut, e, e.pair_m_tiles * c.n_tiles, g->d / kG2BK, st));
  return launch_reduce(c, row_part, col_part, 2 * c.n_tiles, row_sums, col_sums, st);
}

extern "C" int tan_nce_from_logits(const void* logits, int logits_is_f32, const tan_sim_geom* g,
                                   const uint32_t* posbits, const uint8_t* col_valid, const uint8_t* row_kill,
                                   float* row_sums, float* col_sums, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  TAN_CHECK(tan_device_check());
  if (logits == nullptr || row_sums == nullptr || col_sums == nullptr)
    return set_error(TAN_ERR_ARG, "tan_nce_from_logits: null pointer");
  if (g == nullptr) return set_error(TAN_ERR_ARG, "tan_nce_from_logits: null geometry");
  SimCommon c;
  TAN_CHECK(fill_common(&c, g, posbits, col_valid, row_kill));
  if (workspace == nullptr || workspace_bytes < tan_sim_nce_workspace_bytes(g))
    return set_error(TAN_ERR_WORKSPACE, "tan_nce_from_logits: workspace too small (%zu < %zu)", workspace_bytes,
                     tan_sim_nce_workspace_bytes(g));
  if ((reinterpret_cast<uintptr_t>(logits) & 15) != 0)
    return set_error(TAN_ERR_SHAPE, "tan_nce_from_logits: logits must be 16-byte aligned");
  // clips per CTA: enough CTAs for ~8 per resident slot, at most one column partial per clip
  const int64_t base_ctas = static_cast<int64_t>(c.n_tiles) * g->S;
  int64_t want = (static_cast<int64_t>(num_sms()) * 2 * 8 + base_ctas - 1) / base_ctas;
  if (want < 1) want = 1;
  if (want > g->B_loc) want = g->B_loc;
  const int clips_per_cta = static_cast<int>((g->B_loc + want - 1) / want);
  const int chunks = (g->B_loc + clips_per_cta - 1) / clips_per_cta;
  if (chunks > 65535 || g->S > 65535) return set_error(TAN_ERR_SHAPE, "tan_nce_from_logits: grid too large");
  c.P = chunks;
  float* row_part = static_cast<float*>(workspace);
  float* col_part = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + row_part_bytes(g));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(c.n_tiles, g->S, chunks);
  if (logits_is_f32)
    TAN_CHECK(launch_pdl(nce_from_logits_kernel<true>, grid, dim3(kNceWarps * 32), 0, st, 1, logits, c, clips_per_cta,
                         row_part, col_part));
  else
    TAN_CHECK(launch_pdl(nce_from_logits_kernel<false>, grid, dim3(kNceWarps * 32), 0, st, 1, logits, c, clips_per_cta,
                         row_part, col_part));
  return launch_reduce(c, row_part, col_part, c.n_tiles, row_sums, col_sums, st);
}

extern "C" int tan_nce_reduce(const float* row_sums, int64_t R, int S, int T, const uint8_t* row_sel,
                              const float* col_sums, int64_t SC, int C, const uint8_t* col_sel, double* out,
                              void* stream) {
  TAN_CHECK(tan_device_check());
  if (out == nullptr || (row_sums == nullptr && col_sums == nullptr))
    return set_error(TAN_ERR_ARG, "tan_nce_reduce: null pointer");
  if ((row_sums != nullptr && (S <= 0 || T <= 0 || R % (static_cast<int64_t>(S) * T) != 0)) ||
      (col_sums != nullptr && (C <= 0 || SC % C != 0)))
    return set_error(TAN_ERR_SHAPE, "tan_nce_reduce: R must be B*S*T and SC must be S*C");
  const int64_t n = (row_sums ? R : 0) > (col_sums ? SC : 0) ? R : SC;
  int blocks = static_cast<int>((n + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
  return launch_pdl(nce_reduce_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, row_sums, R, S,
                    T, row_sel, col_sums, SC, C, col_sel, out);
}
