// Cosine-similarity matrix + MIL-NCE statistics.
//
//  * tan_sim_nce_fwd      fused: tcgen05 GEMM (umma_gemm.cuh) whose epilogue turns each 128 x BN
//                         accumulator tile into exp-sums per row and per column (all / positive),
//                         optionally storing the bf16 logits once.  In fused mode the
//                         [B*T x B*N] x S matrix never reaches HBM.
//  * tan_nce_from_logits  the same statistics from a materialised logits tensor: one coalesced
//                         streaming pass (HBM-bound), warp-shuffle row reductions, register column
//                         accumulators.
//  * tan_nce_reduce       rows/columns -> the four scalars of train/loss.py:248-256.
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
  const float* start;
  const float* end;
  const uint8_t* col_valid;
  int seg_tiles;     // row tiles per (clip, stage) segment = ceil(T / 128)
  int m_tiles;       // B_loc * S * seg_tiles
  int n_tiles;       // ceil(C / BN)
  int64_t R;         // B_loc * S * T
};

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

// Epilogue of the CTA-pair GEMM (umma_gemm2.cuh) for the similarity matrix.  Pair tile = 256 frames of one
// (clip, stage) segment x 256 text columns; this CTA owns 128 of the frames.  Epilogue warp (quarter, half)
// turns 32 frames x 128 columns of cosines into exp-sums: per frame (thread-local over its columns, one
// partial per (column tile, half)) and per column (transposing butterfly over the warp's 32 frames, then a
// fixed-order combine of the four quarters through shared memory: deterministic, no atomics).
// Optional bf16 logits: staged as [32 x 64] boxes (128-byte swizzle) and written with TMA stores.
struct SimEpi2 {
  static constexpr int kStages = 4;
  static constexpr int kWarpScratch = 2 * 4096 + 1024;    // two logits boxes + this warp's share of the column partials
  struct State {};
  SimCommon c;             // seg_tiles = ceil(T / 256) PAIR tiles per segment; m_tiles counts 128-row CTA tiles
  int pair_m_tiles;        // B_loc * S * seg_tiles
  int64_t b_stage_rows;    // rows of B per stage in the B tensor map (0 for the dual encoder)
  int store_logits;        // 0 none, 1 TMA boxes through tmOut (3-D map [segment][T][C], needs C % 8 == 0), 2 direct
  bf16* logits;            // mode 2
  float* row_part;         // [2][2 * n_tiles][R]
  float* col_part;         // [2][m_tiles][C]

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
    const int t = i * 256 + static_cast<int>(rank) * 128 + quarter * 32 + lane;   // frame index of this thread's row
    const bool row_ok = t < c.g.T;
    const int64_t r = static_cast<int64_t>(seg) * c.g.T + t;
    const int bg = c.g.b_off + seg / c.g.S;                   // global clip of this row
    const int pos_c0 = bg * c.g.N, pos_c1 = pos_c0 + c.g.N;   // the only columns that can be positive
    const float tf = static_cast<float>(t);
    const int n0 = tn * kG2BN + half * 128;                   // first column of this warp
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
    // column partials: [quarter][2 (all, pos)][256] floats spread over the eight 1 KB warp tails
    uint8_t* sbase = ws - ew * kWarpScratch;
    auto scol = [&](int q, int which, int col) -> float& {
      const int idx = (q * 2 + which) * 256 + col;            // 0 .. 2047
      return *reinterpret_cast<float*>(sbase + (idx >> 8) * kWarpScratch + 2 * 4096 + (idx & 255) * 4);
    };
    const bool all_rows = __all_sync(0xffffffffu, row_ok);
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
      // lane j looks up column col0 + j once; shuffled to everyone below
      const int mycol = col0 + lane;
      const bool my_ok = mycol < c.g.C && c.col_valid[mycol] != 0;
      const uint32_t okmask = __ballot_sync(0xffffffffu, my_ok);
      const bool chunk_has_pos = col0 < pos_c1 && col0 + 32 > pos_c0;   // warp-uniform
      float my_start = 0.f, my_end = 0.f;
      if (chunk_has_pos && my_ok && mycol >= pos_c0 && mycol < pos_c1) {
        my_start = c.start[mycol];
        my_end = c.end[mycol];                                           // else start >= end: never positive
      }
      if (store_logits == 2 && row_ok) {                        // generic tail path (C % 8 != 0): per-thread stores
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
      if (all_rows && okmask == 0xffffffffu) {                 // common case: no masking at all
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          e[j] = fast_exp2(fmaf(__uint_as_float(rc[j]), kExpScale, -kExpScale));
          row_all += e[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = fast_exp2(fmaf(__uint_as_float(rc[j]), kExpScale, -kExpScale));
          e[j] = (row_ok && ((okmask >> j) & 1u)) ? x : 0.f;
          row_all += e[j];
        }
      }
      float cpos = 0.f;
      if (chunk_has_pos) {
        float pe[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float s = __shfl_sync(0xffffffffu, my_start, j);
          const float en = __shfl_sync(0xffffffffu, my_end, j);
          pe[j] = (s <= tf && tf < en) ? e[j] : 0.f;
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
        for (int p = 0; p < 2; ++p)
          if (n0 + 64 * p < c.g.C) tma_store_3d(tmOut, ws + p * 4096, n0 + 64 * p, t0, seg);
        tma_store_commit();
      }
    }
    if (row_ok) {
      const int64_t part = static_cast<int64_t>(tn) * 2 + half;
      row_part[(static_cast<int64_t>(0) * 2 * c.n_tiles + part) * c.R + r] = row_all;
      row_part[(static_cast<int64_t>(1) * 2 * c.n_tiles + part) * c.R + r] = row_pos;
    }
    // combine the four quarters in a fixed order (deterministic) and publish this CTA tile's column partials
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int tm = (seg * c.seg_tiles + i) * 2 + static_cast<int>(rank);     // 128-row CTA tile index
    const int tid = ew * 32 + lane;
    for (int j = tid; j < 2 * 256; j += 256) {
      const int which = j >> 8, cj = j & 255;
      const int col = tn * kG2BN + cj;
      if (col < c.g.C) {
        const float sum = ((scol(0, which, cj) + scol(1, which, cj)) + scol(2, which, cj)) + scol(3, which, cj);
        col_part[(static_cast<int64_t>(which) * c.m_tiles + tm) * c.g.C + col] = sum;
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // scratch is reused by the next tile
  }
};

// row_sums[w][r] = sum_p row_part[w][p][r] (p < row_parts);  col_sums[w][s][c] = sum over the m-tiles of stage s
// (seg_parts consecutive m-tiles per (clip, stage) segment).
__global__ void sim_reduce_partials_kernel(const float* __restrict__ row_part, const float* __restrict__ col_part,
                                           SimCommon c, int row_parts, int seg_parts, float* __restrict__ row_sums,
                                           float* __restrict__ col_sums) {
  const int64_t nrow = 2 * c.R;
  const int64_t ncol = 2ll * c.g.S * c.g.C;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nrow + ncol; i += stride) {
    if (i < nrow) {
      const int64_t w = i / c.R, r = i % c.R;
      float s = 0.f;
      for (int p = 0; p < row_parts; ++p) s += row_part[(w * row_parts + p) * c.R + r];
      row_sums[i] = s;
    } else {
      const int64_t k = i - nrow;
      const int64_t w = k / (static_cast<int64_t>(c.g.S) * c.g.C);
      const int64_t rem = k % (static_cast<int64_t>(c.g.S) * c.g.C);
      const int s_idx = static_cast<int>(rem / c.g.C), col = static_cast<int>(rem % c.g.C);
      float s = 0.f;
      for (int b = 0; b < c.g.B_loc; ++b) {
        const int seg = b * c.g.S + s_idx;
        for (int t = 0; t < seg_parts; ++t)
          s += col_part[(w * c.m_tiles + (seg * seg_parts + t)) * c.g.C + col];
      }
      col_sums[k] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Statistics from materialised logits (HBM-bound).
// CTA = 8 warps; a CTA owns a (segment, 256-column slab) and walks the segment's T rows.  Each lane owns
// 8 consecutive columns (one 16-byte load of bf16, two of fp32), so a warp reads 512 contiguous bytes
// per row.  A warp takes kNceRows rows per iteration and issues all their loads before touching any of
// them (memory-level parallelism: 8 x 512 B in flight per warp); column sums live in registers for the
// whole walk; the kNceRows row sums are reduced together by one transposing butterfly (9 shuffles for 8
// values instead of 5 per value).
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
__global__ void __launch_bounds__(kNceWarps * 32, 2)
nce_from_logits_kernel(const void* __restrict__ logits, SimCommon c, float* __restrict__ row_part,
                       float* __restrict__ col_sums_part) {
  __shared__ float scol[kNceWarps][2][kNceCols];
  const int slab = blockIdx.x;                 // column slab
  const int seg = blockIdx.y;                  // (clip, stage)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = slab * kNceCols + lane * 8;
  const int bg = c.g.b_off + seg / c.g.S;
  const int pos_c0 = bg * c.g.N, pos_c1 = pos_c0 + c.g.N;
  const bool aligned = (c.g.C % 8) == 0 && col0 + 8 <= c.g.C;
  const bool slab_has_pos = slab * kNceCols < pos_c1 && (slab + 1) * kNceCols > pos_c0;   // CTA-uniform
  pdl_launch_dependents();
  pdl_wait();

  uint32_t okbits = 0;
  float st[8], en[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = col0 + j;
    const bool ok = col < c.g.C && c.col_valid[col] != 0;
    okbits |= ok ? (1u << j) : 0u;
    const bool pc = ok && col >= pos_c0 && col < pos_c1;
    st[j] = pc ? c.start[col] : 1.f;
    en[j] = pc ? c.end[col] : 0.f;             // start >= end: never positive
  }
  float call[8], cpos[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { call[j] = 0.f; cpos[j] = 0.f; }

  for (int t0 = warp * kNceRows; t0 < c.g.T; t0 += kNceWarps * kNceRows) {
    // ---- issue every load of this row group first (raw 16-byte words: 4 (bf16) / 8 (fp32) registers per row)
    uint4 raw[kNceRows][F32 ? 2 : 1];
#pragma unroll
    for (int i = 0; i < kNceRows; ++i) {
      const int t = t0 + i;
      const int64_t r = static_cast<int64_t>(seg) * c.g.T + (t < c.g.T ? t : c.g.T - 1);
      if (aligned) {
        if (F32) {
          const uint4* p = reinterpret_cast<const uint4*>(static_cast<const float*>(logits) + r * c.g.C + col0);
          raw[i][0] = __ldcs(p);
          raw[i][F32 ? 1 : 0] = __ldcs(p + 1);
        } else {
          raw[i][0] = __ldcs(reinterpret_cast<const uint4*>(static_cast<const bf16*>(logits) + r * c.g.C + col0));
        }
      } else {
        float xs[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = col0 + j;
          xs[j] = 0.f;
          if (col < c.g.C)
            xs[j] = F32 ? static_cast<const float*>(logits)[r * c.g.C + col]
                        : __bfloat162float(static_cast<const bf16*>(logits)[r * c.g.C + col]);
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
    // ---- consume
    float ra[kNceRows], rp[kNceRows];
#pragma unroll
    for (int i = 0; i < kNceRows; ++i) {
      float x[8];
      if (F32) {
        x[0] = __uint_as_float(raw[i][0].x); x[1] = __uint_as_float(raw[i][0].y);
        x[2] = __uint_as_float(raw[i][0].z); x[3] = __uint_as_float(raw[i][0].w);
        x[4] = __uint_as_float(raw[i][F32 ? 1 : 0].x); x[5] = __uint_as_float(raw[i][F32 ? 1 : 0].y);
        x[6] = __uint_as_float(raw[i][F32 ? 1 : 0].z); x[7] = __uint_as_float(raw[i][F32 ? 1 : 0].w);
      } else {
        const float2 a = unpack_bf16x2(raw[i][0].x), b = unpack_bf16x2(raw[i][0].y);
        const float2 cc = unpack_bf16x2(raw[i][0].z), d = unpack_bf16x2(raw[i][0].w);
        x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = cc.x; x[5] = cc.y; x[6] = d.x; x[7] = d.y;
      }
      const bool row_ok = t0 + i < c.g.T;
      const float tf = static_cast<float>(t0 + i);
      ra[i] = 0.f;
      rp[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float e = (((okbits >> j) & 1u) && row_ok) ? fast_exp2(fmaf(x[j], kExpScale, -kExpScale)) : 0.f;
        call[j] += e;
        ra[i] += e;
        if (slab_has_pos) {
          const float pe = (st[j] <= tf && tf < en[j]) ? e : 0.f;
          cpos[j] += pe;
          rp[i] += pe;
        }
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
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    scol[warp][0][lane * 8 + j] = call[j];
    scol[warp][1][lane * 8 + j] = cpos[j];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * kNceCols; j += blockDim.x) {
    const int which = j / kNceCols, cj = j % kNceCols;
    const int col = slab * kNceCols + cj;
    if (col < c.g.C) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kNceWarps; ++w) s += scol[w][which][cj];
      // one partial per segment: [2][B_loc*S][C]
      col_sums_part[(static_cast<int64_t>(which) * gridDim.y + seg) * c.g.C + col] = s;
    }
  }
}

__global__ void nce_reduce_kernel(const float* __restrict__ row_sums, int64_t R, const float* __restrict__ col_sums,
                                  int64_t SC, int do_rows, int do_cols, double* __restrict__ out) {
  // fp64 accumulation: the cross-block atomic order then only perturbs bits far below fp32 epsilon
  double acc[4] = {0., 0., 0., 0.};
  pdl_launch_dependents();
  pdl_wait();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (do_rows) {
    for (int64_t i = i0; i < R; i += stride) {
      const float all = row_sums[i], pos = row_sums[R + i];
      if (pos > 0.f) { acc[0] += static_cast<double>(logf(all) - logf(pos)); acc[1] += 1.; }
    }
  }
  if (do_cols) {
    for (int64_t i = i0; i < SC; i += stride) {
      const float all = col_sums[i], pos = col_sums[SC + i];
      if (pos > 0.f) { acc[2] += static_cast<double>(logf(all) - logf(pos)); acc[3] += 1.; }
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
      if (lane == 0 && v != 0.) atomicAdd(out + k, v);
    }
  }
}

static int fill_common(SimCommon* c, const tan_sim_geom* g, const float* start, const float* end,
                       const uint8_t* col_valid) {
  if (g == nullptr || start == nullptr || end == nullptr || col_valid == nullptr)
    return set_error(TAN_ERR_ARG, "sim/nce: null geometry or mask pointer");
  if (g->B_loc <= 0 || g->S <= 0 || g->T <= 0 || g->C <= 0 || g->N <= 0 || g->d <= 0 || g->b_off < 0 ||
      g->C % g->N != 0)
    return set_error(TAN_ERR_SHAPE, "sim/nce: bad geometry B_loc=%d S=%d T=%d C=%d N=%d d=%d", g->B_loc, g->S, g->T,
                     g->C, g->N, g->d);
  c->g = *g;
  c->start = start;
  c->end = end;
  c->col_valid = col_valid;
  c->seg_tiles = (g->T + 255) / 256;                       // pair tiles (256 frames) per segment
  c->m_tiles = g->B_loc * g->S * c->seg_tiles * 2;         // 128-row CTA tiles
  c->n_tiles = (g->C + 255) / 256;                         // 256-column tiles / slabs
  c->R = static_cast<int64_t>(g->B_loc) * g->S * g->T;
  return TAN_OK;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

static size_t row_part_bytes(const SimCommon& c) { return align256(2 * 2 * static_cast<size_t>(c.n_tiles) * c.R * 4); }

}  // namespace tanb

using namespace tanb;

extern "C" size_t tan_sim_nce_workspace_bytes(const tan_sim_geom* g) {
  if (g == nullptr || g->B_loc <= 0 || g->S <= 0 || g->T <= 0 || g->C <= 0) return 0;
  // one query covers both producers: row partials [2][2 * ceil(C/256)][R] (the streaming kernel uses half of
  // them), column partials [2][B_loc * S * 2 * ceil(T/256)][C] (the streaming kernel uses one per segment)
  const int64_t R = static_cast<int64_t>(g->B_loc) * g->S * g->T;
  const int64_t n_tiles = (g->C + 255) / 256;
  const int64_t m_tiles = static_cast<int64_t>(g->B_loc) * g->S * ((g->T + 255) / 256) * 2;
  return align256(2 * 2 * n_tiles * R * 4) + align256(2 * m_tiles * g->C * 4);
}

extern "C" int tan_sim_nce_fwd(const void* vfeat, const void* tfeat, int64_t tfeat_stage_stride,
                               const tan_sim_geom* g, const float* start, const float* end,
                               const uint8_t* col_valid, void* logits_out, float* row_sums, float* col_sums,
                               void* workspace, size_t workspace_bytes, void* stream) {
  TAN_CHECK(tan_device_check());
  if (vfeat == nullptr || tfeat == nullptr || row_sums == nullptr || col_sums == nullptr)
    return set_error(TAN_ERR_ARG, "tan_sim_nce_fwd: null pointer");
  if (g == nullptr) return set_error(TAN_ERR_ARG, "tan_sim_nce_fwd: null geometry");
  SimCommon c;
  TAN_CHECK(fill_common(&c, g, start, end, col_valid));
  if (g->d % kG2BK != 0) return set_error(TAN_ERR_SHAPE, "tan_sim_nce_fwd: d %% 64 != 0 (d=%d)", g->d);
  if (tfeat_stage_stride != 0 && tfeat_stage_stride != static_cast<int64_t>(g->C) * g->d)
    return set_error(TAN_ERR_SHAPE, "tan_sim_nce_fwd: tfeat_stage_stride must be 0 or C*d");
  if (workspace == nullptr || workspace_bytes < tan_sim_nce_workspace_bytes(g))
    return set_error(TAN_ERR_WORKSPACE, "tan_sim_nce_fwd: workspace too small (%zu < %zu)", workspace_bytes,
                     tan_sim_nce_workspace_bytes(g));
  float* row_part = static_cast<float*>(workspace);
  float* col_part = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + row_part_bytes(c));
  const int64_t b_rows = tfeat_stage_stride == 0 ? g->C : static_cast<int64_t>(g->S) * g->C;
  CUtensorMap tmA, tmB, tmOut;
  TAN_CHECK(make_tmap_2d(&tmA, vfeat, 2, c.R, g->d, g->d, kG2BM));
  TAN_CHECK(make_tmap_2d(&tmB, tfeat, 2, b_rows, g->d, g->d, kG2BN / 2));
  SimEpi2 e;
  e.c = c;
  e.pair_m_tiles = g->B_loc * g->S * c.seg_tiles;
  e.b_stage_rows = tfeat_stage_stride == 0 ? 0 : g->C;
  e.logits = static_cast<bf16*>(logits_out);
  e.store_logits = 0;
  tmOut = tmA;
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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TAN_CHECK(launch_umma_gemm2<SimEpi2>(tmA, tmB, tmOut, tmOut, e, e.pair_m_tiles * c.n_tiles, g->d / kG2BK, st));
  const int64_t total = 2 * c.R + 2ll * g->S * g->C;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  return launch_pdl(sim_reduce_partials_kernel, dim3(blocks), dim3(256), 0, st, 1, row_part, col_part, c,
                    2 * c.n_tiles, 2 * c.seg_tiles, row_sums, col_sums);
}

extern "C" int tan_nce_from_logits(const void* logits, int logits_is_f32, const tan_sim_geom* g,
                                   const float* start, const float* end, const uint8_t* col_valid,
                                   float* row_sums, float* col_sums, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  TAN_CHECK(tan_device_check());
  if (logits == nullptr || row_sums == nullptr || col_sums == nullptr)
    return set_error(TAN_ERR_ARG, "tan_nce_from_logits: null pointer");
  if (g == nullptr) return set_error(TAN_ERR_ARG, "tan_nce_from_logits: null geometry");
  SimCommon c;
  TAN_CHECK(fill_common(&c, g, start, end, col_valid));
  if (workspace == nullptr || workspace_bytes < tan_sim_nce_workspace_bytes(g))
    return set_error(TAN_ERR_WORKSPACE, "tan_nce_from_logits: workspace too small (%zu < %zu)", workspace_bytes,
                     tan_sim_nce_workspace_bytes(g));
  // reuse the partial layout of the fused kernel with one "m tile" per segment
  c.seg_tiles = 1;
  c.m_tiles = g->B_loc * g->S;
  float* row_part = static_cast<float*>(workspace);
  float* col_part = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + row_part_bytes(c));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g->B_loc * g->S > 65535) return set_error(TAN_ERR_SHAPE, "tan_nce_from_logits: B_loc*S > 65535");
  dim3 grid(c.n_tiles, g->B_loc * g->S);
  if (logits_is_f32)
    TAN_CHECK(launch_pdl(nce_from_logits_kernel<true>, grid, dim3(kNceWarps * 32), 0, st, 1, logits, c, row_part, col_part));
  else
    TAN_CHECK(launch_pdl(nce_from_logits_kernel<false>, grid, dim3(kNceWarps * 32), 0, st, 1, logits, c, row_part, col_part));
  const int64_t total = 2 * c.R + 2ll * g->S * g->C;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  return launch_pdl(sim_reduce_partials_kernel, dim3(blocks), dim3(256), 0, st, 1, row_part, col_part, c, c.n_tiles,
                    1, row_sums, col_sums);
}

extern "C" int tan_nce_reduce(const float* row_sums, int64_t R, const float* col_sums, int64_t SC, int do_rows,
                              int do_cols, double* out, void* stream) {
  TAN_CHECK(tan_device_check());
  if (out == nullptr || (do_rows && row_sums == nullptr) || (do_cols && col_sums == nullptr))
    return set_error(TAN_ERR_ARG, "tan_nce_reduce: null pointer");
  const int64_t n = (do_rows ? R : 0) > (do_cols ? SC : 0) ? R : SC;
  int blocks = static_cast<int>((n + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
  return launch_pdl(nce_reduce_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, row_sums, R,
                    col_sums, SC, do_rows, do_cols, out);
}
