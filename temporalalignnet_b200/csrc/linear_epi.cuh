// Epilogue of the linear-layer pair GEMM (shared by tan_linear_bf16 and tan_gemm_tn_bf16).
#pragma once

#include "umma_gemm2.cuh"

namespace tanb {

// Epilogue data path (per epilogue warp: 32 token rows x 128 features of the CTA's 128 x 256 accumulator).
// tcgen05.ld 32x32b.x32 hands each thread ONE token row x 32 consecutive features; global memory wants whole
// 128-byte lines, so every chunk is transposed through a [32 rows x 128 B] staging box in shared memory with
// the 128-byte XOR swizzle (chunk ^= row & 7: conflict-free for the row-per-lane writes AND the reads below).
//   bf16 output:  two x32 chunks fill one [32 x 64] bf16 box -> fence.proxy.async -> TMA store (clips tails).
//   fp32 output (+ fp32 residual, may alias the output): the box is read back with lane = (row % 4, 16-byte
//       column), i.e. a warp instruction covers 4 rows x 128 contiguous bytes; the residual is LDG'ed and the
//       result STG'ed with that mapping (fully coalesced, the residual of the next chunk is prefetched into the
//       same registers).  Routing the fp32 residual through TMA + smem as well was measured smem-bandwidth
//       bound (4 passes of 128 KB per tile = 2.4 us against 2.5 us of MMA per K=512 tile); this way it is 2.
// The accumulator itself leaves TMEM at ~64 B/clk/SM (1.05 us per tile), which is the floor of any epilogue.
constexpr int kModeBf16 = 0;   // out_bf16 only
constexpr int kModeF32 = 1;    // out_f32 (+ residual) (+ bf16 copy)
constexpr int kModeBf16Dual = 2;   // out_bf16 = act(x) AND a second bf16 output with the pre-activation x (training tape)

// d/dx [x sigmoid(1.702 x)] = s + 1.702 x s (1 - s),  s = sigmoid(1.702 x) = 0.5 (1 + tanh(0.851 x))
__device__ __forceinline__ float quick_gelu_grad(float x) {
  const float sg = 0.5f * (1.0f + fast_tanh(0.851f * x));
  return sg * (1.0f + 1.702f * x * (1.0f - sg));
}

__device__ __forceinline__ uint32_t swz128(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// ACT >= 0: the activation is a compile-time constant of the instance (the epilogue of a K = 512 layer is as long
// as its main loop: with every activation variant compiled into one body it was ~2700 instructions per tile and the
// epilogue warps stalled on instruction fetch -- `no_inst` 17 % of all samples, ncu r02aa); ACT = -1: the runtime
// field `act` decides (rarely used combinations).
template <int MODE, int ACT = -1>
struct LinearEpi2 {
  __device__ __forceinline__ bool is_act(int x) const { return ACT >= 0 ? ACT == x : act == x; }
  // Ring depth over staging space: the MMA warp trails the TMA producer by a constant ~1.8 us (per-CTA timelines,
  // tan_debug_set_trace), so bytes in flight pace these GEMMs: the epilogue keeps ONE 4 KB staging box per warp
  // and the ring gets 6 stages (192 KB in flight per SM): +8 % over 4 stages on the K = 512 layers (same box).
  // (the dual-output mode needs a second staging box per warp and gives up one ring stage for it)
  static constexpr int kStages = MODE == kModeBf16Dual ? 5 : 6;
  static constexpr int kWarpScratch = MODE == kModeBf16Dual ? 8192 : 4096;
  struct State {
    float4 res[8];         // kModeF32: residual of the next chunk, lane = (row % 4, 16-byte column)
  };
  int M, N;                // tokens, features
  int f_tiles;             // ceil(N / 256)
  int n_tiles;             // ceil(M / 256) * f_tiles
  const float* bias;
  int act;
  const float* residual;   // kModeF32
  int64_t ldr;
  float* out_f32;
  int64_t ldo;
  bf16* extra_bf16;        // kModeF32: optional second output (bf16 copy).  kModeBf16 with act == TAN_ACT_QUICKGELU_GRAD:
  int64_t ld_extra;        // the pre-activations u [M, N] the result is multiplied with gelu'(u) of (read-only)

  __device__ __forceinline__ int num_tiles() const { return n_tiles; }
  // feature tiles fastest: concurrently resident pair tiles share token rows (A) and cover all of W
  __device__ __forceinline__ PairTile coord(int tile) const {
    PairTile pt;
    pt.a_row = (tile / f_tiles) * (2 * kG2BM);
    pt.b_row = (tile % f_tiles) * kG2BN;
    return pt;
  }

  // residual of chunk c in the read-back mapping: row = row0 + 4 i + (lane >> 3), 4 features at col + 4 (lane & 7)
  __device__ __forceinline__ void load_res(float4 (&res)[8], int row0, int col, int lane) const {
    const float* p = residual + static_cast<int64_t>(row0 + (lane >> 3)) * ldr + col + 4 * (lane & 7);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      res[i] = (row0 + 4 * i + (lane >> 3) < M && col < N) ? *reinterpret_cast<const float4*>(p + 4 * i * ldr)
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
  }

  __device__ __forceinline__ void pre(int tile, uint32_t rank, int ew, int lane, uint8_t*, float* colvec, uint64_t*,
                                      uint32_t, const CUtensorMap*, const CUtensorMap*, State& st) const {
    const int f_base = (tile % f_tiles) * kG2BN;
    // per-tile bias vector (all 256 epilogue threads); the trailing barrier of run() protects its reuse
    const int et = ew * 32 + lane;
    colvec[et] = (bias != nullptr && f_base + et < N) ? __ldg(bias + f_base + et) : 0.f;
    if (MODE == kModeF32 && residual != nullptr) {
      const int row0 = (tile / f_tiles) * (2 * kG2BM) + static_cast<int>(rank) * kG2BM + (ew & 3) * 32;
      load_res(st.res, row0, f_base + (ew >> 2) * 128, lane);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }

  __device__ __forceinline__ void run(int tile, uint32_t rank, uint32_t tmem_acc, int ew, int lane, uint8_t* ws,
                                      const float* colvec, uint64_t*, uint32_t, const CUtensorMap* tmOut,
                                      const CUtensorMap* tmAux, State& st) const {
    const int quarter = ew & 3, half = ew >> 2;
    const int row0 = (tile / f_tiles) * (2 * kG2BM) + static_cast<int>(rank) * kG2BM + quarter * 32;
    const int col0 = (tile % f_tiles) * kG2BN + half * 128;
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
    const float4* bvec = reinterpret_cast<const float4*>(colvec + half * 128);

    // Software pipeline: the tcgen05.ld of chunk c+1 is in flight while chunk c is processed.
    uint32_t r[2][32];
    tmem_ld_32x32(taddr, r[0]);
    if (MODE == kModeF32) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint8_t* buf = ws;
        tmem_ld_wait();
        if (c + 1 < 4) tmem_ld_32x32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
        const uint32_t(&rc)[32] = r[c & 1];
        if (c > 0) __syncwarp();      // chunk c-1's read-back is complete before the box is rewritten
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = bvec[c * 8 + j];
          float a0 = __uint_as_float(rc[4 * j]) + b.x, a1 = __uint_as_float(rc[4 * j + 1]) + b.y;
          float a2 = __uint_as_float(rc[4 * j + 2]) + b.z, a3 = __uint_as_float(rc[4 * j + 3]) + b.w;
          if (is_act(TAN_ACT_QUICKGELU)) { a0 = quick_gelu(a0); a1 = quick_gelu(a1); a2 = quick_gelu(a2); a3 = quick_gelu(a3); }
          else if (is_act(TAN_ACT_RELU)) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
          *reinterpret_cast<float4*>(buf + swz128(lane, j)) = make_float4(a0, a1, a2, a3);
        }
        __syncwarp();
        const int col = col0 + 32 * c;
        const int rr = lane >> 3, cc = lane & 7;
        float* po = out_f32 + static_cast<int64_t>(row0 + rr) * ldo + col + 4 * cc;
        bf16* pe = extra_bf16 != nullptr ? extra_bf16 + static_cast<int64_t>(row0 + rr) * ld_extra + col + 4 * cc : nullptr;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 x = *reinterpret_cast<const float4*>(buf + swz128(4 * i + rr, cc));
          if (residual != nullptr) {
            x.x += st.res[i].x; x.y += st.res[i].y; x.z += st.res[i].z; x.w += st.res[i].w;
          }
          if (row0 + 4 * i + rr < M && col < N) {
            *reinterpret_cast<float4*>(po + 4 * i * ldo) = x;
            if (pe != nullptr)
              *reinterpret_cast<uint2*>(pe + 4 * i * ld_extra) = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
          }
        }
        if (residual != nullptr && c + 1 < 4) load_res(st.res, row0, col + 32, lane);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        if (c + 1 < 4) tmem_ld_32x32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
        const uint32_t(&rc)[32] = r[c & 1];
        uint32_t packed[16];
        uint32_t packed_u[MODE == kModeBf16Dual ? 16 : 1];
        uint4 uu[4];                                    // QUICKGELU_GRAD: this row's 32 pre-activations of the chunk
        const bool mul_grad = MODE == kModeBf16 && is_act(TAN_ACT_QUICKGELU_GRAD);
        if (mul_grad) {
          const bool ok = row0 + lane < M && col0 + 32 * c < N;
          const uint4* pu = reinterpret_cast<const uint4*>(extra_bf16 + static_cast<int64_t>(row0 + lane) * ld_extra + col0 + 32 * c);
#pragma unroll
          for (int j = 0; j < 4; ++j) uu[j] = ok ? __ldg(pu + j) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = bvec[c * 8 + j];
          float a0 = __uint_as_float(rc[4 * j]) + b.x, a1 = __uint_as_float(rc[4 * j + 1]) + b.y;
          float a2 = __uint_as_float(rc[4 * j + 2]) + b.z, a3 = __uint_as_float(rc[4 * j + 3]) + b.w;
          if (MODE == kModeBf16Dual) {
            packed_u[2 * j] = pack_bf16x2(a0, a1);
            packed_u[2 * j + 1] = pack_bf16x2(a2, a3);
          }
          if (is_act(TAN_ACT_QUICKGELU)) { a0 = quick_gelu(a0); a1 = quick_gelu(a1); a2 = quick_gelu(a2); a3 = quick_gelu(a3); }
          else if (is_act(TAN_ACT_RELU)) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
          else if (mul_grad) {
            const uint4 q = uu[j >> 1];
            const float2 u01 = unpack_bf16x2((j & 1) ? q.z : q.x), u23 = unpack_bf16x2((j & 1) ? q.w : q.y);
            a0 *= quick_gelu_grad(u01.x); a1 *= quick_gelu_grad(u01.y);
            a2 *= quick_gelu_grad(u23.x); a3 *= quick_gelu_grad(u23.y);
          }
          packed[2 * j] = pack_bf16x2(a0, a1);
          packed[2 * j + 1] = pack_bf16x2(a2, a3);
        }
        if ((c & 1) == 0) {                             // the box's previous TMA store has read it out
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
        // 32 features = 64 bytes = chunks [4 * (c & 1), +4) of this row of the [32 x 64] bf16 box
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<uint4*>(ws + swz128(lane, 4 * (c & 1) + j)) =
              make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          if (MODE == kModeBf16Dual)
            *reinterpret_cast<uint4*>(ws + 4096 + swz128(lane, 4 * (c & 1) + j)) =
                make_uint4(packed_u[4 * j], packed_u[4 * j + 1], packed_u[4 * j + 2], packed_u[4 * j + 3]);
        }
        if (c & 1) {                                    // box complete: 64 features of 32 tokens
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < M && col0 + 32 * (c - 1) < N) {
            tma_store_2d(tmOut, ws, col0 + 32 * (c - 1), row0);
            if (MODE == kModeBf16Dual) tma_store_2d(tmAux, ws + 4096, col0 + 32 * (c - 1), row0);
            tma_store_commit();
          }
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");     // everyone is done with colvec before the next tile rewrites it
  }
};

}  // namespace tanb
