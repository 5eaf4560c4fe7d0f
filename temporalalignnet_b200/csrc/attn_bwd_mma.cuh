// Register-level building blocks of the mma.sync attention backward (shared by attention_bwd.cu and its
// pipelined variant attention_bwd_pipe.cu): ldmatrix / mma wrappers, tile loads, accumulator <-> operand repacking.
#pragma once
#include "attn_bwd.cuh"

namespace tanb {
namespace abw {

constexpr int kBlk = 64;
constexpr int kPitch = 72;                 // bf16 row pitch: 144 B, ldmatrix rows fall into distinct banks
constexpr float kScale = 0.125f;
constexpr float kLog2e = 1.4426950408889634f;

struct Tiles {
  bf16 q[kBlk][kPitch];
  bf16 dO[kBlk][kPitch];
  bf16 k[kBlk][kPitch];
  bf16 v[kBlk][kPitch];
  float lse[kBlk];
  float delta[kBlk];
  float bias[kBlk];
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}

__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// rows [row0, row0 + 64) of a [*, ld] bf16 matrix, columns col0 .. col0 + 63 -> tile (zero beyond n_rows)
__device__ __forceinline__ void load_tile(bf16 (*dst)[kPitch], const bf16* src, int64_t ld, int row0, int n_rows,
                                          int64_t base_row, int col0) {
  for (int i = threadIdx.x; i < kBlk * 8; i += blockDim.x) {
    const int r = i >> 3, c8 = (i & 7) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row0 + r < n_rows) v = *reinterpret_cast<const uint4*>(src + (base_row + row0 + r) * ld + col0 + c8);
    *reinterpret_cast<uint4*>(&dst[r][c8]) = v;
  }
}

__device__ __forceinline__ void fill_bias(float* bias, const uint8_t* kpm, int b, int Lk, int k0) {
  for (int j = threadIdx.x; j < kBlk; j += blockDim.x) {
    const int key = k0 + j;
    bias[j] = (key < Lk && (kpm == nullptr || kpm[static_cast<int64_t>(b) * Lk + key] == 0)) ? 0.f : -INFINITY;
  }
}

// A fragments (4 k-steps of 16) of the warp's 16 rows m0.. of a [m][k] tile
__device__ __forceinline__ void load_a_rows(uint32_t (&a)[4][4], const bf16 (*t)[kPitch], int m0, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    ldsm_x4(a[kk], &t[m0 + (lane & 7) + 8 * ((lane >> 3) & 1)][kk * 16 + 8 * (lane >> 4)]);
}

// acc[16 x 64] += A[16 x 64] @ T^T with T a [n][k] tile (S = Q K^T and friends)
__device__ __forceinline__ void mm_nt(float (&acc)[8][4], const uint32_t (&a)[4][4], const bf16 (*t)[kPitch], int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      uint32_t b[4];
      ldsm_x4(b, &t[jp * 16 + (lane & 7) + 8 * (lane >> 4)][kk * 16 + 8 * ((lane >> 3) & 1)]);
      mma16816(acc[2 * jp], a[kk], b[0], b[1]);
      mma16816(acc[2 * jp + 1], a[kk], b[2], b[3]);
    }
  }
}

// acc[16 x 64] += A[16 x 64] @ T with T a [k][n] tile (dQ += dS K and friends)
__device__ __forceinline__ void mm_nn(float (&acc)[8][4], const uint32_t (&a)[4][4], const bf16 (*t)[kPitch], int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      uint32_t b[4];
      ldsm_x4_t(b, &t[kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)][jp * 16 + 8 * (lane >> 4)]);
      mma16816(acc[2 * jp], a[kk], b[0], b[1]);
      mma16816(acc[2 * jp + 1], a[kk], b[2], b[3]);
    }
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][4]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
}

// accumulator tiles (row g / g + 8, columns 8 j + 2 t, + 1) -> A fragments of the same 16 x 64 matrix
__device__ __forceinline__ void acc_to_a(uint32_t (&a)[4][4], const float (&acc)[8][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    a[kk][0] = pack_bf16x2(acc[2 * kk][0], acc[2 * kk][1]);
    a[kk][1] = pack_bf16x2(acc[2 * kk][2], acc[2 * kk][3]);
    a[kk][2] = pack_bf16x2(acc[2 * kk + 1][0], acc[2 * kk + 1][1]);
    a[kk][3] = pack_bf16x2(acc[2 * kk + 1][2], acc[2 * kk + 1][3]);
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// store the warp's 16 x 64 accumulator (scaled) as bf16 rows of a [*, ld] matrix
__device__ __forceinline__ void store_acc(const float (&acc)[8][4], float scale, bf16* dst, int64_t ld, int64_t base_row,
                                          int row0, int n_rows, int col0, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = row0 + g + 8 * half;
    if (r < n_rows) {
      bf16* p = dst + (base_row + r) * ld + col0 + 2 * t;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint32_t*>(p + 8 * j) = pack_bf16x2(acc[j][2 * half] * scale, acc[j][2 * half + 1] * scale);
    }
  }
}

}  // namespace abw
}  // namespace tanb
