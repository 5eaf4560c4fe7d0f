// tan_linear_bf16: out = act(A @ W^T + bias) [+ residual] on tcgen05 tensor cores.
// Main loop in umma_gemm.cuh; this file holds the register epilogue and the host launcher.
#include <cstdlib>

#include "umma_gemm.cuh"

namespace tanb {

// Operand roles are SWAPPED relative to the textbook layout: the MMA's M dimension (TMEM lanes) runs
// over 128 output FEATURES (rows of W) and its N dimension (TMEM columns) over BN TOKENS (rows of A):
//     D^T[128 features x BN tokens] = W_tile[128 x K] * A_tile[BN x K]^T        (both operands K-major)
// After tcgen05.ld each epilogue thread owns ONE feature for 32 consecutive tokens (bias = 1 register).
// A 4x4 transpose inside every quad of lanes (4 shuffles per 4 values, registers only) then gives each
// lane 4 CONSECUTIVE features of one token, so the fp32 residual / output move as 16-byte vectors and a
// warp instruction covers 4 token rows x 128 contiguous bytes.
// Why this shape (measured with scripts/gemm_trace.py on B200, per 128x256 tile, MMA time ~2.5 us):
//   row-per-lane 16-byte stores (32 scattered lines / instruction) .. 4.6 us bf16, 11.2 us fp32+residual
//   transposing through shared memory (STS/LDS.128) .................. 6-13 us: smem bandwidth belongs to the MMA
//   feature-per-lane scalar coalesced stores .......................... 3.0 us bf16, 9.4 us fp32+residual
// i.e. the LSU charges per INSTRUCTION (~8 cycles even when coalesced), so: few, wide, coalesced, no smem.
__device__ __forceinline__ void quad_transpose4(float (&x)[4], int lane) {
  // in:  lane q (= lane & 3) holds feature q for tokens 0..3;  out: lane q holds token q, features 0..3
  const bool hi = (lane & 2) != 0, lo = (lane & 1) != 0;
  // step A (partner lane ^ 2): afterwards tokens {2hi, 2hi+1} x features {lo, lo+2}
  const float sA0 = hi ? x[0] : x[2], sA1 = hi ? x[1] : x[3];
  const float kA0 = hi ? x[2] : x[0], kA1 = hi ? x[3] : x[1];
  const float rA0 = __shfl_xor_sync(0xffffffffu, sA0, 2), rA1 = __shfl_xor_sync(0xffffffffu, sA1, 2);
  const float y00 = hi ? rA0 : kA0, y01 = hi ? rA1 : kA1;      // feature lo,     tokens 2hi, 2hi+1
  const float y10 = hi ? kA0 : rA0, y11 = hi ? kA1 : rA1;      // feature lo + 2, tokens 2hi, 2hi+1
  // step B (partner lane ^ 1): afterwards token 2hi + lo = q, features 0..3
  const float sB0 = lo ? y00 : y01, sB1 = lo ? y10 : y11;
  const float kB0 = lo ? y01 : y00, kB1 = lo ? y11 : y10;
  const float rB0 = __shfl_xor_sync(0xffffffffu, sB0, 1), rB1 = __shfl_xor_sync(0xffffffffu, sB1, 1);
  x[0] = lo ? rB0 : kB0;
  x[1] = lo ? kB0 : rB0;
  x[2] = lo ? rB1 : kB1;
  x[3] = lo ? kB1 : rB1;
}

template <int BN>
struct LinearEpi {
  static constexpr int kExtraSmem = 0;
  struct State {
    float bias;
    float4 res[8];
  };
  int M, N;                // tokens, features
  int cs;                  // cluster size (feature blocks per cluster tile)
  int f_blocks, f_groups;  // N / 128, f_blocks / cs
  int t_tiles;             // ceil(M / BN)
  const float* bias;
  const float* residual;
  int64_t ldr;
  float* out_f32;
  int64_t ldo_f32;
  bf16* out_bf16;
  int64_t ldo_bf16;
  int act;

  __device__ __forceinline__ int num_ctiles() const { return t_tiles * f_groups; }
  // cluster tile = (token tile, feature group); the cs CTAs of a cluster share the token tile (B operand)
  __device__ __forceinline__ int tile_id(int ct, int rank) const {
    return (ct / f_groups) * f_blocks + (ct % f_groups) * cs + rank;
  }
  __device__ __forceinline__ TileCoord coord(int tile) const {
    TileCoord tc;
    tc.a_row = (tile % f_blocks) * kGemmBM;     // rows of W
    tc.b_row = (tile / f_blocks) * BN;          // rows of A (tokens)
    return tc;
  }
  __device__ __forceinline__ void init(uint8_t*) const {}

  // After the quad transpose lane l owns, for k = 0..7: token  tok0 + 4k + (l & 3),
  // features f0 + 4*(l >> 2) .. +3   (f0 = first feature of this warp's 32-feature slice).
  template <bool GUARD>
  __device__ __forceinline__ void load_res(float4 (&r)[8], int tok, int col) const {
    const float* p = residual + static_cast<int64_t>(tok) * ldr + col;
    const uint32_t ld4 = 4u * static_cast<uint32_t>(ldr);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      r[k] = (!GUARD || tok + 4 * k < M) ? *reinterpret_cast<const float4*>(p + k * ld4)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void load_res_any(float4 (&r)[8], int tok0, int lane, int f0) const {
    const int tok = tok0 + (lane & 3), col = f0 + 4 * (lane >> 2);
    if (tok0 + 32 <= M) load_res<false>(r, tok, col);
    else load_res<true>(r, tok, col);
  }

  __device__ __forceinline__ void pre(int tile, int quarter, int lane, uint8_t*, State& st) const {
    const int f0 = (tile % f_blocks) * kGemmBM + quarter * 32;
    st.bias = (bias != nullptr) ? __ldg(bias + f0 + lane) : 0.f;
    if (residual != nullptr) load_res_any(st.res, (tile / f_blocks) * BN, lane, f0);
  }

  template <bool GUARD>
  __device__ __forceinline__ void store_chunk(const float (&v)[32], const float4 (&cur)[8], int tok0, int lane,
                                              int f0) const {
    const int tok = tok0 + (lane & 3), col = f0 + 4 * (lane >> 2);
    float* pf = out_f32 != nullptr ? out_f32 + static_cast<int64_t>(tok) * ldo_f32 + col : nullptr;
    bf16* pb = out_bf16 != nullptr ? out_bf16 + static_cast<int64_t>(tok) * ldo_bf16 + col : nullptr;
    const uint32_t ldf4 = 4u * static_cast<uint32_t>(ldo_f32), ldb4 = 4u * static_cast<uint32_t>(ldo_bf16);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float x[4] = {v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]};
      quad_transpose4(x, lane);
      if (residual != nullptr) { x[0] += cur[k].x; x[1] += cur[k].y; x[2] += cur[k].z; x[3] += cur[k].w; }
      if (!GUARD || tok + 4 * k < M) {
        if (pf != nullptr) *reinterpret_cast<float4*>(pf + k * ldf4) = make_float4(x[0], x[1], x[2], x[3]);
        if (pb != nullptr)
          *reinterpret_cast<uint2*>(pb + k * ldb4) = make_uint2(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]));
      }
    }
  }

  __device__ __forceinline__ void run(int tile, uint32_t tmem_acc, int quarter, int lane, uint8_t*, State& st) const {
    const int f0 = (tile % f_blocks) * kGemmBM + quarter * 32;            // this warp's first output feature
    const int t0 = (tile / f_blocks) * BN;
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16);
    int n_chunks = (M - t0 + 31) / 32;
    if (n_chunks > BN / 32) n_chunks = BN / 32;
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
      const int tok0 = t0 + c * 32;
      uint32_t r[32];
      tmem_ld_32x32(taddr + c * 32, r);
      float4 cur[8];
      if (residual != nullptr) {
#pragma unroll
        for (int k = 0; k < 8; ++k) cur[k] = st.res[k];
        if (c + 1 < n_chunks) load_res_any(st.res, tok0 + 32, lane, f0);   // in flight while chunk c is processed
      }
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + st.bias;
      if (act == TAN_ACT_QUICKGELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
      }
      if (tok0 + 32 <= M) store_chunk<false>(v, cur, tok0, lane, f0);
      else store_chunk<true>(v, cur, tok0, lane, f0);
    }
  }
};

}  // namespace tanb

using namespace tanb;

extern "C" int tan_linear_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                               const float* residual, int64_t ldr, float* out_f32, int64_t ldo_f32,
                               void* out_bf16, int64_t ldo_bf16, int M, int N, int K, int act, void* stream) {
  TAN_CHECK(tan_device_check());
  if (A == nullptr || W == nullptr || (out_f32 == nullptr && out_bf16 == nullptr))
    return set_error(TAN_ERR_ARG, "tan_linear_bf16: null A/W or no output");
  if (M <= 0 || N <= 0 || K <= 0 || K % kGemmBK != 0)
    return set_error(TAN_ERR_SHAPE, "tan_linear_bf16: need M>0, N>0, K%%64==0 (M=%d N=%d K=%d)", M, N, K);
  if (lda % 8 != 0 || ldw % 8 != 0 || lda < K || ldw < K)
    return set_error(TAN_ERR_SHAPE, "tan_linear_bf16: lda/ldw must be >= K and multiples of 8");
  if (act != TAN_ACT_NONE && act != TAN_ACT_QUICKGELU) return set_error(TAN_ERR_ARG, "tan_linear_bf16: bad act");
  if (N % 128 != 0)
    return set_error(TAN_ERR_SHAPE, "tan_linear_bf16: N (output features) must be a multiple of 128 (N=%d)", N);
  // the epilogue addresses rows with 32-bit element offsets relative to a 64-bit per-thread base
  if (ldr * 64 > 0x7fffffffll || ldo_f32 * 64 > 0x7fffffffll || ldo_bf16 * 64 > 0x7fffffffll)
    return set_error(TAN_ERR_SHAPE, "tan_linear_bf16: row pitch too large");
  if ((out_f32 && (ldo_f32 % 4 != 0 || (reinterpret_cast<uintptr_t>(out_f32) & 15))) ||
      (out_bf16 && (ldo_bf16 % 4 != 0 || (reinterpret_cast<uintptr_t>(out_bf16) & 7))) ||
      (residual && (ldr % 4 != 0 || (reinterpret_cast<uintptr_t>(residual) & 15))))
    return set_error(TAN_ERR_SHAPE, "tan_linear_bf16: outputs / residual must keep 16-byte (bf16: 8-byte) alignment "
                                    "per 4-feature group");

  // Token-tile width: the widest tile that still gives every SM work (persistent grid of 148 CTAs).
  // Cluster size over feature blocks (token-tile multicast): measured neutral on B200, default 1
  // (TAN_GEMM_CS / TAN_GEMM_BN override for experiments).
  const int f_blocks = N / kGemmBM;
  int bn = 256;
  if (f_blocks * ((M + 255) / 256) < num_sms()) bn = 128;
  if (bn == 128 && f_blocks * ((M + 127) / 128) < num_sms() / 2) bn = 64;
  int cs = 1;
  if (const char* e = getenv("TAN_GEMM_CS")) {
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4) cs = v;
  }
  if (const char* e = getenv("TAN_GEMM_BN")) {
    const int v = atoi(e);
    if (v == 64 || v == 128 || v == 256) bn = v;
  }
  while (cs > 1 && f_blocks % cs != 0) cs >>= 1;

  CUtensorMap tmA, tmB;
  TAN_CHECK(make_tmap_2d_bf16(&tmA, W, N, K, ldw, kGemmBM, kGemmBK));          // MMA "A" = weights (features)
  TAN_CHECK(make_tmap_2d_bf16(&tmB, A, M, K, lda, bn / cs, kGemmBK));           // MMA "B" = activations (tokens)
  cudaStream_t st = static_cast<cudaStream_t>(stream);

#define TAN_LAUNCH_LINEAR(BN_, CS_)                                                            \
  {                                                                                            \
    LinearEpi<BN_> e;                                                                          \
    e.M = M; e.N = N; e.cs = CS_; e.f_blocks = f_blocks; e.f_groups = f_blocks / CS_;          \
    e.t_tiles = (M + BN_ - 1) / BN_;                                                           \
    e.bias = bias; e.residual = residual; e.ldr = ldr;                                         \
    e.out_f32 = out_f32; e.ldo_f32 = ldo_f32;                                                  \
    e.out_bf16 = static_cast<bf16*>(out_bf16); e.ldo_bf16 = ldo_bf16; e.act = act;             \
    return launch_umma_gemm<BN_, CS_, LinearEpi<BN_>>(tmA, tmB, e, e.t_tiles * e.f_groups, K / kGemmBK, st); \
  }
#define TAN_LAUNCH_LINEAR_CS(BN_)            \
  {                                          \
    if (cs == 4) TAN_LAUNCH_LINEAR(BN_, 4)   \
    if (cs == 2) TAN_LAUNCH_LINEAR(BN_, 2)   \
    TAN_LAUNCH_LINEAR(BN_, 1)                \
  }
  if (bn == 256) TAN_LAUNCH_LINEAR_CS(256)
  if (bn == 128) TAN_LAUNCH_LINEAR_CS(128)
  TAN_LAUNCH_LINEAR_CS(64)
#undef TAN_LAUNCH_LINEAR_CS
#undef TAN_LAUNCH_LINEAR
}
