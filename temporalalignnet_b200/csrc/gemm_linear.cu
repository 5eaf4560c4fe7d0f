// tan_linear_bf16: out = act(A @ W^T + bias) [+ residual] on tcgen05 tensor cores.
// Main loop in umma_gemm.cuh; this file holds the register epilogue and the host launcher.
#include "umma_gemm.cuh"

namespace tanb {

template <int BN>
struct LinearEpi {
  static constexpr int kExtraSmem = 0;
  int M, N;
  int m_tiles, n_tiles;
  const float* bias;
  const float* residual;
  int64_t ldr;
  float* out_f32;
  int64_t ldo_f32;
  bf16* out_bf16;
  int64_t ldo_bf16;
  int act;

  __device__ __forceinline__ int num_tiles() const { return m_tiles * n_tiles; }
  // n fastest: consecutive CTAs share the A row block (L2 hit), W stays L2-resident.
  __device__ __forceinline__ TileCoord coord(int tile) const {
    TileCoord tc;
    tc.a_row = (tile / n_tiles) * kGemmBM;
    tc.b_row = (tile % n_tiles) * BN;
    return tc;
  }

  __device__ __forceinline__ void run(int tile, uint32_t tmem_acc, int quarter, int lane, uint8_t*) const {
    const int m0 = (tile / n_tiles) * kGemmBM;
    const int n0 = (tile % n_tiles) * BN;
    const int row = m0 + quarter * 32 + lane;
    const bool row_ok = row < M;
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int col0 = n0 + c * 32;
      if (col0 >= N) break;                      // warp-uniform
      uint32_t r[32];
      tmem_ld_32x32(taddr + c * 32, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (bias != nullptr) {
        const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(b4 + j);
          v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
      }
      if (act == TAN_ACT_QUICKGELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
      }
      if (row_ok) {
        if (residual != nullptr) {
          const float4* r4 = reinterpret_cast<const float4*>(residual + static_cast<int64_t>(row) * ldr + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 x = r4[j];
            v[4 * j + 0] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
          }
        }
        if (out_f32 != nullptr) {
          float4* o4 = reinterpret_cast<float4*>(out_f32 + static_cast<int64_t>(row) * ldo_f32 + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (out_bf16 != nullptr) {
          uint4* o4 = reinterpret_cast<uint4*>(out_bf16 + static_cast<int64_t>(row) * ldo_bf16 + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
            u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
            u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
            o4[j] = u;
          }
        }
      }
    }
  }
};

template <int BN>
static int launch_linear(const CUtensorMap& tmA, const CUtensorMap& tmB, LinearEpi<BN> epi, int K,
                         cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = umma_gemm_kernel<BN, LinearEpi<BN>>;
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    TAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = epi.m_tiles * epi.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, epi, K / kGemmBK);
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}

}  // namespace tanb

using namespace tanb;

extern "C" int tan_linear_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                               const float* residual, int64_t ldr, float* out_f32, int64_t ldo_f32,
                               void* out_bf16, int64_t ldo_bf16, int M, int N, int K, int act, void* stream) {
  TAN_CHECK(tan_device_check());
  if (A == nullptr || W == nullptr || (out_f32 == nullptr && out_bf16 == nullptr))
    return set_error(TAN_ERR_ARG, "tan_linear_bf16: null A/W or no output");
  if (M <= 0 || N <= 0 || K <= 0 || K % kGemmBK != 0 || N % 32 != 0)
    return set_error(TAN_ERR_SHAPE, "tan_linear_bf16: need M>0, K%%64==0, N%%32==0 (M=%d N=%d K=%d)", M, N, K);
  if (lda % 8 != 0 || ldw % 8 != 0 || lda < K || ldw < K)
    return set_error(TAN_ERR_SHAPE, "tan_linear_bf16: lda/ldw must be >= K and multiples of 8");
  if ((out_f32 && ldo_f32 % 4 != 0) || (out_bf16 && ldo_bf16 % 8 != 0) || (residual && ldr % 4 != 0))
    return set_error(TAN_ERR_SHAPE, "tan_linear_bf16: output/residual pitches must keep 16-byte row alignment");
  if (act != TAN_ACT_NONE && act != TAN_ACT_QUICKGELU) return set_error(TAN_ERR_ARG, "tan_linear_bf16: bad act");

  // Tile-width choice: the widest tile that still gives every SM work (grid = 148 persistent CTAs).
  const int m_tiles = (M + kGemmBM - 1) / kGemmBM;
  int bn = 256;
  if (N % 256 != 0 || m_tiles * (N / 256) < num_sms()) bn = 128;
  if (bn == 128 && (N % 128 != 0 || m_tiles * ((N + 127) / 128) < num_sms() / 2)) bn = 64;

  CUtensorMap tmA, tmB;
  TAN_CHECK(make_tmap_2d_bf16(&tmA, A, M, K, lda, kGemmBM, kGemmBK));
  TAN_CHECK(make_tmap_2d_bf16(&tmB, W, N, K, ldw, bn, kGemmBK));
  cudaStream_t st = static_cast<cudaStream_t>(stream);

#define TAN_LAUNCH_LINEAR(BN_)                                                         \
  {                                                                                    \
    LinearEpi<BN_> e;                                                                  \
    e.M = M; e.N = N; e.m_tiles = m_tiles; e.n_tiles = (N + BN_ - 1) / BN_;            \
    e.bias = bias; e.residual = residual; e.ldr = ldr;                                 \
    e.out_f32 = out_f32; e.ldo_f32 = ldo_f32;                                          \
    e.out_bf16 = static_cast<bf16*>(out_bf16); e.ldo_bf16 = ldo_bf16; e.act = act;     \
    return launch_linear<BN_>(tmA, tmB, e, K, st);                                     \
  }
  if (bn == 256) TAN_LAUNCH_LINEAR(256)
  if (bn == 128) TAN_LAUNCH_LINEAR(128)
  TAN_LAUNCH_LINEAR(64)
#undef TAN_LAUNCH_LINEAR
}
