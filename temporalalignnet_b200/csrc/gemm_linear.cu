// tan_linear_bf16: out = act(A @ W^T + bias) [+ residual] on tcgen05 tensor cores (CTA pairs).
// Main loop in umma_gemm2.cuh; this file holds the epilogue and the host launcher.
#include <cstdlib>

#include "linear_epi.cuh"

using namespace tanb;

extern "C" int tan_linear_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                               const float* residual, int64_t ldr, float* out_f32, int64_t ldo_f32,
                               void* out_bf16, int64_t ldo_bf16, int M, int N, int K, int act, void* stream) {
  TAN_CHECK(tan_device_check());
  if (A == nullptr || W == nullptr || (out_f32 == nullptr && out_bf16 == nullptr))
    return set_error(TAN_ERR_ARG, "tan_linear_bf16: null A/W or no output");
  if (M <= 0 || N <= 0 || K <= 0 || K % kG2BK != 0)
    return set_error(TAN_ERR_SHAPE, "tan_linear_bf16: need M>0, N>0, K%%64==0 (M=%d N=%d K=%d)", M, N, K);
  if (lda % 8 != 0 || ldw % 8 != 0 || lda < K || ldw < K)
    return set_error(TAN_ERR_SHAPE, "tan_linear_bf16: lda/ldw must be >= K and multiples of 8");
  if (act != TAN_ACT_NONE && act != TAN_ACT_QUICKGELU && act != TAN_ACT_RELU) return set_error(TAN_ERR_ARG, "tan_linear_bf16: bad act");
  if (N % 128 != 0)
    return set_error(TAN_ERR_SHAPE, "tan_linear_bf16: N (output features) must be a multiple of 128 (N=%d)", N);
  if (residual != nullptr && out_f32 == nullptr)
    return set_error(TAN_ERR_ARG, "tan_linear_bf16: a residual needs the fp32 output");
  if ((out_f32 && (ldo_f32 % 4 != 0 || ldo_f32 < N || (reinterpret_cast<uintptr_t>(out_f32) & 15))) ||
      (out_bf16 && (ldo_bf16 % 8 != 0 || ldo_bf16 < N || (reinterpret_cast<uintptr_t>(out_bf16) & 15))) ||
      (residual && (ldr % 4 != 0 || ldr < N || (reinterpret_cast<uintptr_t>(residual) & 15))))
    return set_error(TAN_ERR_SHAPE, "tan_linear_bf16: outputs / residual need 16-byte aligned bases and row pitches "
                                    ">= N that keep rows 16-byte aligned");

  const int f_tiles = (N + kG2BN - 1) / kG2BN;
  const int t_tiles = (M + 2 * kG2BM - 1) / (2 * kG2BM);
  CUtensorMap tmA, tmB;
  TAN_CHECK(make_tmap_2d(&tmA, A, 2, M, K, lda, kG2BM));                  // MMA "A" = activations (tokens)
  TAN_CHECK(make_tmap_2d(&tmB, W, 2, N, K, ldw, kG2BN / 2));              // MMA "B" = weights (features)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_f32 != nullptr) {
    LinearEpi2<kModeF32> e;
    e.M = M; e.N = N; e.f_tiles = f_tiles; e.n_tiles = f_tiles * t_tiles;
    e.bias = bias; e.act = act; e.residual = residual; e.ldr = ldr; e.out_f32 = out_f32; e.ldo = ldo_f32;
    e.extra_bf16 = static_cast<bf16*>(out_bf16); e.ld_extra = ldo_bf16;
    return launch_umma_gemm2<LinearEpi2<kModeF32>>(tmA, tmB, tmA, tmA, e, e.n_tiles, K / kG2BK, st);
  }
  LinearEpi2<kModeBf16> e;
  e.M = M; e.N = N; e.f_tiles = f_tiles; e.n_tiles = f_tiles * t_tiles;
  e.bias = bias; e.act = act; e.residual = nullptr; e.ldr = 0; e.out_f32 = nullptr; e.ldo = 0;
  e.extra_bf16 = nullptr; e.ld_extra = 0;
  CUtensorMap tmOut;
  TAN_CHECK(make_tmap_2d(&tmOut, out_bf16, 2, M, N, ldo_bf16, 32));
  return launch_umma_gemm2<LinearEpi2<kModeBf16>>(tmA, tmB, tmOut, tmOut, e, e.n_tiles, K / kG2BK, st);
}
