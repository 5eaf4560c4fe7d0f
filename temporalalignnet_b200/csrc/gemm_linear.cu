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
    auto go = [&](auto e) {
      e.M = M; e.N = N; e.f_tiles = f_tiles; e.n_tiles = f_tiles * t_tiles;
      e.bias = bias; e.act = act; e.residual = residual; e.ldr = ldr; e.out_f32 = out_f32; e.ldo = ldo_f32;
      e.extra_bf16 = static_cast<bf16*>(out_bf16); e.ld_extra = ldo_bf16;
      return launch_umma_gemm2<decltype(e)>(tmA, tmB, tmA, tmA, e, e.n_tiles, K / kG2BK, st);
    };
    if (act == TAN_ACT_NONE) return go(LinearEpi2<kModeF32, TAN_ACT_NONE>{});
    return go(LinearEpi2<kModeF32>{});
  }
  CUtensorMap tmOut;
  TAN_CHECK(make_tmap_2d(&tmOut, out_bf16, 2, M, N, ldo_bf16, 32));
  auto go = [&](auto e) {
    e.M = M; e.N = N; e.f_tiles = f_tiles; e.n_tiles = f_tiles * t_tiles;
    e.bias = bias; e.act = act; e.residual = nullptr; e.ldr = 0; e.out_f32 = nullptr; e.ldo = 0;
    e.extra_bf16 = nullptr; e.ld_extra = 0;
    return launch_umma_gemm2<decltype(e)>(tmA, tmB, tmOut, tmOut, e, e.n_tiles, K / kG2BK, st);
  };
  if (act == TAN_ACT_QUICKGELU) return go(LinearEpi2<kModeBf16, TAN_ACT_QUICKGELU>{});
  if (act == TAN_ACT_RELU) return go(LinearEpi2<kModeBf16, TAN_ACT_RELU>{});
  return go(LinearEpi2<kModeBf16, TAN_ACT_NONE>{});
}

namespace {

int check_linear_common(const char* fn, const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K) {
  if (A == nullptr || W == nullptr) return set_error(TAN_ERR_ARG, "%s: null A / W", fn);
  if (M <= 0 || N <= 0 || K <= 0 || K % kG2BK != 0 || N % 128 != 0)
    return set_error(TAN_ERR_SHAPE, "%s: need M>0, N%%128==0, K%%64==0 (M=%d N=%d K=%d)", fn, M, N, K);
  if (lda % 8 != 0 || ldw % 8 != 0 || lda < K || ldw < K)
    return set_error(TAN_ERR_SHAPE, "%s: lda/ldw must be >= K and multiples of 8", fn);
  return TAN_OK;
}

bool bad_bf16_out(const void* p, int64_t ld, int N) {
  return p == nullptr || ld % 8 != 0 || ld < N || (reinterpret_cast<uintptr_t>(p) & 15);
}

}  // namespace

// Training forward of c_fc (model/tfm_model.py:23-25,:37): out_act = act(A W^T + bias) AND out_pre = A W^T + bias in
// one pass -- the backward of QuickGELU needs the pre-activation, the following GEMM the activation; writing both
// from the epilogue replaces a separate elementwise pass over the [M, 4d] activations.
extern "C" int tan_linear_dual_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                    void* out_act, int64_t ldo_act, void* out_pre, int64_t ldo_pre, int M, int N, int K,
                                    int act, void* stream) {
  TAN_CHECK(tan_device_check());
  TAN_CHECK(check_linear_common("tan_linear_dual_bf16", A, lda, W, ldw, M, N, K));
  if (act != TAN_ACT_NONE && act != TAN_ACT_QUICKGELU && act != TAN_ACT_RELU)
    return set_error(TAN_ERR_ARG, "tan_linear_dual_bf16: bad act");
  if (bad_bf16_out(out_act, ldo_act, N) || bad_bf16_out(out_pre, ldo_pre, N))
    return set_error(TAN_ERR_SHAPE, "tan_linear_dual_bf16: outputs need 16-byte aligned bases and pitches >= N, %% 8 == 0");
  CUtensorMap tmA, tmB, tmOut, tmAux;
  TAN_CHECK(make_tmap_2d(&tmA, A, 2, M, K, lda, kG2BM));
  TAN_CHECK(make_tmap_2d(&tmB, W, 2, N, K, ldw, kG2BN / 2));
  TAN_CHECK(make_tmap_2d(&tmOut, out_act, 2, M, N, ldo_act, 32));
  TAN_CHECK(make_tmap_2d(&tmAux, out_pre, 2, M, N, ldo_pre, 32));
  auto go = [&](auto e) {
    e.M = M; e.N = N; e.f_tiles = (N + kG2BN - 1) / kG2BN; e.n_tiles = e.f_tiles * ((M + 2 * kG2BM - 1) / (2 * kG2BM));
    e.bias = bias; e.act = act; e.residual = nullptr; e.ldr = 0; e.out_f32 = nullptr; e.ldo = 0;
    e.extra_bf16 = nullptr; e.ld_extra = 0;
    return launch_umma_gemm2<decltype(e)>(tmA, tmB, tmOut, tmAux, e, e.n_tiles, K / kG2BK, static_cast<cudaStream_t>(stream));
  };
  if (act == TAN_ACT_QUICKGELU) return go(LinearEpi2<kModeBf16Dual, TAN_ACT_QUICKGELU>{});
  return go(LinearEpi2<kModeBf16Dual>{});
}

// Backward through c_proj AND QuickGELU in one GEMM: out = (A W^T) o gelu'(u), i.e. du = (dx W_proj) o gelu'(u)
// (autograd of model/tfm_model.py:11-13,:37).  u [M, N] bf16 are the pre-activations tan_linear_dual_bf16 stored.
extern "C" int tan_linear_gelu_bwd_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* u,
                                        int64_t ldu, void* out, int64_t ldo, int M, int N, int K, void* stream) {
  TAN_CHECK(tan_device_check());
  TAN_CHECK(check_linear_common("tan_linear_gelu_bwd_bf16", A, lda, W, ldw, M, N, K));
  if (bad_bf16_out(out, ldo, N) || bad_bf16_out(u, ldu, N))
    return set_error(TAN_ERR_SHAPE, "tan_linear_gelu_bwd_bf16: u / out need 16-byte aligned bases and pitches >= N, %% 8 == 0");
  LinearEpi2<kModeBf16, TAN_ACT_QUICKGELU_GRAD> e;
  e.M = M; e.N = N; e.f_tiles = (N + kG2BN - 1) / kG2BN; e.n_tiles = e.f_tiles * ((M + 2 * kG2BM - 1) / (2 * kG2BM));
  e.bias = nullptr; e.act = TAN_ACT_QUICKGELU_GRAD; e.residual = nullptr; e.ldr = 0; e.out_f32 = nullptr; e.ldo = 0;
  e.extra_bf16 = const_cast<bf16*>(static_cast<const bf16*>(u)); e.ld_extra = ldu;
  CUtensorMap tmA, tmB, tmOut;
  TAN_CHECK(make_tmap_2d(&tmA, A, 2, M, K, lda, kG2BM));
  TAN_CHECK(make_tmap_2d(&tmB, W, 2, N, K, ldw, kG2BN / 2));
  TAN_CHECK(make_tmap_2d(&tmOut, out, 2, M, N, ldo, 32));
  return launch_umma_gemm2<decltype(e)>(tmA, tmB, tmOut, tmOut, e, e.n_tiles, K / kG2BK,
                                        static_cast<cudaStream_t>(stream));
}
