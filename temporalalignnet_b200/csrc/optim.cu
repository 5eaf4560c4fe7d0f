// Multi-tensor optimizer step of the training loop (train/main.py:112-122): per-parameter gradient clipping
// (utils/train_utils.py:3-13 -- one `.item()` host synchronisation per parameter in the reference), AdamW
// (torch.optim.AdamW, train/main.py:397) and the EMA update of the target network (model/tan_model.py:340-344) as
// TWO launches over ALL parameters, without any host synchronisation:
//   optim_sqnorm_kernel   partial sums of squares of every gradient, one CTA per 16 K-element chunk
//   optim_adamw_kernel    per chunk: the tensor's norm from its partials (fixed order), the clip coefficient,
//                         the AdamW update of param / exp_avg / exp_avg_sq and, when a target is attached,
//                         target = m * target + (1 - m) * param_new
// The arithmetic mirrors torch's multi-tensor AdamW operation by operation (one IEEE rounding per torch op, no
// cross-op contraction), so that an unclipped step is bit-identical to torch.optim.AdamW(foreach=True) in fp32.
// HBM-bound: 16 B read + 12 B written per element (+ 8 B with the EMA).
#include <cmath>

#include "common.cuh"

namespace tanb {

namespace {

constexpr int kOptThreads = 256;

__device__ __forceinline__ float block_sum_fixed(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kOptThreads / 32; ++i) t += red[i];
    red[0] = t;
  }
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(kOptThreads)
optim_sqnorm_kernel(const tan_optim_tensor* __restrict__ tab, const int* __restrict__ chunk_tensor,
                    const int64_t* __restrict__ chunk_start, int chunk_elems, float* __restrict__ partial) {
  __shared__ float red[kOptThreads / 32];
  pdl_launch_dependents();
  pdl_wait();
  const int j = blockIdx.x;
  const tan_optim_tensor t = tab[chunk_tensor[j]];
  const int64_t s = chunk_start[j];
  const int64_t e = min(s + static_cast<int64_t>(chunk_elems), t.numel);
  float acc = 0.f;
  if (t.grad != nullptr) {
    for (int64_t i = s + threadIdx.x; i < e; i += kOptThreads) {
      const float g = t.grad[i];
      acc = fmaf(g, g, acc);
    }
  }
  const float tot = block_sum_fixed(acc, red);
  if (threadIdx.x == 0) partial[j] = tot;
}

struct OptHyper {
  float clip_grad;          // <= 0: no clipping
  float beta1_w;            // 1 - beta1 (lerp weight)
  float beta2, beta2_w;     // beta2, 1 - beta2
  float bc2_sqrt;           // sqrt(1 - beta2^step)
  float eps;
  float ema_m, ema_w;       // m, 1 - m
  const float* inv_scale;   // optional device scalar multiplied into the gradients first (GradScaler.unscale_)
};

__global__ void __launch_bounds__(kOptThreads)
optim_adamw_kernel(const tan_optim_tensor* __restrict__ tab, const int* __restrict__ chunk_tensor,
                   const int64_t* __restrict__ chunk_start, const int* __restrict__ tensor_first_chunk, int chunk_elems,
                   const float* __restrict__ partial, float* __restrict__ norms_out, const OptHyper hp) {
  __shared__ float s_coef;
  pdl_launch_dependents();
  pdl_wait();
  const int j = blockIdx.x;
  const int ti = chunk_tensor[j];
  const tan_optim_tensor t = tab[ti];
  if (t.grad == nullptr) return;                      // parameter without a gradient this step: untouched, as torch
  const float inv_scale = hp.inv_scale != nullptr ? *hp.inv_scale : 1.f;
  if (threadIdx.x == 0) {
    const int c0 = tensor_first_chunk[ti], c1 = tensor_first_chunk[ti + 1];
    float sq = 0.f;
    for (int c = c0; c < c1; ++c) sq += partial[c];   // fixed order: every chunk of the tensor computes the same value
    const float norm = sqrtf(sq) * fabsf(inv_scale);
    if (j == c0 && norms_out != nullptr) norms_out[ti] = norm;
    float coef = 1.f;
    if (hp.clip_grad > 0.f) coef = fminf(1.f, hp.clip_grad / (norm + 1e-6f));
    s_coef = coef;
  }
  __syncthreads();
  const float coef = s_coef;
  const bool scale_g = coef < 1.f || inv_scale != 1.f;
  const float gmul = coef * inv_scale;
  const int64_t s = chunk_start[j];
  const int64_t e = min(s + static_cast<int64_t>(chunk_elems), t.numel);
  for (int64_t i = s + threadIdx.x; i < e; i += kOptThreads) {
    float g = t.grad[i];
    if (scale_g) g = __fmul_rn(g, gmul);
    float p = t.param[i], m = t.exp_avg[i], v = t.exp_avg_sq[i];
    p = __fmul_rn(p, t.decay);                                        // _foreach_mul_(params, 1 - lr * wd)
    m = __fmaf_rn(hp.beta1_w, __fsub_rn(g, m), m);                    // _foreach_lerp_(exp_avgs, grads, 1 - beta1)
    v = __fmul_rn(v, hp.beta2);                                       // _foreach_mul_(exp_avg_sqs, beta2)
    v = __fmaf_rn(hp.beta2_w, __fmul_rn(g, g), v);                    // _foreach_addcmul_(.., grads, grads, 1 - beta2)
    float den = __fsqrt_rn(v);                                        // _foreach_sqrt
    den = __fdiv_rn(den, hp.bc2_sqrt);                                // _foreach_div_(.., bias_correction2_sqrt)
    den = __fadd_rn(den, hp.eps);                                     // _foreach_add_(.., eps)
    p = __fmaf_rn(t.neg_step_size, __fdiv_rn(m, den), p);             // _foreach_addcdiv_(params, m, den, -lr / bc1)
    t.param[i] = p;
    t.exp_avg[i] = m;
    t.exp_avg_sq[i] = v;
    if (t.ema != nullptr)                                             // param_k * m + param_q * (1 - m)
      t.ema[i] = __fadd_rn(__fmul_rn(t.ema[i], hp.ema_m), __fmul_rn(p, hp.ema_w));
  }
}

__global__ void __launch_bounds__(kOptThreads)
optim_ema_kernel(const tan_optim_tensor* __restrict__ tab, const int* __restrict__ chunk_tensor,
                 const int64_t* __restrict__ chunk_start, int chunk_elems, float m, float w) {
  pdl_launch_dependents();
  pdl_wait();
  const int j = blockIdx.x;
  const tan_optim_tensor t = tab[chunk_tensor[j]];
  if (t.ema == nullptr) return;
  const int64_t s = chunk_start[j];
  const int64_t e = min(s + static_cast<int64_t>(chunk_elems), t.numel);
  for (int64_t i = s + threadIdx.x; i < e; i += kOptThreads)
    t.ema[i] = __fadd_rn(__fmul_rn(t.ema[i], m), __fmul_rn(t.param[i], w));
}

}  // namespace

}  // namespace tanb

using namespace tanb;

extern "C" int tan_optim_adamw_step(const tan_optim_tensor* table, int n_tensors, const int* chunk_tensor,
                                    const int64_t* chunk_start, const int* tensor_first_chunk, int n_chunks,
                                    int chunk_elems, double clip_grad, double beta1, double beta2, double eps, int step,
                                    double ema_m, const float* inv_scale, float* partial, float* norms_out,
                                    void* stream) {
  TAN_CHECK(tan_device_check());
  if (table == nullptr || chunk_tensor == nullptr || chunk_start == nullptr || tensor_first_chunk == nullptr ||
      partial == nullptr)
    return set_error(TAN_ERR_ARG, "tan_optim_adamw_step: null pointer");
  if (n_tensors <= 0 || n_chunks <= 0 || chunk_elems <= 0 || step <= 0)
    return set_error(TAN_ERR_SHAPE, "tan_optim_adamw_step: bad sizes (tensors %d, chunks %d, chunk %d, step %d)", n_tensors,
                     n_chunks, chunk_elems, step);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TAN_CHECK(launch_pdl(optim_sqnorm_kernel, dim3(n_chunks), dim3(kOptThreads), 0, st, 1, table, chunk_tensor, chunk_start,
                       chunk_elems, partial));
  OptHyper hp;
  hp.clip_grad = static_cast<float>(clip_grad);
  // scalars exactly as torch forms them: Python doubles rounded once to fp32 where the tensor op consumes them
  hp.beta1_w = static_cast<float>(1.0 - beta1);
  hp.beta2 = static_cast<float>(beta2);
  hp.beta2_w = static_cast<float>(1.0 - beta2);
  hp.bc2_sqrt = static_cast<float>(std::sqrt(1.0 - std::pow(beta2, step)));
  hp.eps = static_cast<float>(eps);
  hp.ema_m = static_cast<float>(ema_m);
  hp.ema_w = static_cast<float>(1.0 - ema_m);
  hp.inv_scale = inv_scale;
  return launch_pdl(optim_adamw_kernel, dim3(n_chunks), dim3(kOptThreads), 0, st, 1, table, chunk_tensor, chunk_start,
                    tensor_first_chunk, chunk_elems, static_cast<const float*>(partial), norms_out, hp);
}

extern "C" int tan_ema_update(const tan_optim_tensor* table, const int* chunk_tensor, const int64_t* chunk_start,
                              int n_chunks, int chunk_elems, double m, void* stream) {
  TAN_CHECK(tan_device_check());
  if (table == nullptr || chunk_tensor == nullptr || chunk_start == nullptr)
    return set_error(TAN_ERR_ARG, "tan_ema_update: null pointer");
  if (n_chunks <= 0 || chunk_elems <= 0) return set_error(TAN_ERR_SHAPE, "tan_ema_update: bad sizes");
  return launch_pdl(optim_ema_kernel, dim3(n_chunks), dim3(kOptThreads), 0, static_cast<cudaStream_t>(stream), 1, table,
                    chunk_tensor, chunk_start, chunk_elems, static_cast<float>(m), static_cast<float>(1.0 - m));
}
