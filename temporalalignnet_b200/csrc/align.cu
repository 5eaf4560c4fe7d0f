// Sliding-window alignment inference: the stitching of eval/eval_zeroshot_align.py:198-205 and
// the per-sentence decision of :222-238 as two kernels.  The reference runs a python loop over the windows of a
// video, each iteration a handful of slice-add launches; here every (sentence, frame) output sums, in window order,
// the windows that cover it, and a warp per sentence takes the soft-max arg-max over time.
#include "common.cuh"

namespace tanb {
namespace {

// `logits / 0.07` as torch evaluates it on a float tensor: a multiplication by the fp32 reciprocal of the fp32 scalar
// (ATen div_true_kernel_cuda with a scalar divisor), so that the stitched values equal the reference's bit for bit
constexpr float kInvTau = 1.0f / 0.07f;

// sim_x[n, t] = sum over windows w covering (n, t), in window order, of blk_x[w, t - t0_w, n - n0_w] / 0.07 (kInvTau)
//             / max(cover[n, t], 1e-5);  cover[n, t] = number of those windows.
// win: [W, 4] = (t0, t1, n0, n1).  blk: [W, T, N] cosines (the last stage's own-clip blocks).
// A long video comes in several batches of windows: `accumulate` continues the running (un-normalised) sums and
// counts a previous call left in the outputs, `finalize` divides.
__global__ void __launch_bounds__(256) align_stitch_kernel(const float* __restrict__ blk_j, const float* __restrict__ blk_d,
                                                           const int* __restrict__ win, int W, int T, int N,
                                                           float* __restrict__ sim_j, float* __restrict__ sim_d,
                                                           float* __restrict__ cover, int n_text, int vlen,
                                                           int accumulate, int finalize) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(n_text) * vlen) return;
  const int n = static_cast<int>(idx / vlen), t = static_cast<int>(idx - static_cast<int64_t>(n) * vlen);
  float aj = 0.f, ad = 0.f, c = 0.f;
  if (accumulate) {
    aj = sim_j[idx];
    ad = sim_d[idx];
    c = cover[idx];
  }
  for (int w = 0; w < W; ++w) {
    const int4 q = __ldg(reinterpret_cast<const int4*>(win) + w);       // warp-uniform address: one broadcast load
    if (t >= q.x && t < q.y && n >= q.z && n < q.w) {
      const int64_t off = (static_cast<int64_t>(w) * T + (t - q.x)) * N + (n - q.z);
      aj = __fadd_rn(aj, __fmul_rn(blk_j[off], kInvTau));      // two roundings, as the reference's mul and add kernels
      ad = __fadd_rn(ad, __fmul_rn(blk_d[off], kInvTau));      // (no contraction into one FMA)
      c += 1.f;
    }
  }
  const float den = finalize ? fmaxf(c, 1e-5f) : 1.f;
  sim_j[idx] = finalize ? aj / den : aj;
  sim_d[idx] = finalize ? ad / den : ad;
  cover[idx] = c;
}

// out[n] = argmax_t softmax_t(s[n, :]) with s = sim where sim != 0 else -6e4 (uncovered entries), first index on ties.
__global__ void __launch_bounds__(256) align_argmax_kernel(const float* __restrict__ sim, int n_text, int vlen,
                                                           int64_t* __restrict__ out) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= n_text) return;
  const float* row = sim + static_cast<int64_t>(n) * vlen;
  float mx = -INFINITY;
  for (int t = lane; t < vlen; t += 32) {
    const float v = row[t];
    mx = fmaxf(mx, v == 0.f ? -6e4f : v);
  }
  mx = warp_max(mx);
  // the soft-max is monotonic, but its exponentials ROUND: entries within one ulp of the maximum tie at 1.0 and the
  // first of them wins, as in torch.softmax(...).argmax(...)
  float best = -1.f;
  int arg = vlen;
  for (int t = lane; t < vlen; t += 32) {
    const float v = row[t];
    const float e = expf((v == 0.f ? -6e4f : v) - mx);
    if (e > best) { best = e; arg = t; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  if (lane == 0) out[n] = arg < vlen ? arg : 0;
}

}  // namespace
}  // namespace tanb

using namespace tanb;

extern "C" int tan_align_stitch(const float* blk_joint, const float* blk_dual, const int* windows, int W, int T, int N,
                                float* sim_joint, float* sim_dual, float* cover, int n_text, int vlen, int accumulate,
                                int finalize, void* stream) {
  TAN_CHECK(tan_device_check());
  if (blk_joint == nullptr || blk_dual == nullptr || windows == nullptr || sim_joint == nullptr || sim_dual == nullptr ||
      cover == nullptr)
    return set_error(TAN_ERR_ARG, "tan_align_stitch: null pointer");
  if (W < 0 || T <= 0 || N <= 0 || n_text <= 0 || vlen <= 0)
    return set_error(TAN_ERR_SHAPE, "tan_align_stitch: bad dims W=%d T=%d N=%d n_text=%d vlen=%d", W, T, N, n_text, vlen);
  if (reinterpret_cast<uintptr_t>(windows) & 15)
    return set_error(TAN_ERR_SHAPE, "tan_align_stitch: the window table must be 16-byte aligned");
  const int64_t total = static_cast<int64_t>(n_text) * vlen;
  const int64_t blocks = (total + 255) / 256;
  if (blocks > 0x7fffffffll) return set_error(TAN_ERR_SHAPE, "tan_align_stitch: output too large");
  align_stitch_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      blk_joint, blk_dual, windows, W, T, N, sim_joint, sim_dual, cover, n_text, vlen, accumulate, finalize);
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}

extern "C" int tan_align_argmax(const float* sim, int n_text, int vlen, int64_t* out, void* stream) {
  TAN_CHECK(tan_device_check());
  if (sim == nullptr || out == nullptr) return set_error(TAN_ERR_ARG, "tan_align_argmax: null pointer");
  if (n_text <= 0 || vlen <= 0) return set_error(TAN_ERR_SHAPE, "tan_align_argmax: bad dims n_text=%d vlen=%d", n_text, vlen);
  align_argmax_kernel<<<(n_text + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(sim, n_text, vlen, out);
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}
