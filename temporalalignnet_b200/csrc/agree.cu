// Own-clip similarity blocks and the agreement self-labelling of train/loss.py:88-229.
//
//  * tan_own_clip_sim    cos[b, s, t, n] = <vfeat[b,s,t], tfeat_(s)[b,n]>: the diagonal blocks of the [B,S,T,B,N]
//                        matrix (torch.diagonal(dim1=0, dim2=3), train/loss.py:92-95,:150-153,:280-283 and the
//                        eval einsum 'bstc,b(s)kc->bstk', model/tan_model.py:261-262,:280-281) on the CTA-pair
//                        tcgen05 GEMM: one 256-frame x 256-column tile per (clip, stage, frame block), of which
//                        the first N columns are kept.
//  * tan_agree_scan      per sentence: two-way softmax (over sentences, then over time), the best window of the
//                        sentence's original duration (the reference materialises a [B,N,T,T] circulant box
//                        filter, 2.1 GB at BASELINE config 5; here a warp scans the T window starts), its mean
//                        logit, and the best single-frame logit (train/loss.py:280, loss threshold).
//  * tan_agree_targets   new packed target bits from the dual / joint windows: replacement rule
//                        (temporal_agreement_type), one sentence per frame, restore emptied sentences
//                        (train/loss.py:196-226).
#include "umma_gemm2.cuh"

namespace tanb {

constexpr float kAgInvTemp = 1.0f / 0.07f;

// ---------------------------------------------------------------------------------------------
// own-clip similarity: epilogue of the pair GEMM
// ---------------------------------------------------------------------------------------------
struct DiagEpi {
  static constexpr int kStages = 4;
  static constexpr int kWarpScratch = 1024;
  struct State {};
  int B, S, T, N;          // clips, stages of vfeat, frames, sentences per clip
  int s_first, s_count;    // stage range to compute
  int seg_tiles;           // ceil(T / 256)
  int f_tiles;             // ceil(N / 256)
  int64_t b_stage_rows;    // rows of the text operand per stage (0: shared by all stages)
  float* out;              // [B, s_count, T, N]

  __device__ __forceinline__ int num_tiles() const { return B * s_count * seg_tiles * f_tiles; }
  __device__ __forceinline__ void split(int tile, int& b, int& sj, int& i, int& f) const {
    f = tile % f_tiles;
    int r = tile / f_tiles;
    i = r % seg_tiles;
    r /= seg_tiles;
    sj = r % s_count;
    b = r / s_count;
  }
  __device__ __forceinline__ PairTile coord(int tile) const {
    int b, sj, i, f;
    split(tile, b, sj, i, f);
    PairTile pt;
    pt.a_row = (b * S + s_first + sj) * T + i * 256;
    pt.b_row = static_cast<int>((s_first + sj) * b_stage_rows) + b * N + f * kG2BN;
    return pt;
  }
  __device__ __forceinline__ void pre(int, uint32_t, int, int, uint8_t*, float*, uint64_t*, uint32_t,
                                      const CUtensorMap*, const CUtensorMap*, State&) const {}
  __device__ __forceinline__ void run(int tile, uint32_t rank, uint32_t tmem_acc, int ew, int lane, uint8_t*,
                                      float*, uint64_t*, uint32_t, const CUtensorMap*, const CUtensorMap*,
                                      State&) const {
    int b, sj, i, f;
    split(tile, b, sj, i, f);
    const int quarter = ew & 3, half = ew >> 2;
    const int t = i * 256 + static_cast<int>(rank) * 128 + quarter * 32 + lane;
    const int n_base = f * kG2BN + half * 128;
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
    float* dst = out + ((static_cast<int64_t>(b) * s_count + sj) * T + t) * N;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const int n0 = n_base + ch * 32;
      if (n0 >= N) break;                              // warp-uniform
      uint32_t r[32];
      tmem_ld_32x32(taddr + ch * 32, r);
      tmem_ld_wait();
      if (t < T) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + j < N) dst[n0 + j] = __uint_as_float(r[j]);
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// agreement scan
// ---------------------------------------------------------------------------------------------
// p[b,t,n] = softmax_n(z[b,t,:]) / 0.07 with z = cos/0.07 and -6e4 on padded frames / sentences
// (train/loss.py:96-103).  One warp per (clip, frame).
__global__ void agree_rowsoftmax_kernel(const float* __restrict__ own, const uint8_t* __restrict__ vpm,
                                        const uint8_t* __restrict__ tpm, int B, int T, int N,
                                        float* __restrict__ p_scaled) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < B * T; row += warps) {
    const int b = row / T;
    const bool vpad = vpm != nullptr && vpm[row] != 0;
    const float* src = own + static_cast<int64_t>(row) * N;
    float mx = -INFINITY;
    for (int n = lane; n < N; n += 32) {
      const float z = (vpad || tpm[b * N + n] != 0) ? -6e4f : src[n] * kAgInvTemp;
      mx = fmaxf(mx, z);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int n = lane; n < N; n += 32) {
      const float z = (vpad || tpm[b * N + n] != 0) ? -6e4f : src[n] * kAgInvTemp;
      sum += __expf(z - mx);
    }
    sum = warp_sum(sum);
    for (int n = lane; n < N; n += 32) {
      const float z = (vpad || tpm[b * N + n] != 0) ? -6e4f : src[n] * kAgInvTemp;
      p_scaled[static_cast<int64_t>(row) * N + n] = __expf(z - mx) / sum * kAgInvTemp;
    }
  }
}

// One warp per sentence (b, n).  q_t = softmax_t(p_scaled[b,t,n]); windows of `dur` frames starting at
// i = 0 .. T-dur, never frames 0 and T-1 (train/loss.py:119-131); score = mean of q over the kept frames; first
// best window wins.  Outputs: win[b,n] = (lo, hi) kept frame range (lo >= hi: none), mean_logit = mean of z over it
// (:141-142), max_logit = max_t z[b,t,n] (:280; padded frames count as -6e4 only when fill_max != 0).
__global__ void agree_scan_kernel(const float* __restrict__ own, const float* __restrict__ p_scaled,
                                  const uint32_t* __restrict__ posbits, const uint8_t* __restrict__ vpm,
                                  const uint8_t* __restrict__ tpm, int B, int T, int N, int W, int fill_max,
                                  int* __restrict__ win, float* __restrict__ mean_logit,
                                  float* __restrict__ max_logit) {
  extern __shared__ float sh[];                        // [warps][T] q values
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  float* q = sh + warp * T;
  for (int sent = blockIdx.x * wpb + warp; sent < B * N; sent += gridDim.x * wpb) {
    const int b = sent / N, n = sent % N;
    const bool tpad = tpm[sent] != 0;
    // duration of the original target (>= 1; 0 for padded sentences)
    int cnt = 0;
    for (int t = lane; t < T; t += 32)
      cnt += (posbits[(static_cast<int64_t>(b) * T + t) * W + (n >> 5)] >> (n & 31)) & 1u;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    const int dur = tpad ? 0 : max(cnt, 1);
    // softmax over time
    float mx = -INFINITY, zmax = -INFINITY;
    for (int t = lane; t < T; t += 32) {
      mx = fmaxf(mx, p_scaled[(static_cast<int64_t>(b) * T + t) * N + n]);
      const bool vpad = vpm != nullptr && vpm[b * T + t] != 0;
      const float z = ((fill_max && vpad) || tpad) ? -6e4f : own[(static_cast<int64_t>(b) * T + t) * N + n] * kAgInvTemp;
      zmax = fmaxf(zmax, z);
    }
    mx = warp_max(mx);
    zmax = warp_max(zmax);
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float e = __expf(p_scaled[(static_cast<int64_t>(b) * T + t) * N + n] - mx);
      q[t] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;
    float best = 0.f;                                  // invalid windows score 0
    int best_i = 0;
    if (dur >= 1) {
      for (int i = lane; i <= T - dur; i += 32) {
        const int lo = max(i, 1), hi = min(i + dur, T - 1);
        float sc = 0.f;
        if (hi > lo) {
          const float w = 1.f / static_cast<float>(hi - lo);
          for (int t = lo; t < hi; ++t) sc += (q[t] * inv) * w;
        }
        if (sc > best) { best = sc; best_i = i; }       // strict: the first best start of this lane
      }
    }
    // warp argmax, lowest index on ties
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    int lo = 0, hi = 0;
    if (dur >= 1 && best_i <= T - dur) { lo = max(best_i, 1); hi = min(best_i + dur, T - 1); }
    float ml = 0.f;
    if (hi > lo) {
      const float w = 1.f / static_cast<float>(hi - lo);
      for (int t = lo + lane; t < hi; t += 32) {
        const bool vpad = vpm != nullptr && vpm[b * T + t] != 0;
        const float z = (vpad || tpad) ? -6e4f : own[(static_cast<int64_t>(b) * T + t) * N + n] * kAgInvTemp;
        ml += z * w;
      }
      ml = warp_sum(ml);
    } else {
      lo = hi = 0;
    }
    if (lane == 0) {
      win[2 * sent] = lo;
      win[2 * sent + 1] = hi;
      mean_logit[sent] = ml;
      max_logit[sent] = zmax;
    }
    __syncwarp();
  }
}

// One CTA per clip.  kind: 0 'i', 1 'u', 2 'keep', 3 'keep-joint' (train/loss.py:196-212).
// replace[b,n] = the per-sentence flag of the rule: 'i'/'u' use it as "keep the self-label" (confidence_mask),
// 'keep'/'keep-joint' as "replace the original timestamps" (confidence_iou).
__global__ void agree_targets_kernel(const uint32_t* __restrict__ old_bits, const int* __restrict__ win_joint,
                                     const int* __restrict__ win_dual, const uint8_t* __restrict__ replace,
                                     int T, int N, int W, int kind, uint32_t* __restrict__ new_bits) {
  extern __shared__ uint32_t sany[];                   // [W] sentences that kept at least one frame
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  for (int w = threadIdx.x; w < W; w += blockDim.x) sany[w] = 0u;
  __syncthreads();
  const int* wj = win_joint + static_cast<int64_t>(b) * N * 2;
  const int* wd = win_dual + static_cast<int64_t>(b) * N * 2;
  const uint8_t* rp = replace + static_cast<int64_t>(b) * N;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const uint32_t* ob = old_bits + (static_cast<int64_t>(b) * T + t) * W;
    uint32_t* nb = new_bits + (static_cast<int64_t>(b) * T + t) * W;
    bool found = false;                                // one sentence per frame: the first one keeps it (:217-221)
    for (int w = 0; w < W; ++w) {
      uint32_t bits = 0;
      for (int k = 0; k < 32 && w * 32 + k < N; ++k) {
        const int n = w * 32 + k;
        const bool inj = wj[2 * n] <= t && t < wj[2 * n + 1];
        const bool ind = wd[2 * n] <= t && t < wd[2 * n + 1];
        const bool old = (ob[w] >> k) & 1u;
        bool v;
        if (kind == 0) v = rp[n] && inj && ind;
        else if (kind == 1) v = rp[n] && (inj || ind);
        else if (kind == 2) v = rp[n] ? (inj || ind) : old;
        else v = rp[n] ? inj : old;
        bits |= v ? (1u << k) : 0u;
      }
      uint32_t keep = 0;
      if (!found && bits != 0) { keep = bits & (0u - bits); found = true; }     // lowest set bit
      nb[w] = keep;
      if (keep) atomicOr(&sany[w], keep);
    }
  }
  __syncthreads();
  // sentences left with nothing get their original timestamps back (:223-225)
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const uint32_t* ob = old_bits + (static_cast<int64_t>(b) * T + t) * W;
    uint32_t* nb = new_bits + (static_cast<int64_t>(b) * T + t) * W;
    for (int w = 0; w < W; ++w) nb[w] |= ob[w] & ~sany[w];
  }
}

}  // namespace tanb

using namespace tanb;

extern "C" int tan_own_clip_sim(const void* vfeat, const void* tfeat, int64_t tfeat_stage_stride, int B, int S, int T,
                                int N, int d, int s_first, int s_count, float* out, void* stream) {
  TAN_CHECK(tan_device_check());
  if (vfeat == nullptr || tfeat == nullptr || out == nullptr)
    return set_error(TAN_ERR_ARG, "tan_own_clip_sim: null pointer");
  if (B <= 0 || S <= 0 || T <= 0 || N <= 0 || d <= 0 || d % kG2BK != 0 || s_first < 0 || s_count <= 0 ||
      s_first + s_count > S)
    return set_error(TAN_ERR_SHAPE, "tan_own_clip_sim: bad dims B=%d S=%d T=%d N=%d d=%d stages [%d,+%d)", B, S, T, N, d,
                     s_first, s_count);
  const int64_t CN = static_cast<int64_t>(B) * N;
  if (tfeat_stage_stride != 0 && tfeat_stage_stride != CN * d)
    return set_error(TAN_ERR_SHAPE, "tan_own_clip_sim: tfeat_stage_stride must be 0 or B*N*d");
  DiagEpi e;
  e.B = B; e.S = S; e.T = T; e.N = N; e.s_first = s_first; e.s_count = s_count;
  e.seg_tiles = (T + 255) / 256;
  e.f_tiles = (N + kG2BN - 1) / kG2BN;
  e.b_stage_rows = tfeat_stage_stride == 0 ? 0 : CN;
  e.out = out;
  CUtensorMap tmA, tmB;
  TAN_CHECK(make_tmap_2d(&tmA, vfeat, 2, static_cast<uint64_t>(B) * S * T, d, d, kG2BM));
  TAN_CHECK(make_tmap_2d(&tmB, tfeat, 2, tfeat_stage_stride == 0 ? CN : CN * S, d, d, kG2BN / 2));
  return launch_umma_gemm2<DiagEpi>(tmA, tmB, tmA, tmA, e, B * s_count * e.seg_tiles * e.f_tiles, d / kG2BK,
                                    static_cast<cudaStream_t>(stream));
}

extern "C" size_t tan_agree_scan_workspace_bytes(int B, int T, int N) {
  if (B <= 0 || T <= 0 || N <= 0) return 0;
  return static_cast<size_t>(B) * T * N * 4;
}

extern "C" int tan_agree_scan(const float* own, const uint32_t* posbits, const uint8_t* video_padding_mask,
                              const uint8_t* text_padding_mask, int B, int T, int N, int fill_max, int* win,
                              float* mean_logit, float* max_logit, void* workspace, size_t workspace_bytes,
                              void* stream) {
  TAN_CHECK(tan_device_check());
  if (own == nullptr || posbits == nullptr || text_padding_mask == nullptr || win == nullptr ||
      mean_logit == nullptr || max_logit == nullptr)
    return set_error(TAN_ERR_ARG, "tan_agree_scan: null pointer");
  if (B <= 0 || T <= 1 || N <= 0) return set_error(TAN_ERR_SHAPE, "tan_agree_scan: bad dims B=%d T=%d N=%d", B, T, N);
  if (workspace == nullptr || workspace_bytes < tan_agree_scan_workspace_bytes(B, T, N))
    return set_error(TAN_ERR_WORKSPACE, "tan_agree_scan: workspace too small");
  if (T > 8192) return set_error(TAN_ERR_SHAPE, "tan_agree_scan: T > 8192");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* p_scaled = static_cast<float*>(workspace);
  const int W = (N + 31) / 32;
  int blocks = (B * T + 7) / 8;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  TAN_CHECK(launch_pdl(agree_rowsoftmax_kernel, dim3(blocks), dim3(256), 0, st, 1, own, video_padding_mask,
                       text_padding_mask, B, T, N, p_scaled));
  const int wpb = T <= 1024 ? 8 : 4;
  const size_t smem = static_cast<size_t>(wpb) * T * 4;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    TAN_CUDA(cudaFuncSetAttribute(agree_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    smem_set = smem;
  }
  blocks = (B * N + wpb - 1) / wpb;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  return launch_pdl(agree_scan_kernel, dim3(blocks), dim3(wpb * 32), smem, st, 1, own, static_cast<const float*>(p_scaled),
                    posbits, video_padding_mask, text_padding_mask, B, T, N, W, fill_max, win, mean_logit, max_logit);
}

extern "C" int tan_agree_targets(const uint32_t* old_posbits, const int* win_joint, const int* win_dual,
                                 const uint8_t* replace, int B, int T, int N, int kind, uint32_t* new_posbits,
                                 void* stream) {
  TAN_CHECK(tan_device_check());
  if (old_posbits == nullptr || win_joint == nullptr || win_dual == nullptr || replace == nullptr ||
      new_posbits == nullptr)
    return set_error(TAN_ERR_ARG, "tan_agree_targets: null pointer");
  if (B <= 0 || T <= 0 || N <= 0 || kind < 0 || kind > 3)
    return set_error(TAN_ERR_SHAPE, "tan_agree_targets: bad dims / kind");
  const int W = (N + 31) / 32;
  return launch_pdl(agree_targets_kernel, dim3(B), dim3(256), static_cast<size_t>(W) * 4,
                    static_cast<cudaStream_t>(stream), 1, old_posbits, win_joint, win_dual, replace, T, N, W, kind,
                    new_posbits);
}
