// Text embedder of the reference (model/word2vec_model.py:76-102, the step right before the hot path: SURVEY.md 8(f)
// f3): Embedding(66250 x 300) gather -> fc1 (300 -> 2048) + ReLU -> masked max-pool over the (up to) 32 words of a
// sentence -> fc2 (2048 -> 512).
//
//   tan_embed_gather_bf16   rows of the (frozen) bf16 embedding table -> the token matrix [S*32, K] (K = 320: 300
//                           padded to a multiple of 64 with zero columns), one warp per token
//   tan_text_pool_fc1       the fc1 GEMM on the tcgen05 pair GEMM (umma_gemm2.cuh) with the pooling FUSED into its
//                           epilogue: a 256-token tile is 8 sentences x 32 words and an epilogue warp's 32 TMEM lanes
//                           are exactly ONE sentence, so bias + ReLU + the `-6e4` fill of ignored words (:94) +
//                           max over the words is a transposing warp butterfly per 32-column chunk.  The
//                           [S*32, 2048] activations (1 GB at 8192 sentences) never reach HBM; the arg-max word per
//                           (sentence, feature) is kept for the backward pass (1 byte each).
//   tan_text_pool_bwd       d pooled -> dH [S*32, 2048] bf16 (the arg-max word gets the gradient where the ReLU is
//                           open, :87,:95), the operand of dW1 = dH^T X on tan_gemm_tn_bf16.
// fc2 and its gradients run on tan_linear_bf16 / tan_gemm_tn_bf16.  The "all words ignored" rule (:93: a sentence
// of stop words only keeps all of its words) is applied per sentence inside the epilogue.
#include <algorithm>

#include "umma_gemm2.cuh"

namespace tanb {

namespace {

constexpr int kWords = 32;

__global__ void __launch_bounds__(256)
embed_gather_kernel(const int64_t* __restrict__ ids, const uint4* __restrict__ table, int64_t ld16, int V, int64_t n,
                    uint4* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = w0; r < n; r += nw) {
    int64_t id = ids[r];
    if (id < 0 || id >= V) id = 0;                  // out-of-vocabulary -> the padding row, as the tokenizer does (:44-47)
    const uint4* src = table + id * ld16;
    uint4* dst = out + r * ld16;
    for (int64_t c = lane; c < ld16; c += 32) dst[c] = __ldg(src + c);
  }
}

// lane l receives max over the warp's 32 rows of column l (v[j] = value(row = lane, column j) on entry)
__device__ __forceinline__ float warp_column_max(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float keep = upper ? v[j + half] : v[j];
      const float send = upper ? v[j] : v[j + half];
      v[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, half));
    }
  }
  return v[0];
}

struct PoolEpi {
  static constexpr int kStages = 6;
  static constexpr int kWarpScratch = 1024;     // unused (no staging): the outputs are 64-byte row segments
  struct State {};
  int M, N;                 // tokens (S * 32), features
  int f_tiles, n_tiles;
  const float* bias;
  const uint8_t* keep;      // [M] attention mask (1 = keep the word) or NULL
  bf16* pooled;             // [S, N]
  uint8_t* argmax;          // [S, N] or NULL

  __device__ __forceinline__ int num_tiles() const { return n_tiles; }
  __device__ __forceinline__ PairTile coord(int tile) const {
    PairTile pt;
    pt.a_row = (tile / f_tiles) * (2 * kG2BM);
    pt.b_row = (tile % f_tiles) * kG2BN;
    return pt;
  }
  __device__ __forceinline__ void pre(int tile, uint32_t, int ew, int lane, uint8_t*, float* colvec, uint64_t*, uint32_t,
                                      const CUtensorMap*, const CUtensorMap*, State&) const {
    const int f_base = (tile % f_tiles) * kG2BN;
    const int et = ew * 32 + lane;
    colvec[et] = (bias != nullptr && f_base + et < N) ? __ldg(bias + f_base + et) : 0.f;
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }
  __device__ __forceinline__ void run(int tile, uint32_t rank, uint32_t tmem_acc, int ew, int lane, uint8_t*,
                                      const float* colvec, uint64_t*, uint32_t, const CUtensorMap*, const CUtensorMap*,
                                      State&) const {
    const int quarter = ew & 3, half = ew >> 2;
    const int row0 = (tile / f_tiles) * (2 * kG2BM) + static_cast<int>(rank) * kG2BM + quarter * 32;   // one sentence
    const int col0 = (tile % f_tiles) * kG2BN + half * 128;
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
    const int row = row0 + lane;
    // this word's keep flag; a sentence without any kept word keeps all of them (model/word2vec_model.py:93)
    const bool k_raw = row < M && (keep == nullptr || keep[row] != 0);
    const uint32_t any = __ballot_sync(0xffffffffu, k_raw);
    const bool kp = (any == 0u) ? (row < M) : k_raw;
    const int64_t sent = row0 / kWords;
    uint32_t r[2][32];
    tmem_ld_32x32(taddr, r[0]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      tmem_ld_wait();
      if (c + 1 < 4) tmem_ld_32x32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
      const uint32_t(&rc)[32] = r[c & 1];
      float v[32], m[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = fmaxf(__uint_as_float(rc[j]) + colvec[half * 128 + c * 32 + j], 0.f);
        v[j] = kp ? x : -6e4f;
        m[j] = v[j];
      }
      const float mx = warp_column_max(m, lane);           // lane l: max of column l over the sentence's words
      int arg = 0;
      if (argmax != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float mj = __shfl_sync(0xffffffffu, mx, j);
          const uint32_t eq = __ballot_sync(0xffffffffu, v[j] == mj);
          if (lane == j) arg = __ffs(eq) - 1;               // first word that attains the maximum
        }
      }
      const int col = col0 + 32 * c + lane;
      if (row0 < M && col < N) {
        pooled[sent * N + col] = __float2bfloat16(mx);
        if (argmax != nullptr) argmax[sent * N + col] = static_cast<uint8_t>(arg);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");        // colvec is rewritten by the next tile's pre()
  }
};

// dH[s * 32 + w, f] = (w == argmax[s, f] && pooled[s, f] > 0) ? dpool[s, f] : 0.   CTA = one sentence; a thread owns
// 8 consecutive features (16-byte stores), loops over the 32 words.
__global__ void __launch_bounds__(256)
text_pool_bwd_kernel(const bf16* __restrict__ dpool, const bf16* __restrict__ pooled, const uint8_t* __restrict__ argmax,
                     int F, bf16* __restrict__ dH) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t s = blockIdx.x;
  for (int f0 = threadIdx.x * 8; f0 < F; f0 += blockDim.x * 8) {
    const uint4 g4 = *reinterpret_cast<const uint4*>(dpool + s * F + f0);
    const uint4 p4 = *reinterpret_cast<const uint4*>(pooled + s * F + f0);
    const uint2 a2 = *reinterpret_cast<const uint2*>(argmax + s * F + f0);
    const uint32_t gs[4] = {g4.x, g4.y, g4.z, g4.w}, ps[4] = {p4.x, p4.y, p4.z, p4.w};
    uint16_t gv[8];
    uint8_t av[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 p = unpack_bf16x2(ps[i]);
      gv[2 * i] = p.x > 0.f ? static_cast<uint16_t>(gs[i] & 0xffffu) : 0;
      gv[2 * i + 1] = p.y > 0.f ? static_cast<uint16_t>(gs[i] >> 16) : 0;
      av[2 * i] = static_cast<uint8_t>(((i < 2 ? a2.x : a2.y) >> (16 * (i & 1))) & 0xffu);
      av[2 * i + 1] = static_cast<uint8_t>(((i < 2 ? a2.x : a2.y) >> (16 * (i & 1) + 8)) & 0xffu);
    }
    for (int w = 0; w < kWords; ++w) {
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        o[i] = static_cast<uint32_t>(av[2 * i] == w ? gv[2 * i] : 0) |
               (static_cast<uint32_t>(av[2 * i + 1] == w ? gv[2 * i + 1] : 0) << 16);
      *reinterpret_cast<uint4*>(dH + (s * kWords + w) * F + f0) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

}  // namespace

}  // namespace tanb

using namespace tanb;

extern "C" int tan_embed_gather_bf16(const int64_t* ids, const void* table, int64_t ld, int V, int64_t n, void* out,
                                     void* stream) {
  TAN_CHECK(tan_device_check());
  if (ids == nullptr || table == nullptr || out == nullptr) return set_error(TAN_ERR_ARG, "tan_embed_gather_bf16: null pointer");
  if (n <= 0 || V <= 0 || ld <= 0 || ld % 8 != 0 || (reinterpret_cast<uintptr_t>(table) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(TAN_ERR_SHAPE, "tan_embed_gather_bf16: need n, V > 0, ld %% 8 == 0, 16-byte aligned table / out");
  const int blocks = static_cast<int>(std::min<int64_t>((n + 7) / 8, 16ll * num_sms()));
  return launch_pdl(embed_gather_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1, ids,
                    static_cast<const uint4*>(table), ld / 8, V, n, static_cast<uint4*>(out));
}

extern "C" int tan_text_pool_fc1(const void* x, int64_t ldx, const void* w1, int64_t ldw, const float* b1,
                                 const uint8_t* keep, int S, int F, int K, void* pooled, uint8_t* argmax, void* stream) {
  TAN_CHECK(tan_device_check());
  if (x == nullptr || w1 == nullptr || pooled == nullptr) return set_error(TAN_ERR_ARG, "tan_text_pool_fc1: null pointer");
  if (S <= 0 || F <= 0 || K <= 0 || K % kG2BK != 0 || F % 128 != 0 || ldx % 8 != 0 || ldw % 8 != 0 || ldx < K || ldw < K)
    return set_error(TAN_ERR_SHAPE, "tan_text_pool_fc1: need S > 0, F %% 128 == 0, K %% 64 == 0 (S=%d F=%d K=%d)", S, F, K);
  if (static_cast<int64_t>(S) * kWords > 0x7fffffffll) return set_error(TAN_ERR_SHAPE, "tan_text_pool_fc1: too many tokens");
  const int M = S * kWords;
  CUtensorMap tmA, tmB;
  TAN_CHECK(make_tmap_2d(&tmA, x, 2, M, K, ldx, kG2BM));
  TAN_CHECK(make_tmap_2d(&tmB, w1, 2, F, K, ldw, kG2BN / 2));
  PoolEpi e;
  e.M = M; e.N = F; e.f_tiles = (F + kG2BN - 1) / kG2BN; e.n_tiles = e.f_tiles * ((M + 2 * kG2BM - 1) / (2 * kG2BM));
  e.bias = b1; e.keep = keep; e.pooled = static_cast<bf16*>(pooled); e.argmax = argmax;
  return launch_umma_gemm2<PoolEpi>(tmA, tmB, tmA, tmA, e, e.n_tiles, K / kG2BK, static_cast<cudaStream_t>(stream));
}

extern "C" int tan_text_pool_bwd(const void* dpool, const void* pooled, const uint8_t* argmax, int S, int F, void* dH,
                                 void* stream) {
  TAN_CHECK(tan_device_check());
  if (dpool == nullptr || pooled == nullptr || argmax == nullptr || dH == nullptr)
    return set_error(TAN_ERR_ARG, "tan_text_pool_bwd: null pointer");
  if (S <= 0 || F <= 0 || F % 8 != 0) return set_error(TAN_ERR_SHAPE, "tan_text_pool_bwd: need S > 0, F %% 8 == 0");
  return launch_pdl(text_pool_bwd_kernel, dim3(S), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                    static_cast<const bf16*>(dpool), static_cast<const bf16*>(pooled), argmax, F, static_cast<bf16*>(dH));
}
