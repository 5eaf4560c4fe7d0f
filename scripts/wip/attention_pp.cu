// NOT BUILT, NOT SHIPPED: the persistent ping-pong attention kernel of round 2, measured slower than attention.cu on
// every benchmarked shape and rejected (profiles/r02q_attention_pingpong_rejected.txt).  Kept as the record of the
// experiment; it was compiled in place of csrc/attention.cu (non-rdc object, see the setmaxnreg note below).
//
// tan_attention_bf16: multi-head softmax attention core on tcgen05 tensor cores, head_dim 64, arbitrary
// key-padding mask, Lq != Lk allowed (cross-attention).  Replaces torch's nn.MultiheadAttention core
// (model/tfm_model.py:30-32 of the reference: softmax(q k^T / 8 + key_padding_mask) v per head).
//
// PERSISTENT kernel, one CTA per SM, 640 threads = 5 warpgroups.  A task is (clip, head, PAIR of 128-query tiles A | B); a CTA
// walks its tasks (task = blockIdx.x + i * gridDim.x) and, inside each, the 128-key blocks.  Every pipeline keeps
// running across task boundaries (flat block counter s = i * nb + j), so TMEM allocation, barrier set-up and the
// first-load latency are paid once per CTA, and a task's last PV / output epilogue overlaps the next task's
// first QK / softmax.
//   warp 0      TMA producer: per task the Q pair (double buffered by task parity) and the key-mask bits of the
//               clip; per block K (3-stage ring, freed by the pair's last QK) and V (3-stage ring, freed by the
//               pair's last PV).  K / V blocks are loaded ONCE per pair of query tiles.
//   warps 1, 2  MMA issuers of tile A / tile B:  S_X = Q_X K_s^T  (M=128, N=128, K=64; both operands K-major, 128B
//               swizzle) as soon as the softmax warps hold the previous block's scores in registers, O_X += P_X V_s
//               (M=128, N=64, K=128; V consumed as it lies in HBM: [keys, 64] = MN-major B operand, no transpose
//               anywhere) as soon as P_X is staged.  The two tiles run in PING-PONG, tile B half a softmax behind
//               tile A: the MMAs, barrier round trips, TMEM loads and the output epilogue of one tile execute
//               under the exponentials of the other.
//   warps 4-11  softmax of tile A, warps 12-19 softmax of tile B, TWO threads per query row (TMEM lane): thread
//               (half hh) owns the key columns [64 hh, 64 hh + 64) of every block -- 64 scores in registers from two
//               tcgen05.ld.x32, four softmax warps per scheduler (the warpgroups trade registers with setmaxnreg:
//               64 for the TMA / MMA warpgroup, 104 for the four softmax warpgroups; the sum must stay within the 640 x 96 registers of the launch).  The halves of a row agree on
//               the block maximum through shared memory and a 64-thread named barrier; P = exp2(S - m) goes back to
//               shared memory as the bf16 A operand of PV (each half writes one 128-byte row of ITS 64-key half
//               tile, 128B-swizzled K-major: the layout TMA would have produced).
// TMEM (512 columns): S_A [0,128), S_B [128,256), O_A [256,320), O_B [320,384).  O ACCUMULATES IN TMEM across the
// key blocks and is read once per task.  The reference max m of a row is only raised when a block's maximum
// exceeds it by more than 8 (log2 domain, i.e. P <= 256: harmless for bf16 P and fp32 sums); only then the warp
// rescales its 32 rows of O in TMEM -- after the first block this almost never happens.
// Why this shape (round-2 per-CTA timelines of the previous kernel, profiles/r02i_attention_timeline_before.txt:
// one (clip, head) per CTA, 64-key blocks, two CTAs per SM): the exp2 of a block (MUFU, 16 / clk / SM) bounds
// head_dim-64 attention at 512 clk per 128 x 64 scores, the old kernel spent ~1000 clk per block and SM plus
// ~12 k of each CTA's 26 k clk in prologue and tile epilogues: four barrier round trips of the single MMA thread
// per 64 keys, K / V loaded once per query tile.  Here a block is 128 keys (half the round trips), the other
// tile's softmax hides them, and the prologue is paid once per SM.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace tanb {

constexpr int kPpBQ = 128;                     // query rows per tile (UMMA M)
constexpr int kPpBK = 128;                     // keys per block (UMMA N of QK, K of PV)
constexpr int kPpThreads = 640;                // 5 warpgroups: producer / MMA issuers, 4 x softmax
constexpr int kPpSoftmaxRegs = 104;            // setmaxnreg: 64 for warpgroup 0, 104 for the softmax warpgroups (640 x 96 at launch)
constexpr int kPpTile = 128 * 128;             // bytes: 128 rows x 64 bf16 (16 KB): a Q tile, a K / V block, half a P tile
constexpr int kPpStages = 3;
constexpr int kPpMaskWords = 64;               // mask bits for Lk <= 2048 (longer: per-block ballots)
constexpr int kPpSmem = 2 * 2 * kPpTile /*Q pair x 2*/ + 2 * kPpStages * kPpTile /*K, V rings*/ +
                        2 * 2 * kPpTile /*P_A, P_B*/ + 512 /*barriers*/ + 2 * kPpMaskWords * 4 + 2 * 1024 /*row exchange*/;
static_assert(kPpSmem <= 227 * 1024, "one CTA per SM");

// warpgroup-wide register reallocation (all 128 threads of the warpgroup execute it)
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// one lane of the (converged) warp; ptxas keeps the code under `if (elect_one())` on the uniform datapath
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t pp_swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// Rare path of the lazy online softmax, kept out of line so that its 32 staging registers do not add to the
// pressure of the block loop (the caller holds 128 scores): multiply this warp's 32 rows of O (2 x 32 columns)
// by alpha in TMEM.
__device__ __noinline__ void pp_rescale(uint32_t t_o, float alpha) {
  uint32_t a0[32];
#pragma unroll 1
  for (int part = 0; part < 2; ++part) {
    const uint32_t t_x = t_o + part * 32;
    tmem_ld_32x32(t_x, a0);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 32; ++c) a0[c] = __float_as_uint(__uint_as_float(a0[c]) * alpha);
    tmem_st_32x32(t_x, a0);
  }
  tmem_st_wait();
}

struct PpArgs {
  const uint8_t* kpm;       // [B, Lk] 1 = ignore key, or null
  bf16* out;
  int64_t ldo;
  float* lse;               // [B, H, lse_ld] log2-domain log-sum-exp per query row (backward pass), or null
  int lse_ld;
  int H, Lq, Lk;
  int npairs;               // query-tile pairs per (clip, head)
  int ntasks;               // B * H * npairs
  long long* trace;         // development aid (tan_debug_set_trace): 256 clock stamps per CTA, or null
};

// Trace slots of a CTA: 0 globaltimer, 1 start, 2 prologue done, 3 exit; flat block s < 14 at 8 + 16 s:
// +0 K issued, +1 V issued (producer) | +2 QK_A, +3 QK_B issued, +4 p_ready_A seen, +5 PV_A issued, +6 p_ready_B
// seen, +7 PV_B issued (MMA warp) | +8 s_full seen, +9 S in registers, +10 P staged (warp 4, tile A), +11..+13 the
// same for tile B (warp 8) | +14 / +15 task epilogue of tile A / B done (stamped at the task's last block)
// Bounded wait of this kernel: on a time-out (a protocol bug) the waiter leaves (site, value, parity) in slot
// 200 + warp of the CTA's trace record -- readable after the trap when the trace buffer is mapped host memory --
// and traps (surfacing as a CUDA error instead of a hung GPU).
#ifndef TAN_ATT_WAIT
#define TAN_ATT_WAIT 1      // 0: try_wait with a suspend-time hint, 1: plain try_wait, 2: test_wait spin (A/B builds)
#endif
__device__ __forceinline__ bool pp_poll(uint64_t* bar, uint32_t parity) {
#if TAN_ATT_WAIT == 0
  return mbar_try_wait_hint(bar, parity);
#elif TAN_ATT_WAIT == 1
  return mbar_try_wait(bar, parity);
#else
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
#endif
}
__device__ __forceinline__ void pp_wait(uint64_t* bar, uint32_t parity, long long* trace, int site, int val) {
  if (pp_poll(bar, parity)) return;
  const long long t0 = clock64();
  int spins = 0;
  while (!pp_poll(bar, parity)) {
    if ((++spins & 255) == 0 && clock64() - t0 > (1ll << 29)) {
      if (trace != nullptr) {
        trace[static_cast<int64_t>(blockIdx.x) * 256 + 232 + (threadIdx.x >> 5)] =
            (static_cast<long long>(site) << 40) | (static_cast<long long>(val & 0xffffff) << 8) | parity;
        __threadfence_system();
      }
      __trap();
    }
  }
}

// (compiled in with -DTAN_ATT_TRACE only: the stamps cost instructions in the softmax loop)
__device__ __forceinline__ void pp_trace(long long* tr, int s, int k) {
#ifdef TAN_ATT_TRACE
  if (tr != nullptr && s < 14) tr[8 + 16 * s + k] = clock64();
#endif
}

__global__ void __launch_bounds__(kPpThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const PpArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();     // the 128-byte swizzle atoms need 1024-byte aligned tiles
  uint8_t* sQ = smem;                               // [2 task parities][A | B][16 KB]
  uint8_t* sK = sQ + 4 * kPpTile;                   // [3][16 KB]
  uint8_t* sV = sK + kPpStages * kPpTile;           // [3][16 KB]
  uint8_t* sP = sV + kPpStages * kPpTile;           // [A | B][2 half tiles][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kPpTile);
  uint64_t* q_full = bars;             // [2] by task parity
  uint64_t* q_empty = bars + 2;        // [2] the task's last QK has completed
  uint64_t* k_full = bars + 4;         // [3]
  uint64_t* k_empty = bars + 7;        // [3]
  uint64_t* v_full = bars + 10;        // [3]
  uint64_t* v_empty = bars + 13;       // [3]
  uint64_t* s_full = bars + 16;        // [A | B] QK of the tile's current block has completed
  uint64_t* p_ready = bars + 18;       // [A | B] count 8 (softmax warps of the tile): P staged (O rescaled if needed)
  uint64_t* pv_done = bars + 20;       // [A | B] PV of the tile's current block has completed: P free, O stable
  uint64_t* o_free = bars + 22;        // [A | B] count 8: the task's O has been read out of TMEM
  uint64_t* mask_free = bars + 24;     // [2] by task parity, count 16: every softmax warp is done with the task
  uint64_t* s_free = bars + 26;        // [A | B] count 8: the block's scores are in registers, S may be overwritten
  uint64_t* turn = bars + 28;          // [A | B] count 8: the OTHER tile has finished a block's exponentials, this one may start
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(bars + 64);     // [2 task parities][kPpMaskWords]
  uint8_t* s_xch = reinterpret_cast<uint8_t*>(s_mask + 2 * kPpMaskWords);   // [tile][1 KB]: block maxima bf16 [parity][half][row], or sums fp32 [half][row]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* const tr = (a.trace != nullptr && lane == 0) ? a.trace + static_cast<int64_t>(blockIdx.x) * 256 : nullptr;
  if (tr != nullptr && threadIdx.x == 0) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tr[0] = gt;
    tr[1] = clock64();
  }
  const int nb = (a.Lk + kPpBK - 1) / kPpBK;
  const int G = static_cast<int>(gridDim.x);
  const int ncta = (static_cast<int>(blockIdx.x) < a.ntasks) ? (a.ntasks - 1 - static_cast<int>(blockIdx.x)) / G + 1 : 0;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 2);
        mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 8); mbar_init(&pv_done[i], 1); mbar_init(&o_free[i], 8);
        mbar_init(&mask_free[i], 16);
        mbar_init(&s_free[i], 8);
      }
      mbar_init(&turn[0], 8);
      mbar_init(&turn[1], 8);
      for (int i = 0; i < kPpStages; ++i) {
        mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  if (tr != nullptr && threadIdx.x == 0) tr[2] = clock64();

  // task i of this CTA -> (clip, head, pair); heads fastest, so the 8 heads that share a clip's packed rows are in
  // flight on neighbouring SMs at the same time
  auto decode = [&](int i, int& b, int& h, int& qp) {
    const int t = static_cast<int>(blockIdx.x) + i * G;
    h = t % a.H;
    const int bp = t / a.H;
    qp = bp % a.npairs;
    b = bp / a.npairs;
  };
  auto has_b = [&](int qp) { return qp * 2 * kPpBQ + kPpBQ < a.Lq; };

  // each role's code must be DOMINATED by its setmaxnreg (ptxas allocates a region with the count of the
  // instruction that dominates it; after a merge point it would fall back to the smaller one)
  if (warp < 4) {
  reg_dealloc<64>();
  if (warp == 0) {
    // ===== TMA producer (+ the clip's key mask as bits) =====
    for (int i = 0; i < ncta; ++i) {
      int b, h, qp;
      decode(i, b, h, qp);
      const int tp = i & 1;
      if (i >= 2) pp_wait(&mask_free[tp], ((i >> 1) - 1) & 1, a.trace, 1, i);
      {
        const uint8_t* mb = a.kpm != nullptr ? a.kpm + static_cast<int64_t>(b) * a.Lk : nullptr;
        uint32_t* mk = s_mask + tp * kPpMaskWords;
        const int words = min(nb * 4, kPpMaskWords);
#pragma unroll 4
        for (int wd = 0; wd < words; ++wd) {
          const int key = wd * 32 + lane;
          const bool ig = key >= a.Lk || (mb != nullptr && mb[key] != 0);
          const uint32_t bits = __ballot_sync(0xffffffffu, ig);
          if (lane == 0) mk[wd] = bits;
        }
      }
      __syncwarp();
      if (lane == 0) {
        if (i >= 2) pp_wait(&q_empty[tp], ((i >> 1) - 1) & 1, a.trace, 2, i);
        const int q0 = qp * 2 * kPpBQ;
        const bool hb = has_b(qp);
        mbar_arrive_expect_tx(&q_full[tp], hb ? 2 * kPpTile : kPpTile);     // publishes the mask words as well
        tma_load_2d(sQ + tp * 2 * kPpTile, &tmQ, &q_full[tp], h * 64, b * a.Lq + q0);
        if (hb) tma_load_2d(sQ + tp * 2 * kPpTile + kPpTile, &tmQ, &q_full[tp], h * 64, b * a.Lq + q0 + kPpBQ);
        for (int j = 0; j < nb; ++j) {
          const int s = i * nb + j;
          const int st = s % kPpStages;
          const uint32_t ph = ((s / kPpStages) & 1) ^ 1;
          pp_wait(&k_empty[st], ph, a.trace, 3, s);
          mbar_arrive_expect_tx(&k_full[st], kPpTile);
          tma_load_2d(sK + st * kPpTile, &tmK, &k_full[st], h * 64, b * a.Lk + j * kPpBK);
          pp_trace(tr, s, 0);
          pp_wait(&v_empty[st], ph, a.trace, 4, s);
          mbar_arrive_expect_tx(&v_full[st], kPpTile);
          tma_load_2d(sV + st * kPpTile, &tmV, &v_full[st], h * 64, b * a.Lk + j * kPpBK);
          pp_trace(tr, s, 1);
        }
      }
      __syncwarp();
    }
  } else if (warp <= 2) {
    // ===== MMA issuers: warp 1 serves tile A, warp 2 tile B =====
    // Each issuer walks ITS tile's blocks in the fixed order  QK_X(s+1) [as soon as S_X(s) is in registers]
    // PV_X(s) [as soon as P_X(s) is staged]; the two tiles never wait for each other (one issuer serving both
    // in a fixed order re-synchronised the warpgroups: profiles/r02q).  K / V / Q buffers are released by BOTH
    // issuers (barrier count 2), also in tasks without tile B.
    const int X = warp - 1;
    constexpr uint32_t idesc_qk = umma_idesc_bf16(kPpBQ, kPpBK);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(kPpBQ, 64) | (1u << 16);     // B (= V) is MN-major
    uint32_t pvc = 0;                  // blocks of this tile whose PV has been issued
    uint32_t qkc = 0;                  // blocks of this tile whose QK has been issued
    uint32_t tsk = 0;                  // tasks of this tile whose first PV has been issued
    const int S = ncta * nb;
    auto pair_of = [&](int i) { int b, h, qp; decode(i, b, h, qp); return qp; };
    // all operand tiles are 16 KB apart: Q 0..3 (task parity x tile), K 4..6, V 7..9, P 10..13 (tile x 64-key half)
    const uint64_t d0 = umma_desc_k_sw128(smem_u32(smem));
    constexpr uint32_t kTileDesc = kPpTile >> 4;

    // S_X = Q_X K_s^T.  Issued as soon as the softmax warps hold the previous block's scores in registers (s_free),
    // i.e. a whole softmax ahead of its consumer.
    auto issue_qk = [&](int s) {
      const int i = s / nb, j = s - i * nb;
      const int tp = i & 1, st = s % kPpStages;
      if (qkc > 0) pp_wait(&s_free[X], (qkc - 1) & 1, a.trace, 5, s * 2 + X);
      if (j == 0) pp_wait(&q_full[tp], (i >> 1) & 1, a.trace, 6, i);
      pp_wait(&k_full[st], (s / kPpStages) & 1, a.trace, 7, s * 2 + X);
      tc_fence_after();
      const uint64_t dq = d0 + static_cast<uint32_t>(tp * 2 + X) * kTileDesc;
      const uint64_t dk = d0 + static_cast<uint32_t>(4 + st) * kTileDesc;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + X * 128, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
        pp_trace(tr, s, 2 + X);
        tc_commit(&s_full[X]);
        tc_commit(&k_empty[st]);
        if (j == nb - 1) tc_commit(&q_empty[tp]);
      }
      __syncwarp();
      ++qkc;
    };
    auto issue_pv = [&](int s) {
      const int i = s / nb, j = s - i * nb;
      const int st = s % kPpStages;
      pp_wait(&p_ready[X], pvc & 1, a.trace, 8, s * 2 + X);        // P staged (O rescaled if needed)
      pp_trace(tr, s, 4 + 2 * X);
      if (j == 0 && tsk > 0) pp_wait(&o_free[X], (tsk - 1) & 1, a.trace, 9, s * 2 + X);   // the previous task has left O
      pp_wait(&v_full[st], (s / kPpStages) & 1, a.trace, 10, s * 2 + X);
      tc_fence_after();
      const uint64_t dp = d0 + static_cast<uint32_t>(10 + 2 * X) * kTileDesc;
      const uint64_t dv = d0 + static_cast<uint32_t>(7 + st) * kTileDesc;
      // 16 keys per MMA: +32 B along P's rows inside a 64-key half tile (K-major; the next half tile is
      // 16 KB further), +16 rows x 128 B = 2048 B in V (MN-major)
      const int ksteps = min(8, (a.Lk - j * kPpBK + 15) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < ksteps)
            umma_bf16_ss(tmem_base + 256 + X * 64, dp + (k >> 2) * kTileDesc + 2 * (k & 3), dv + 128 * k, idesc_pv,
                         k != 0 || j != 0);
        pp_trace(tr, s, 5 + 2 * X);
        tc_commit(&pv_done[X]);
        tc_commit(&v_empty[st]);
      }
      __syncwarp();
      ++pvc;
      if (j == 0) ++tsk;
    };

    int prev = -1;
    for (int s = 0; s < S; ++s) {
      if (X == 1 && !has_b(pair_of(s / nb))) {
        // A task without tile B.  First the PV this tile still owes (its softmax warps wait for it in their task
        // epilogue, and the producer waits for them before it loads the Q tiles two tasks ahead).  Then this
        // issuer still OBSERVES every phase of the shared full barriers and ARRIVES on the empty ones: a waiter
        // that skips phases, or that the ring may overtake by two phases, would alias the parity of its next wait.
        if (prev >= 0) {
          issue_pv(prev);
          prev = -1;
        }
        const int i = s / nb, j = s - i * nb, st = s % kPpStages;
        if (j == 0) pp_wait(&q_full[i & 1], (i >> 1) & 1, a.trace, 6, i);
        pp_wait(&k_full[st], (s / kPpStages) & 1, a.trace, 7, s * 2 + X);
        pp_wait(&v_full[st], (s / kPpStages) & 1, a.trace, 10, s * 2 + X);
        if (lane == 0) {
          mbar_arrive(&k_empty[st]);
          mbar_arrive(&v_empty[st]);
          if (j == nb - 1) mbar_arrive(&q_empty[i & 1]);
        }
        __syncwarp();
        continue;
      }
      issue_qk(s);
      if (prev >= 0) issue_pv(prev);
      prev = s;
    }
    if (prev >= 0) issue_pv(prev);
  }
  } else {
    reg_alloc<kPpSoftmaxRegs>();
    // ===== softmax / output: warps 4..11 tile A, warps 12..19 tile B; TWO threads per query row =====
    // Thread (quarter, lane) of half hh owns row quarter * 32 + lane and the key columns [64 hh, 64 hh + 64) of
    // every block: four softmax warps per scheduler instead of two (the two-warp version spent a third of a block
    // in its TMEM-load / mask / maximum phase with the MUFU pipe idle, profiles/r02q), 64 scores in registers
    // instead of 128.  The halves of a row exchange their block maxima (and, once per task, their sums) through
    // shared memory and a 64-thread named barrier per (tile, quarter).
    const int X = (warp - 4) >> 3;
    const int hh = ((warp - 4) >> 2) & 1;
    const int quarter = warp & 3;                      // TMEM lane quarter this warp may address
    const int row = quarter * 32 + lane;               // query row inside the tile = TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + X * 128 + hh * 64;
    const uint32_t t_o = t_lane + 256 + X * 64 + hh * 32;
    uint8_t* sPh = sP + (X * 2 + hh) * kPpTile;        // this half's 64 keys = one 128-byte row of its half tile
    bf16* xmax = reinterpret_cast<bf16*>(s_xch + X * 1024);       // [parity][half][row]
    float* xsum = reinterpret_cast<float*>(s_xch + X * 1024);     // [half][row] (same bytes, used between tasks)
    const int pair_bar = 1 + X * 4 + quarter;          // named barrier of the two warps that share these rows
    const float sl2 = 0.125f * 1.4426950408889634f;    // 1/sqrt(64) folded with log2(e)
    uint32_t cnt = 0;                                  // blocks of this tile processed so far
    uint32_t tcnt = 0;                                 // turns at the MUFU pipe taken so far (incl. the passed ones)
    // PING-PONG of the exponentials: the tiles take strict turns A B A B ... at the MUFU pipe (turn[X] = the other
    // tile has finished its block).  Left alone, the two tiles fall into lock step -- every warp in the same phase,
    // the MUFU pipe idle during all TMEM loads / maxima / stores / barrier round trips (profiles/r02q); with the
    // turns one tile's exponentials run under the other tile's everything-else.  A tile without work (no tile B in
    // the task, no real rows in the warp) still passes its turns.
    auto take_turn = [&]() {
      if (X == 1 || tcnt > 0) pp_wait(&turn[X], (X == 1 ? tcnt : tcnt - 1) & 1, a.trace, 17, tcnt * 2 + X);
    };
    auto pass_turn = [&]() {
      __syncwarp();
      if (lane == 0) mbar_arrive(&turn[X ^ 1]);
      ++tcnt;
    };
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
    // the row's block maximum over both halves, as the SAME bf16-rounded value in both (any reference close to the
    // maximum serves the softmax; the halves must only agree on it).  Double buffered by block parity.
    auto row_max = [&](float v, uint32_t par) {
      const bf16 mine = __float2bfloat16_rn(v);
      xmax[(par * 2 + hh) * 128 + row] = mine;
      pair_sync();
      return fmaxf(__bfloat162float(mine), __bfloat162float(xmax[(par * 2 + (hh ^ 1)) * 128 + row]));
    };
    auto row_sum = [&](float v) {                      // once per task, no block exchange in flight
      xsum[hh * 128 + row] = v;
      pair_sync();
      const float other = xsum[(hh ^ 1) * 128 + row];
      pair_sync();                                     // read before the next task's maxima reuse the bytes
      return v + other;
    };

    for (int i = 0; i < ncta; ++i) {
      int b, h, qp;
      decode(i, b, h, qp);
      const int tp = i & 1;
      pp_wait(&q_full[tp], (i >> 1) & 1, a.trace, 12, i * 2 + X);   // mask words visible; waited for by EVERY warp in
      if (X == 1 && !has_b(qp)) {                          // EVERY task: a skipped phase would alias the parity
        for (int j = 0; j < nb; ++j) {
          take_turn();
          pass_turn();
        }
        if (lane == 0) mbar_arrive(&mask_free[tp]);
        continue;
      }
      const uint32_t* mk = s_mask + tp * kPpMaskWords;
      const bool mask_in_smem = nb * 4 <= kPpMaskWords;
      const uint8_t* mb = a.kpm != nullptr ? a.kpm + static_cast<int64_t>(b) * a.Lk : nullptr;
      const int q0 = qp * 2 * kPpBQ + X * kPpBQ;
      const bool live = q0 + quarter * 32 < a.Lq;      // warp-uniform: this warp pair owns at least one real query row
      float m_ref = -INFINITY, l_run = 0.f;            // l_run: this half's share of the row sum

      for (int j = 0; j < nb; ++j) {
        pp_wait(&s_full[X], cnt & 1, a.trace, 13, (i * nb + j) * 2 + X);
        tc_fence_after();
        long long* const trw = (quarter == 0 && hh == 0) ? tr : nullptr;
        const int s = i * nb + j;
        pp_trace(trw, s, 8 + 3 * X);
        if (live) {
          // one instance per number of 32-key chunks of THIS half that hold a real key (2 except in a sequence's
          // last block): a chunk without a real key is never touched (no exponentials, no P -- the PV MMA stops at
          // the last real key)
          auto block = [&](auto nch_c) {
            constexpr int nch = decltype(nch_c)::value;
            uint32_t r[nch > 0 ? nch : 1][32];
#pragma unroll
            for (int c2 = 0; c2 < nch; ++c2) tmem_ld_32x32(t_s + c2 * 32, r[c2]);
            if (nch > 0) tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[X]);    // the next block's QK may overwrite S
            pp_trace(trw, s, 9 + 3 * X);
            // key mask (warp-uniform words, bit set = ignore key) and this half's row maximum
            float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int c2 = 0; c2 < nch; ++c2) {
              uint32_t w;
              if (mask_in_smem) {
                w = mk[4 * j + 2 * hh + c2];
              } else {
                const int key = j * kPpBK + (2 * hh + c2) * 32 + lane;
                w = __ballot_sync(0xffffffffu, key >= a.Lk || (mb != nullptr && mb[key] != 0));
              }
              if (w != 0u) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if ((w >> c) & 1u) r[c2][c] = 0xff800000u;           // -inf
              }
#pragma unroll
              for (int c = 0; c < 32; c += 2)
                mxa[(c >> 1) & 3] = fmaxf(mxa[(c >> 1) & 3], fmaxf(__uint_as_float(r[c2][c]), __uint_as_float(r[c2][c + 1])));
            }
            const float mx_own = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])) * sl2;   // sl2 > 0: -inf stays -inf
            const float mx = row_max(mx_own, cnt & 1);                       // the row's maximum over the block
            // lazy reference max: raise it only when this block exceeds it by more than 2^8 (both halves of a row
            // see the same mx and m_ref, so they decide alike)
            const bool need = (j == 0) ? (mx > m_ref) : (mx > m_ref + 8.0f);
            if (j > 0 && __any_sync(0xffffffffu, need)) {
              // rescale this warp's rows and columns of O (and l) to the new reference; the previous PV must have completed
              pp_wait(&pv_done[X], (cnt - 1) & 1, a.trace, 14, (i * nb + j) * 2 + X);
              tc_fence_after();
              const float alpha = need ? fast_exp2(m_ref - mx) : 1.f;    // m_ref = -inf -> 0 (O and l are 0 then)
              pp_rescale(t_o, alpha);
              l_run *= alpha;
              if (need) m_ref = mx;
            } else if (need) {
              m_ref = mx;                              // first block of the task
            }
            const float nm = (m_ref == -INFINITY) ? 0.f : -m_ref;
            // all exponentials first, packed in registers: the wait for the P buffer (the previous block's PV, a
            // barrier round trip through the MMA issuer) comes AFTER them, right before the first store
            uint32_t pk[nch > 0 ? nch : 1][16];
            float ls0 = 0.f, ls1 = 0.f;
            take_turn();
#pragma unroll
            for (int c2 = 0; c2 < nch; ++c2) {
#pragma unroll
              for (int c = 0; c < 16; c += 2) {
                const float p0 = fast_exp2(fmaf(__uint_as_float(r[c2][2 * c]), sl2, nm));
                const float p1 = fast_exp2(fmaf(__uint_as_float(r[c2][2 * c + 1]), sl2, nm));
                const float p2 = fast_exp2(fmaf(__uint_as_float(r[c2][2 * c + 2]), sl2, nm));
                const float p3 = fast_exp2(fmaf(__uint_as_float(r[c2][2 * c + 3]), sl2, nm));
                ls0 += p0 + p1;
                ls1 += p2 + p3;
                pk[c2][c] = pack_bf16x2(p0, p1);       // keys 2c, 2c+1 of the chunk
                pk[c2][c + 1] = pack_bf16x2(p2, p3);
              }
            }
            l_run += ls0 + ls1;
            pass_turn();
            // the P buffer of this tile was the A operand of the previous block's PV
            if (cnt >= 1) pp_wait(&pv_done[X], (cnt - 1) & 1, a.trace, 15, (i * nb + j) * 2 + X);
            // P row of this half: 64 keys = 128 bytes = 8 chunks of 8 keys
#pragma unroll
            for (int c2 = 0; c2 < nch; ++c2) {
#pragma unroll
              for (int ch = 0; ch < 4; ++ch)
                *reinterpret_cast<uint4*>(sPh + pp_swz(row, c2 * 4 + ch)) =
                    make_uint4(pk[c2][4 * ch], pk[c2][4 * ch + 1], pk[c2][4 * ch + 2], pk[c2][4 * ch + 3]);
            }
            if (nch > 0) fence_proxy_async_smem();     // P visible to the tensor core (async proxy)
            pp_trace(trw, s, 10 + 3 * X);
          };
          const int nchh = min(2, max(0, ((a.Lk - j * kPpBK + 31) >> 5) - 2 * hh));
          switch (nchh) {
            case 2: block(std::integral_constant<int, 2>{}); break;
            case 1: block(std::integral_constant<int, 1>{}); break;
            default: block(std::integral_constant<int, 0>{}); break;
          }
        } else {
          if (lane == 0) mbar_arrive(&s_free[X]);
          take_turn();
          pass_turn();
          // a warp without real rows must not run ahead of the others: with S released early it could arrive on
          // p_ready for block j + 1 while a live warp still owes its arrival for block j (and every warp waits for
          // every phase of pv_done, so that its parity never aliases)
          if (cnt >= 1) pp_wait(&pv_done[X], (cnt - 1) & 1, a.trace, 15, (i * nb + j) * 2 + X);
        }
        tc_fence_before();                             // S reads / O writes retired before the MMAs behind p_ready
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[X]);
        ++cnt;
      }
      // all MMAs of the task's tile have completed when its last PV has
      pp_wait(&pv_done[X], (cnt - 1) & 1, a.trace, 16, (i * nb + nb - 1) * 2 + X);
      tc_fence_after();
      if (live) {
        const float l_row = row_sum(l_run);            // both halves: the row's sum of P
        // log2-domain log-sum-exp of the row's scaled scores for the backward pass (rows Lq .. lse_ld get +inf, so
        // that exp2(s - lse) of a padding row is 0 there without a predicate)
        if (hh == 0 && a.lse != nullptr && q0 + row < a.lse_ld)
          a.lse[(static_cast<int64_t>(b) * a.H + h) * a.lse_ld + q0 + row] =
              q0 + row < a.Lq ? m_ref + __log2f(l_row) : INFINITY;
        float o[32];
        {
          uint32_t a0[32];
          tmem_ld_32x32(t_o, a0);                      // this half's 32 of the 64 output columns
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = __uint_as_float(a0[c]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[X]);        // the next task may overwrite this O
        // normalise, stage this warp's 32 rows x 64 bytes in ITS rows of its P half tile (their last reader, this
        // task's last PV, has completed and only this warp writes them), then write 64-byte row pieces:
        // lane = (row % 8, 16-byte chunk)
        const float inv = 1.f / l_row;                 // l == 0 (all keys masked) -> NaN, as torch
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint4 u;
          u.x = pack_bf16x2(o[8 * ch] * inv, o[8 * ch + 1] * inv);
          u.y = pack_bf16x2(o[8 * ch + 2] * inv, o[8 * ch + 3] * inv);
          u.z = pack_bf16x2(o[8 * ch + 4] * inv, o[8 * ch + 5] * inv);
          u.w = pack_bf16x2(o[8 * ch + 6] * inv, o[8 * ch + 7] * inv);
          *reinterpret_cast<uint4*>(sPh + pp_swz(row, ch)) = u;
        }
        __syncwarp();
        const int rr = lane >> 2, cc = lane & 3;
        bf16* ob = a.out + (static_cast<int64_t>(b) * a.Lq + q0 + quarter * 32) * a.ldo + h * 64 + hh * 32 + cc * 8;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int rl = 8 * t + rr;
          if (q0 + quarter * 32 + rl < a.Lq)
            *reinterpret_cast<uint4*>(ob + static_cast<int64_t>(rl) * a.ldo) =
                *reinterpret_cast<const uint4*>(sPh + pp_swz(quarter * 32 + rl, cc));
        }
        __syncwarp();                                  // the staged rows are read before the next block's P overwrites them
      } else {
        if (hh == 0 && a.lse != nullptr && q0 + row < a.lse_ld)
          a.lse[(static_cast<int64_t>(b) * a.H + h) * a.lse_ld + q0 + row] = INFINITY;
        if (lane == 0) mbar_arrive(&o_free[X]);
      }
      if (lane == 0) mbar_arrive(&mask_free[tp]);
      pp_trace((quarter == 0 && hh == 0) ? tr : nullptr, i * nb + nb - 1, 14 + X);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tr != nullptr && threadIdx.x == 0) tr[3] = clock64();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tanb

using namespace tanb;

extern "C" int tan_attention_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                  int64_t ldv, const uint8_t* key_padding_mask, void* out, int64_t ldo, int B, int H,
                                  int Lq, int Lk, float* lse, void* stream) {
  TAN_CHECK(tan_device_check());
  if (q == nullptr || k == nullptr || v == nullptr || out == nullptr)
    return set_error(TAN_ERR_ARG, "tan_attention_bf16: null pointer");
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0 || B > 65535 || H > 65535)
    return set_error(TAN_ERR_SHAPE, "tan_attention_bf16: bad dims B=%d H=%d Lq=%d Lk=%d", B, H, Lq, Lk);
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8 || ldq < H * 64 || ldk < H * 64 || ldv < H * 64 || ldo < H * 64)
    return set_error(TAN_ERR_SHAPE, "tan_attention_bf16: row pitches must cover H*64 columns and be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(out)) & 15)
    return set_error(TAN_ERR_SHAPE, "tan_attention_bf16: q/k/v/out must be 16-byte aligned");
  if (static_cast<int64_t>(B) * Lq > 0x7fffffffll || static_cast<int64_t>(B) * Lk > 0x7fffffffll)
    return set_error(TAN_ERR_SHAPE, "tan_attention_bf16: B*L exceeds the TMA coordinate range");
  const int npairs = (Lq + 2 * kPpBQ - 1) / (2 * kPpBQ);
  const int64_t ntasks = static_cast<int64_t>(B) * H * npairs;
  if (ntasks > 0x7fffffffll) return set_error(TAN_ERR_SHAPE, "tan_attention_bf16: too many (clip, head, tile) tasks");
  CUtensorMap tmQ, tmK, tmV;
  TAN_CHECK(make_tmap_2d(&tmQ, q, 2, static_cast<uint64_t>(B) * Lq, static_cast<uint64_t>(H) * 64, ldq, kPpBQ));
  TAN_CHECK(make_tmap_2d(&tmK, k, 2, static_cast<uint64_t>(B) * Lk, static_cast<uint64_t>(H) * 64, ldk, kPpBK));
  TAN_CHECK(make_tmap_2d(&tmV, v, 2, static_cast<uint64_t>(B) * Lk, static_cast<uint64_t>(H) * 64, ldv, kPpBK));
  TAN_CHECK(set_max_dyn_smem(reinterpret_cast<const void*>(attention_kernel), kPpSmem));
  PpArgs a;
  a.kpm = key_padding_mask;
  a.out = static_cast<bf16*>(out);
  a.ldo = ldo;
  a.lse = lse;
  a.lse_ld = (Lq + 63) / 64 * 64;
  a.H = H;
  a.Lq = Lq;
  a.Lk = Lk;
  a.npairs = npairs;
  a.ntasks = static_cast<int>(ntasks);
  a.trace = debug_trace_ptr();
  const int grid = static_cast<int>(ntasks < num_sms() ? ntasks : num_sms());
  return launch_pdl(attention_kernel, dim3(grid), dim3(kPpThreads), kPpSmem, static_cast<cudaStream_t>(stream), 1, tmQ,
                    tmK, tmV, a);
}
