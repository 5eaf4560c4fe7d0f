// tan_sim_grad_gemm: the similarity recomputation of the backward pass fused with the gradient of the MIL-NCE loss.
//   G[r, c] = d loss / d cos[r, c] = e (ra[r] + cb[c] - positive(r, c) (rap[r] + cbp[c])) / 0.07,
//   e = exp((cos[r, c] - 1) / 0.07) on valid columns,  cos = <vfeat[r], tfeat[c]>
// computed in the epilogue of the tcgen05 pair GEMM (umma_gemm2.cuh) from the fp32 accumulator in TMEM and stored
// once as bf16 through the swizzled staging box + TMA store, exactly like tan_linear_bf16's bf16 mode.  Replaces
// "tan_linear_bf16 -> fp32 cosines in HBM -> tan_sim_grad_tiles" (8 bytes of HBM traffic per matrix element
// saved); the transposed copy for dB = G^T @ vfeat is made by tan_transpose_bf16.
#include "umma_gemm2.cuh"

namespace tanb {

namespace {

__device__ __forceinline__ uint32_t swz128g(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

constexpr float kInvTau = 1.0f / 0.07f;
constexpr float kK = kInvTau * 1.4426950408889634f;       // exp(x / 0.07) = exp2(x * kK)

struct SimGradEpi {
  static constexpr int kStages = 5;
  // 4 KB staging box per warp + 1 KB: the extras of warps 0 / 1 / 2 hold the tile's three 256-entry column vectors
  static constexpr int kWarpScratch = 5120;
  struct State {
    float a, ap;           // row coefficients
    int own0, own_n;       // first column / number of columns of the row's own clip
    uint32_t bits[2];      // the row's target bits (N <= 64)
    int kill;
  };
  int M, N;                // rows of the chunk, padded columns (output width)
  int C;                   // real columns
  int f_tiles, n_tiles;
  int r0, T, Ns, W, b_off;
  const float* ra;
  const float* rap;
  const float* cb;
  const float* cbp;
  const uint8_t* col_valid;
  const uint8_t* row_kill;
  const uint32_t* posbits;
  const int32_t* col_off;  // ragged columns (tan_sim_geom.col_off) or NULL

  __device__ __forceinline__ int num_tiles() const { return n_tiles; }
  __device__ __forceinline__ PairTile coord(int tile) const {
    PairTile pt;
    pt.a_row = (tile / f_tiles) * (2 * kG2BM);
    pt.b_row = (tile % f_tiles) * kG2BN;
    return pt;
  }
  __device__ __forceinline__ float* vec(uint8_t* ws, int ew, int i) const {
    return reinterpret_cast<float*>(ws + (i - ew) * kWarpScratch + 4096);
  }

  __device__ __forceinline__ void pre(int tile, uint32_t rank, int ew, int lane, uint8_t* ws, float*, uint64_t*,
                                      uint32_t, const CUtensorMap*, const CUtensorMap*, State& st) const {
    const int f_base = (tile % f_tiles) * kG2BN;
    const int et = ew * 32 + lane;
    const int c = f_base + et;
    const bool ok = c < C && col_valid[c] != 0;
    vec(ws, ew, 0)[et] = ok ? -kK : -INFINITY;             // exponent bias: e = exp2(cos * kK + bias)
    vec(ws, ew, 1)[et] = ok ? __ldg(cb + c) : 0.f;
    vec(ws, ew, 2)[et] = ok ? __ldg(cbp + c) : 0.f;
    const int rl = (tile / f_tiles) * (2 * kG2BM) + static_cast<int>(rank) * kG2BM + (ew & 3) * 32 + lane;
    st.a = 0.f; st.ap = 0.f; st.own0 = 0; st.own_n = 0; st.bits[0] = 0; st.bits[1] = 0; st.kill = 0;
    if (rl < M) {
      const int r = r0 + rl;
      const int b = r / T, t = r - b * T;
      st.a = __ldg(ra + r);
      st.ap = __ldg(rap + r);
      if (col_off != nullptr) {
        st.own0 = __ldg(col_off + b_off + b);
        st.own_n = __ldg(col_off + b_off + b + 1) - st.own0;
      } else {
        st.own0 = (b_off + b) * Ns;
        st.own_n = Ns;
      }
      const uint32_t* pw = posbits + (static_cast<int64_t>(b) * T + t) * W;
      st.bits[0] = pw[0];
      st.bits[1] = W > 1 ? pw[1] : 0u;
      st.kill = (row_kill != nullptr && row_kill[r] != 0) ? 1 : 0;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }

  __device__ __forceinline__ void run(int tile, uint32_t rank, uint32_t tmem_acc, int ew, int lane, uint8_t* ws,
                                      const float*, uint64_t*, uint32_t, const CUtensorMap* tmOut, const CUtensorMap*,
                                      State& st) const {
    const int quarter = ew & 3, half = ew >> 2;
    const int row0 = (tile / f_tiles) * (2 * kG2BM) + static_cast<int>(rank) * kG2BM + quarter * 32;
    const int col0 = (tile % f_tiles) * kG2BN + half * 128;
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
    const float4* bias4 = reinterpret_cast<const float4*>(vec(ws, ew, 0) + half * 128);
    const float4* cb4 = reinterpret_cast<const float4*>(vec(ws, ew, 1) + half * 128);
    const float* cbp1 = vec(ws, ew, 2) + half * 128;

    uint32_t r[2][32];
    tmem_ld_32x32(taddr, r[0]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      tmem_ld_wait();
      if (c + 1 < 4) tmem_ld_32x32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
      const uint32_t(&rc)[32] = r[c & 1];
      float g[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = bias4[c * 8 + j], cc = cb4[c * 8 + j];
        g[4 * j] = fast_exp2(fmaf(__uint_as_float(rc[4 * j]), kK, bb.x)) * (st.a + cc.x);
        g[4 * j + 1] = fast_exp2(fmaf(__uint_as_float(rc[4 * j + 1]), kK, bb.y)) * (st.a + cc.y);
        g[4 * j + 2] = fast_exp2(fmaf(__uint_as_float(rc[4 * j + 2]), kK, bb.z)) * (st.a + cc.z);
        g[4 * j + 3] = fast_exp2(fmaf(__uint_as_float(rc[4 * j + 3]), kK, bb.w)) * (st.a + cc.w);
      }
      // the row's own clip: positives subtract their (rap + cbp) share, killed frames lose the block
      const int n_lo = col0 + 32 * c - st.own0;              // sentence index of this chunk's first column
      if (n_lo + 32 > 0 && n_lo < st.own_n) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n_lo + j;
          if (n >= 0 && n < st.own_n) {
            if (st.kill) {
              g[j] = 0.f;
            } else if ((((n >> 5) ? st.bits[1] : st.bits[0]) >> (n & 31)) & 1u) {
              const float e = fast_exp2(fmaf(__uint_as_float(rc[j]), kK, vec(ws, ew, 0)[half * 128 + c * 32 + j]));
              g[j] -= e * (st.ap + cbp1[c * 32 + j]);
            }
          }
        }
      }
      uint32_t packed[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) packed[j] = pack_bf16x2(g[2 * j] * kInvTau, g[2 * j + 1] * kInvTau);
      if ((c & 1) == 0) {                             // the box's previous TMA store has read it out
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(ws + swz128g(lane, 4 * (c & 1) + j)) =
            make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
      if (c & 1) {                                    // box complete: 64 columns of 32 rows
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && row0 < M && col0 + 32 * (c - 1) < N) {
          tma_store_2d(tmOut, ws, col0 + 32 * (c - 1), row0);
          tma_store_commit();
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");     // the column vectors may be rewritten by the next tile's pre()
  }
};

}  // namespace

}  // namespace tanb

using namespace tanb;

extern "C" int tan_sim_grad_gemm(const void* vfeat, int64_t ldv, const void* tfeat, int64_t ldt, int Rc, int r0,
                                 const tan_sim_geom* g, int C_pad, const uint32_t* posbits, const uint8_t* col_valid,
                                 const uint8_t* row_kill, const float* ra, const float* rap, const float* cb,
                                 const float* cbp, void* G, int64_t ldg, void* stream) {
  TAN_CHECK(tan_device_check());
  if (vfeat == nullptr || tfeat == nullptr || g == nullptr || posbits == nullptr || col_valid == nullptr ||
      ra == nullptr || rap == nullptr || cb == nullptr || cbp == nullptr || G == nullptr)
    return set_error(TAN_ERR_ARG, "tan_sim_grad_gemm: null pointer");
  const int K = g->d;
  if (Rc <= 0 || r0 < 0 || r0 + Rc > g->B_loc * g->T || g->C <= 0 || C_pad < g->C || C_pad % 128 != 0 || K <= 0 ||
      K % kG2BK != 0)
    return set_error(TAN_ERR_SHAPE, "tan_sim_grad_gemm: bad shape (Rc=%d r0=%d C=%d C_pad=%d d=%d)", Rc, r0, g->C, C_pad, K);
  if (g->N > 64) return set_error(TAN_ERR_SHAPE, "tan_sim_grad_gemm: at most 64 sentences per clip (N=%d)", g->N);
  if (ldv % 8 != 0 || ldt % 8 != 0 || ldv < K || ldt < K || ldg % 8 != 0 || ldg < C_pad ||
      (reinterpret_cast<uintptr_t>(G) & 15))
    return set_error(TAN_ERR_SHAPE, "tan_sim_grad_gemm: pitches must be multiples of 8 and cover the rows; G 16-byte aligned");
  SimGradEpi e;
  e.M = Rc; e.N = C_pad; e.C = g->C;
  e.f_tiles = (C_pad + kG2BN - 1) / kG2BN;
  e.n_tiles = e.f_tiles * ((Rc + 2 * kG2BM - 1) / (2 * kG2BM));
  e.r0 = r0; e.T = g->T; e.Ns = g->N; e.W = (g->N + 31) / 32; e.b_off = g->b_off;
  e.ra = ra; e.rap = rap; e.cb = cb; e.cbp = cbp; e.col_valid = col_valid; e.row_kill = row_kill; e.posbits = posbits;
  e.col_off = g->col_off;
  CUtensorMap tmA, tmB, tmOut;
  TAN_CHECK(make_tmap_2d(&tmA, vfeat, 2, Rc, K, ldv, kG2BM));
  TAN_CHECK(make_tmap_2d(&tmB, tfeat, 2, C_pad, K, ldt, kG2BN / 2));
  TAN_CHECK(make_tmap_2d(&tmOut, G, 2, Rc, C_pad, ldg, 32));
  return launch_umma_gemm2<SimGradEpi>(tmA, tmB, tmOut, tmOut, e, e.n_tiles, K / kG2BK, static_cast<cudaStream_t>(stream));
}
