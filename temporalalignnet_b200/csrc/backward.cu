// Backward-pass kernels of the TAN hot path (training step): everything that is NOT a GEMM or attention.
// The GEMM-shaped part of the backward pass runs on the tcgen05 pair GEMMs (dgrad = dY @ W and the similarity
// recomputation: tan_linear_bf16 / tan_sim_grad_gemm; wgrad = dY^T @ X and dB = G^T V: tan_gemm_tn_bf16 on
// MN-major operands), the attention backward in attention_bwd_tc.cu (tcgen05); this file provides the bf16
// transpose (weight shadows), the bias / LayerNorm parameter reductions, QuickGELU, LayerNorm, L2-normalisation,
// the similarity-gradient tile kernel (N > 64) and the entry point of the attention backward.
#include <cstdlib>

#include "common.cuh"
#include "attn_bwd.cuh"

namespace tanb {

// ------------------------------------------------------------------------------------------------
// bf16 transpose with zero padding of the contraction tail: out[c, r] = in[r, c] for r < R, 0 for R <= r < Rp.
// 64 x 64 tiles through shared memory, 32-bit global accesses on both sides.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const bf16* __restrict__ in, int64_t ldi,
                                                             bf16* __restrict__ out, int64_t ldo, int R, int C,
                                                             int Rp) {
  __shared__ uint16_t tile[64][66];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const uint16_t* src = reinterpret_cast<const uint16_t*>(in);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + 2 * tx;
    uint32_t v = 0;
    if (r < R && c < C) v = *reinterpret_cast<const uint32_t*>(src + static_cast<int64_t>(r) * ldi + c);   // C is even
    tile[ty + 8 * i][2 * tx] = static_cast<uint16_t>(v & 0xffffu);
    tile[ty + 8 * i][2 * tx + 1] = static_cast<uint16_t>(v >> 16);
  }
  __syncthreads();
  uint16_t* dst = reinterpret_cast<uint16_t*>(out);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + 2 * tx;
    if (c < C && r < Rp) {                                     // Rp is even
      const uint32_t v = static_cast<uint32_t>(tile[2 * tx][ty + 8 * i]) |
                         (static_cast<uint32_t>(tile[2 * tx + 1][ty + 8 * i]) << 16);
      *reinterpret_cast<uint32_t*>(dst + static_cast<int64_t>(c) * ldo + r) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Column sums (bias gradients): partial[p, n] = sum over the rows of slab p of in[m, n]; finished by
// colsum_finish_kernel in a fixed order (deterministic, no atomics).
// ------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const void* __restrict__ in, int64_t ld, int M, int N,
                                                             int rows_per_slab, float* __restrict__ partial) {
  __shared__ float red[8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 64 + 2 * tx;
  const int m0 = blockIdx.y * rows_per_slab;
  const int m1 = min(m0 + rows_per_slab, M);
  float s0 = 0.f, s1 = 0.f;
  if (n < N) {
    for (int m = m0 + ty; m < m1; m += 8) {
      if (BF16) {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(static_cast<const bf16*>(in) + static_cast<int64_t>(m) * ld + n);
        const float2 f = unpack_bf16x2(u);
        s0 += f.x; s1 += f.y;
      } else {
        const float2 f = *reinterpret_cast<const float2*>(static_cast<const float*>(in) + static_cast<int64_t>(m) * ld + n);
        s0 += f.x; s1 += f.y;
      }
    }
  }
  red[ty][2 * tx] = s0;
  red[ty][2 * tx + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    const int nn = blockIdx.x * 64 + threadIdx.x;
    if (nn < N) partial[static_cast<int64_t>(blockIdx.y) * N + nn] = s;
  }
}

// The same for 16-byte aligned rows: a lane owns 8 bf16 (or 4 fp32) columns = one 16-byte load per row, a warp row is
// 512 contiguous bytes, four rows are in flight per thread (the 4-byte-per-row version above kept one small load in
// flight per thread: 0.5 of the HBM peak on the [M, 1536] / [M, 2048] bias-gradient sums of the training step).
template <bool BF16>
__global__ void __launch_bounds__(256) colsum_partial_wide_kernel(const void* __restrict__ in, int64_t ld, int M, int N,
                                                                  int rows_per_slab, float* __restrict__ partial) {
  constexpr int kC = BF16 ? 8 : 4;                 // columns per lane
  __shared__ float red[8][32 * kC];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = (blockIdx.x * 32 + tx) * kC;
  const int m0 = blockIdx.y * rows_per_slab;
  const int m1 = min(m0 + rows_per_slab, M);
  float acc[kC];
#pragma unroll
  for (int k = 0; k < kC; ++k) acc[k] = 0.f;
  if (n < N) {
    auto add = [&](const uint4& u) {
      if (BF16) {
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
        if (kC == 8) { acc[kC - 4] += c.x; acc[kC - 3] += c.y; acc[kC - 2] += d.x; acc[kC - 1] += d.y; }
      } else {
        acc[0] += __uint_as_float(u.x); acc[1] += __uint_as_float(u.y);
        acc[2] += __uint_as_float(u.z); acc[3] += __uint_as_float(u.w);
      }
    };
    const char* base = static_cast<const char*>(in) + static_cast<int64_t>(n) * (BF16 ? 2 : 4);
    const int64_t pitch = ld * (BF16 ? 2 : 4);
    int m = m0 + ty;
    for (; m + 24 < m1; m += 32) {                 // rows m, m + 8, m + 16, m + 24 of this warp
      const uint4 u0 = *reinterpret_cast<const uint4*>(base + static_cast<int64_t>(m) * pitch);
      const uint4 u1 = *reinterpret_cast<const uint4*>(base + static_cast<int64_t>(m + 8) * pitch);
      const uint4 u2 = *reinterpret_cast<const uint4*>(base + static_cast<int64_t>(m + 16) * pitch);
      const uint4 u3 = *reinterpret_cast<const uint4*>(base + static_cast<int64_t>(m + 24) * pitch);
      add(u0); add(u1); add(u2); add(u3);
    }
    for (; m < m1; m += 8) add(*reinterpret_cast<const uint4*>(base + static_cast<int64_t>(m) * pitch));
  }
#pragma unroll
  for (int k = 0; k < kC; ++k) red[ty][tx * kC + k] = acc[k];
  __syncthreads();
  for (int j = threadIdx.x; j < 32 * kC; j += 256) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][j];
    const int nn = blockIdx.x * 32 * kC + j;
    if (nn < N) partial[static_cast<int64_t>(blockIdx.y) * N + nn] = t;
  }
}

// out[n] = (accumulate ? out[n] : 0) + sum_p partial[p, n]
__global__ void colsum_finish_kernel(const float* __restrict__ partial, int P, int N, float* __restrict__ out,
                                     int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int p = 0; p < P; ++p) s += partial[static_cast<int64_t>(p) * N + n];
  out[n] = accumulate ? out[n] + s : s;
}

// ------------------------------------------------------------------------------------------------
// QuickGELU (model/tfm_model.py:11-13) forward on the stored pre-activation and its backward.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(256) quickgelu_fwd_kernel(const uint4* __restrict__ u, uint4* __restrict__ h, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 v = u[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      o[j] = pack_bf16x2(f.x * sigmoid_f(1.702f * f.x), f.y * sigmoid_f(1.702f * f.y));
    }
    h[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__device__ __forceinline__ float quickgelu_grad(float x) {
  const float s = sigmoid_f(1.702f * x);
  return s * (1.0f + 1.702f * x * (1.0f - s));
}

__global__ void __launch_bounds__(256) quickgelu_bwd_kernel(const uint4* __restrict__ dh, const uint4* __restrict__ u,
                                                            uint4* __restrict__ du, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 a = dh[i], b = u[i];
    const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 g = unpack_bf16x2(wa[j]), x = unpack_bf16x2(wb[j]);
      o[j] = pack_bf16x2(g.x * quickgelu_grad(x.x), g.y * quickgelu_grad(x.y));
    }
    du[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  Forward: y = (x - mean) * rstd * gamma + beta, row r of x -> row map(r) of y with
// map(r) = (r / L_in) * L_out + l_off + r % L_in (the concat scatter of tan_layernorm).
//   dx[r] (+)= rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy[map(r)] * gamma
//   partial dgamma / dbeta per block, finished in a fixed order by ln_param_finish_kernel.
// Warp per row, the row lives in registers (statistics are recomputed from x, as in the forward kernel).
// ------------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ gamma, float* __restrict__ dx,
                                                            int accumulate, int rows, int d, int L_in, int L_out,
                                                            int l_off, float* __restrict__ partial,
                                                            bf16* __restrict__ dx_bf16, int want_colsum) {
  __shared__ float red[8][V * 128];
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float dg[V * 4], db[V * 4], dsum[V * 4];     // dsum: column sums of the UPDATED dx (the next linear's bias gradient)
#pragma unroll
  for (int i = 0; i < V * 4; ++i) { dg[i] = 0.f; db[i] = 0.f; dsum[i] = 0.f; }
  for (int r = blockIdx.x * warps_per_block + warp; r < rows; r += gridDim.x * warps_per_block) {
    const int b = r / L_in, l = r - b * L_in;
    const int64_t yr = static_cast<int64_t>(b) * L_out + l_off + l;
    float xv[V * 4], gv[V * 4];
    const float4* px = reinterpret_cast<const float4*>(x + static_cast<int64_t>(r) * d);
    const float4* py = reinterpret_cast<const float4*>(dy + yr * d);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 a = px[i * 32 + lane], g = py[i * 32 + lane];
      xv[4 * i] = a.x; xv[4 * i + 1] = a.y; xv[4 * i + 2] = a.z; xv[4 * i + 3] = a.w;
      gv[4 * i] = g.x; gv[4 * i + 1] = g.y; gv[4 * i + 2] = g.z; gv[4 * i + 3] = g.w;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V * 4; ++i) s += xv[i];
    const float mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V * 4; ++i) { xv[i] -= mean; q += xv[i] * xv[i]; }
    const float rstd = rsqrtf(warp_sum(q) / d + 1e-5f);
    // the residual-stream gradient this row accumulates into: requested NOW, next to x and dy, so that its latency
    // is not a second, serial round trip to HBM at the end of the row (ncu r02p: 43 % of the DRAM peak, the warps
    // parked on the first use of a load)
    float4* pd = reinterpret_cast<float4*>(dx + static_cast<int64_t>(r) * d);
    float4 pv[V];
    if (accumulate) {
#pragma unroll
      for (int i = 0; i < V; ++i) pv[i] = pd[i * 32 + lane];
    }
    // gamma is re-read per row and per 4 columns (L1-resident): holding it would cost 4 V registers and, with the
    // column-sum accumulators, push the kernel from 2 CTAs per SM to 1 (measured: 3.5 -> 7.7 ms per step)
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f);
      if (gamma != nullptr) g4 = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
      const float gm[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = 4 * i + k;
        xv[e] *= rstd;                     // xhat
        dg[e] += gv[e] * xv[e];
        db[e] += gv[e];
        gv[e] *= gm[k];                    // g
        sg += gv[e];
        sgx += gv[e] * xv[e];
      }
    }
    const float mg = warp_sum(sg) / d, mgx = warp_sum(sgx) / d;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 o;
      o.x = rstd * (gv[4 * i] - mg - xv[4 * i] * mgx);
      o.y = rstd * (gv[4 * i + 1] - mg - xv[4 * i + 1] * mgx);
      o.z = rstd * (gv[4 * i + 2] - mg - xv[4 * i + 2] * mgx);
      o.w = rstd * (gv[4 * i + 3] - mg - xv[4 * i + 3] * mgx);
      if (accumulate) {
        o.x += pv[i].x; o.y += pv[i].y; o.z += pv[i].z; o.w += pv[i].w;
      }
      pd[i * 32 + lane] = o;
      if (dx_bf16 != nullptr)      // the bf16 copy the following dgrad / wgrad GEMMs read (was a separate cast pass)
        *reinterpret_cast<uint2*>(dx_bf16 + static_cast<int64_t>(r) * d + 4 * (i * 32 + lane)) =
            make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      dsum[4 * i] += o.x; dsum[4 * i + 1] += o.y; dsum[4 * i + 2] += o.z; dsum[4 * i + 3] += o.w;
    }
  }
  if (partial == nullptr) return;
  // block reduction of dgamma then dbeta over the 8 warps (column j of lane: 128 i + 4 lane + k)
  const int passes = want_colsum ? 3 : 2;
  for (int pass = 0; pass < passes; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        red[warp][i * 128 + 4 * lane + k] = pass == 0 ? dg[4 * i + k] : (pass == 1 ? db[4 * i + k] : dsum[4 * i + k]);
    __syncthreads();
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
      float t = 0.f;
      for (int w = 0; w < warps_per_block; ++w) t += red[w][j];
      partial[(static_cast<int64_t>(blockIdx.x) * 3 + pass) * d + j] = t;
    }
  }
}

// dgamma[j] += sum_p partial[p][0][j], dbeta[j] += sum_p partial[p][1][j]  (accumulated: one LayerNorm may serve
// several calls of a step, e.g. ln_video_init in the video and the joint stack)
__global__ void ln_param_finish_kernel(const float* __restrict__ partial, int blocks, int d, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ dx_colsum) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  float g = 0.f, b = 0.f, c = 0.f;
  for (int p = 0; p < blocks; ++p) {
    g += partial[(static_cast<int64_t>(p) * 3) * d + j];
    b += partial[(static_cast<int64_t>(p) * 3 + 1) * d + j];
    if (dx_colsum != nullptr) c += partial[(static_cast<int64_t>(p) * 3 + 2) * d + j];
  }
  dgamma[j] += g;
  dbeta[j] += b;
  if (dx_colsum != nullptr) dx_colsum[j] += c;
}

// ------------------------------------------------------------------------------------------------
// Backward of y = x / ||x|| (model/tan_model.py:116-117,:136-137) with row maps on both sides:
//   src row of the raw features x: (r / L_in) * src_stride + r % L_in;  of the incoming gradient g: ... * g_stride ...
//   dst row (token-major gradient buffer):            (r / L_in) * L_out + l_off + r % L_in
//   dst = (g - y * <y, g>) / ||x||        (written, or added when accumulate != 0)
// ------------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                         float* __restrict__ dst, int accumulate, int rows, int d,
                                                         int L_in, int64_t src_stride, int64_t g_stride, int L_out,
                                                         int l_off) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
    const int b = r / L_in, l = r - b * L_in;
    const int64_t sr = static_cast<int64_t>(b) * src_stride + l;
    const int64_t gr = static_cast<int64_t>(b) * g_stride + l;
    const int64_t dr = static_cast<int64_t>(b) * L_out + l_off + l;
    float xv[V * 4], gv[V * 4];
    const float4* px = reinterpret_cast<const float4*>(x + sr * d);
    const float4* pg = reinterpret_cast<const float4*>(g + gr * d);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 a = px[i * 32 + lane], c = pg[i * 32 + lane];
      xv[4 * i] = a.x; xv[4 * i + 1] = a.y; xv[4 * i + 2] = a.z; xv[4 * i + 3] = a.w;
      gv[4 * i] = c.x; gv[4 * i + 1] = c.y; gv[4 * i + 2] = c.z; gv[4 * i + 3] = c.w;
    }
    float q = 0.f, xg = 0.f;
#pragma unroll
    for (int i = 0; i < V * 4; ++i) { q += xv[i] * xv[i]; xg += xv[i] * gv[i]; }
    q = warp_sum(q);
    xg = warp_sum(xg);
    const float inv = 1.0f / sqrtf(q);
    const float k = xg / q;                    // <y, g> / ||x|| * (1 / ||x||) applied to x:  y <y,g> = x * xg / q
    float4* pd = reinterpret_cast<float4*>(dst + dr * d);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 o;
      o.x = (gv[4 * i] - xv[4 * i] * k) * inv;
      o.y = (gv[4 * i + 1] - xv[4 * i + 1] * k) * inv;
      o.z = (gv[4 * i + 2] - xv[4 * i + 2] * k) * inv;
      o.w = (gv[4 * i + 3] - xv[4 * i + 3] * k) * inv;
      if (accumulate) {
        const float4 p = pd[i * 32 + lane];
        o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
      }
      pd[i * 32 + lane] = o;
    }
  }
}

// out[l, :] (+)= sum_b in[(b * L_out + l_off + l), :]   (gradient of a table broadcast over the batch)
__global__ void __launch_bounds__(256) batch_sum_kernel(const float* __restrict__ in, float* __restrict__ out, int B,
                                                        int L, int d, int L_out, int l_off, int accumulate) {
  const int64_t n = static_cast<int64_t>(L) * d;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int l = static_cast<int>(i / d), j = static_cast<int>(i - static_cast<int64_t>(l) * d);
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += in[(static_cast<int64_t>(b) * L_out + l_off + l) * d + j];
    out[i] = accumulate ? out[i] + s : s;
  }
}

// ------------------------------------------------------------------------------------------------
// Similarity-gradient tiles.  z [Rc, ldz] fp32 cosines of rows (b, t) = r0 + row of ONE stage against all
// columns; with e = exp((z - 1) / 0.07) on valid columns (the forward's fixed shift):
//   G[r, c] = e * (ra[r] + cb[c] - pos(r, c) * (rap[r] + cbp[c])) / 0.07      = d loss / d cos[r, c]
// ra / rap = w_row / sum_all, w_row / sum_pos of the row (0 for rows that do not count), cb / cbp the same for
// columns of this stage (train/loss.py:248-256 differentiated; see loss.py for the weights).
// Writes G [Rc, ldg] bf16 and its transpose GT [C, ldgt] bf16 (zero for Rc <= r < Rcp) -- the operands of the
// two gradient GEMMs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sim_grad_kernel(const float* __restrict__ z, int64_t ldz, int Rc, int Rcp, int C,
                                                       int Cp, int r0, int T, int N, int W, int b_off,
                                                       const uint32_t* __restrict__ posbits,
                                                       const uint8_t* __restrict__ col_valid,
                                                       const uint8_t* __restrict__ row_kill,
                                                       const float* __restrict__ ra, const float* __restrict__ rap,
                                                       const float* __restrict__ cb, const float* __restrict__ cbp,
                                                       bf16* __restrict__ G, int64_t ldg, bf16* __restrict__ GT,
                                                       int64_t ldgt) {
  __shared__ uint16_t tile[64][66];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int rt = blockIdx.x * 64, ct = blockIdx.y * 64;
  constexpr float kInvTau = 1.0f / 0.07f;
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr float kK = kInvTau * kLog2e;
  // this thread's two columns: validity and coefficients are loaded once
  const int c = ct + 2 * tx;
  bool cv[2];
  float cbv[2], cbpv[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int cc = c + k;
    cv[k] = cc < C && col_valid[cc] != 0;
    cbv[k] = cv[k] ? cb[cc] : 0.f;
    cbpv[k] = cv[k] ? cbp[cc] : 0.f;
  }
  const float invT = 1.0f / static_cast<float>(T);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = rt + ty + 8 * i;              // row inside the chunk
    float gv[2] = {0.f, 0.f};
    if (rl < Rc && c < C) {
      const int r = r0 + rl;                     // row of the stage: (b, t)
      const int b = __float2int_rd((static_cast<float>(r) + 0.5f) * invT);   // exact for r < 2^21
      const int t = r - b * T;
      const float2 zz = *reinterpret_cast<const float2*>(z + static_cast<int64_t>(rl) * ldz + c);
      const float a = ra[r], ap = rap[r];
      const bool kill = row_kill != nullptr && row_kill[r] != 0;
      const int n0 = c - (b_off + b) * N;        // sentence index of column c inside the row's own clip
      const float zv[2] = {zz.x, zz.y};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int n = n0 + k;
        const bool own = n >= 0 && n < N;
        if (cv[k] && !(own && kill)) {
          const float e = fast_exp2(fmaf(zv[k], kK, -kK));
          float coef = a + cbv[k];
          if (own && ((posbits[(static_cast<int64_t>(b) * T + t) * W + (n >> 5)] >> (n & 31)) & 1u))
            coef -= ap + cbpv[k];
          gv[k] = e * coef * kInvTau;
        }
      }
    }
    const uint32_t packed = pack_bf16x2(gv[0], gv[1]);
    tile[ty + 8 * i][2 * tx] = static_cast<uint16_t>(packed & 0xffffu);
    tile[ty + 8 * i][2 * tx + 1] = static_cast<uint16_t>(packed >> 16);
    if (rl < Rc && c < Cp) *reinterpret_cast<uint32_t*>(G + static_cast<int64_t>(rl) * ldg + c) = packed;
  }
  __syncthreads();
  uint16_t* dst = reinterpret_cast<uint16_t*>(GT);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = ct + ty + 8 * i, rl = rt + 2 * tx;
    if (c < C && rl < Rcp) {
      const uint32_t v = static_cast<uint32_t>(tile[2 * tx][ty + 8 * i]) |
                         (static_cast<uint32_t>(tile[2 * tx + 1][ty + 8 * i]) << 16);
      *reinterpret_cast<uint32_t*>(dst + static_cast<int64_t>(c) * ldgt + rl) = v;
    }
  }
}

static int elementwise_blocks(size_t n, int per_block) {
  size_t blocks = (n + per_block - 1) / per_block;
  const size_t cap = static_cast<size_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace tanb

using namespace tanb;

extern "C" int tan_transpose_bf16(const void* in, int64_t ldi, void* out, int64_t ldo, int R, int C, int R_pad,
                                  void* stream) {
  TAN_CHECK(tan_device_check());
  if (in == nullptr || out == nullptr) return set_error(TAN_ERR_ARG, "tan_transpose_bf16: null pointer");
  if (R <= 0 || C <= 0 || C % 2 != 0 || R_pad < R || R_pad % 2 != 0 || ldi < C || ldo < R_pad || ldi % 2 != 0 ||
      ldo % 2 != 0)
    return set_error(TAN_ERR_SHAPE, "tan_transpose_bf16: need even C / R_pad / pitches, R_pad >= R (R=%d C=%d R_pad=%d)",
                     R, C, R_pad);
  if ((reinterpret_cast<uintptr_t>(in) & 3) || (reinterpret_cast<uintptr_t>(out) & 3))
    return set_error(TAN_ERR_SHAPE, "tan_transpose_bf16: pointers must be 4-byte aligned");
  dim3 grid((R_pad + 63) / 64, (C + 63) / 64);
  transpose_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(in), ldi, static_cast<bf16*>(out), ldo, R, C, R_pad);
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}

static int colsum_slabs(int M) {
  int slabs = (M + 511) / 512;
  if (slabs > 64) slabs = 64;
  return slabs < 1 ? 1 : slabs;
}

extern "C" size_t tan_colsum_workspace_bytes(int M, int N) {
  return static_cast<size_t>(colsum_slabs(M)) * static_cast<size_t>(N) * sizeof(float);
}

extern "C" int tan_colsum(const void* in, int in_is_bf16, int64_t ld, int M, int N, float* out, int accumulate,
                          void* workspace, size_t workspace_bytes, void* stream) {
  TAN_CHECK(tan_device_check());
  if (in == nullptr || out == nullptr) return set_error(TAN_ERR_ARG, "tan_colsum: null pointer");
  if (M <= 0 || N <= 0 || N % 2 != 0 || ld < N || ld % 2 != 0)
    return set_error(TAN_ERR_SHAPE, "tan_colsum: need M > 0, even N and pitch (M=%d N=%d)", M, N);
  if (workspace == nullptr || workspace_bytes < tan_colsum_workspace_bytes(M, N))
    return set_error(TAN_ERR_WORKSPACE, "tan_colsum: workspace too small");
  const int slabs = colsum_slabs(M);
  const int rows_per_slab = (M + slabs - 1) / slabs;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((N + 63) / 64, slabs);
  float* partial = static_cast<float*>(workspace);
  const int per_lane = in_is_bf16 ? 8 : 4;
  if (N % per_lane == 0 && ld % per_lane == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    dim3 gw((N + 32 * per_lane - 1) / (32 * per_lane), slabs);
    if (in_is_bf16)
      colsum_partial_wide_kernel<true><<<gw, 256, 0, st>>>(in, ld, M, N, rows_per_slab, partial);
    else
      colsum_partial_wide_kernel<false><<<gw, 256, 0, st>>>(in, ld, M, N, rows_per_slab, partial);
  } else if (in_is_bf16)
    colsum_partial_kernel<true><<<grid, 256, 0, st>>>(in, ld, M, N, rows_per_slab, partial);
  else
    colsum_partial_kernel<false><<<grid, 256, 0, st>>>(in, ld, M, N, rows_per_slab, partial);
  TAN_CUDA(cudaGetLastError());
  colsum_finish_kernel<<<(N + 255) / 256, 256, 0, st>>>(partial, slabs, N, out, accumulate);
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}

extern "C" int tan_quickgelu_fwd(const void* u, void* h, size_t n, void* stream) {
  TAN_CHECK(tan_device_check());
  if (u == nullptr || h == nullptr) return set_error(TAN_ERR_ARG, "tan_quickgelu_fwd: null pointer");
  if (n % 8 != 0 || (reinterpret_cast<uintptr_t>(u) & 15) || (reinterpret_cast<uintptr_t>(h) & 15))
    return set_error(TAN_ERR_SHAPE, "tan_quickgelu_fwd: n %% 8 == 0 and 16-byte aligned pointers required");
  if (n == 0) return TAN_OK;
  quickgelu_fwd_kernel<<<elementwise_blocks(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(u), static_cast<uint4*>(h), n / 8);
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}

extern "C" int tan_quickgelu_bwd(const void* dh, const void* u, void* du, size_t n, void* stream) {
  TAN_CHECK(tan_device_check());
  if (dh == nullptr || u == nullptr || du == nullptr) return set_error(TAN_ERR_ARG, "tan_quickgelu_bwd: null pointer");
  if (n % 8 != 0 || (reinterpret_cast<uintptr_t>(u) & 15) || (reinterpret_cast<uintptr_t>(dh) & 15) ||
      (reinterpret_cast<uintptr_t>(du) & 15))
    return set_error(TAN_ERR_SHAPE, "tan_quickgelu_bwd: n %% 8 == 0 and 16-byte aligned pointers required");
  if (n == 0) return TAN_OK;
  quickgelu_bwd_kernel<<<elementwise_blocks(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dh), static_cast<const uint4*>(u), static_cast<uint4*>(du), n / 8);
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}

static int ln_bwd_blocks(int rows) {
  int blocks = (rows + 7) / 8;
  const int cap = num_sms() * 2;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : blocks;
}

extern "C" size_t tan_layernorm_bwd_workspace_bytes(int rows, int d) {
  return static_cast<size_t>(ln_bwd_blocks(rows)) * 3 * static_cast<size_t>(d) * sizeof(float);
}

extern "C" int tan_layernorm_bwd(const float* dy, const float* x, const float* gamma, float* dx, int accumulate_dx,
                                 int rows, int d, int L_in, int L_out, int l_off, float* dgamma, float* dbeta,
                                 void* dx_bf16, float* dx_colsum, void* workspace, size_t workspace_bytes, void* stream) {
  TAN_CHECK(tan_device_check());
  if (dy == nullptr || x == nullptr || dx == nullptr) return set_error(TAN_ERR_ARG, "tan_layernorm_bwd: null pointer");
  if (rows <= 0) return TAN_OK;
  if (d % 128 != 0 || d <= 0 || d > 1024)
    return set_error(TAN_ERR_SHAPE, "tan_layernorm_bwd: d must be a multiple of 128 and <= 1024 (d=%d)", d);
  if (L_in <= 0 || L_out < L_in + l_off || l_off < 0)
    return set_error(TAN_ERR_SHAPE, "tan_layernorm_bwd: bad row map (L_in=%d L_out=%d l_off=%d)", L_in, L_out, l_off);
  if ((dgamma == nullptr) != (dbeta == nullptr)) return set_error(TAN_ERR_ARG, "tan_layernorm_bwd: dgamma/dbeta mismatch");
  if (dx_colsum != nullptr && dgamma == nullptr)
    return set_error(TAN_ERR_ARG, "tan_layernorm_bwd: dx_colsum needs the parameter-gradient pass (dgamma / dbeta)");
  if (dx_bf16 != nullptr && (reinterpret_cast<uintptr_t>(dx_bf16) & 7))
    return set_error(TAN_ERR_SHAPE, "tan_layernorm_bwd: dx_bf16 must be 8-byte aligned");
  const int blocks = ln_bwd_blocks(rows);
  float* partial = nullptr;
  if (dgamma != nullptr) {
    if (workspace == nullptr || workspace_bytes < tan_layernorm_bwd_workspace_bytes(rows, d))
      return set_error(TAN_ERR_WORKSPACE, "tan_layernorm_bwd: workspace too small");
    partial = static_cast<float*>(workspace);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define TAN_LNB(V)                                                                                              \
  layernorm_bwd_kernel<V><<<blocks, 256, 0, st>>>(dy, x, gamma, dx, accumulate_dx, rows, d, L_in, L_out, l_off, \
                                                  partial, static_cast<bf16*>(dx_bf16), dx_colsum != nullptr ? 1 : 0)
  switch (d / 128) {
    case 1: TAN_LNB(1); break;
    case 2: TAN_LNB(2); break;
    case 3: TAN_LNB(3); break;
    case 4: TAN_LNB(4); break;
    case 5: TAN_LNB(5); break;
    case 6: TAN_LNB(6); break;
    case 7: TAN_LNB(7); break;
    default: TAN_LNB(8); break;
  }
#undef TAN_LNB
  TAN_CUDA(cudaGetLastError());
  if (dgamma != nullptr) {
    ln_param_finish_kernel<<<(d + 255) / 256, 256, 0, st>>>(partial, blocks, d, dgamma, dbeta, dx_colsum);
    TAN_CUDA(cudaGetLastError());
  }
  return TAN_OK;
}

extern "C" int tan_l2norm_bwd(const float* x, const float* g, float* dst, int accumulate, int rows, int d, int L_in,
                              int64_t src_stride, int64_t g_stride, int L_out, int l_off, void* stream) {
  TAN_CHECK(tan_device_check());
  if (x == nullptr || g == nullptr || dst == nullptr) return set_error(TAN_ERR_ARG, "tan_l2norm_bwd: null pointer");
  if (rows <= 0) return TAN_OK;
  if (d % 128 != 0 || d <= 0 || d > 1024)
    return set_error(TAN_ERR_SHAPE, "tan_l2norm_bwd: d must be a multiple of 128 and <= 1024 (d=%d)", d);
  if (L_in <= 0 || src_stride < L_in || g_stride < L_in || L_out < L_in + l_off || l_off < 0)
    return set_error(TAN_ERR_SHAPE, "tan_l2norm_bwd: bad row maps");
  int blocks = (rows + 7) / 8;
  const int cap = num_sms() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define TAN_L2B(V) \
  l2norm_bwd_kernel<V><<<blocks, 256, 0, st>>>(x, g, dst, accumulate, rows, d, L_in, src_stride, g_stride, L_out, l_off)
  switch (d / 128) {
    case 1: TAN_L2B(1); break;
    case 2: TAN_L2B(2); break;
    case 3: TAN_L2B(3); break;
    case 4: TAN_L2B(4); break;
    case 5: TAN_L2B(5); break;
    case 6: TAN_L2B(6); break;
    case 7: TAN_L2B(7); break;
    default: TAN_L2B(8); break;
  }
#undef TAN_L2B
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}

extern "C" int tan_batch_sum(const float* in, float* out, int B, int L, int d, int L_out, int l_off, int accumulate,
                             void* stream) {
  TAN_CHECK(tan_device_check());
  if (in == nullptr || out == nullptr) return set_error(TAN_ERR_ARG, "tan_batch_sum: null pointer");
  if (B <= 0 || L <= 0 || d <= 0 || L_out < L + l_off || l_off < 0)
    return set_error(TAN_ERR_SHAPE, "tan_batch_sum: bad shape");
  batch_sum_kernel<<<elementwise_blocks(static_cast<size_t>(L) * d, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, out, B, L, d, L_out, l_off, accumulate);
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}

extern "C" int tan_sim_grad_tiles(const float* z, int64_t ldz, int Rc, int Rc_pad, int r0, const tan_sim_geom* g,
                                  const uint32_t* posbits, const uint8_t* col_valid, const uint8_t* row_kill,
                                  const float* ra, const float* rap, const float* cb, const float* cbp, void* G,
                                  int64_t ldg, void* GT, int64_t ldgt, void* stream) {
  TAN_CHECK(tan_device_check());
  if (g != nullptr && g->col_off != nullptr)
    return set_error(TAN_ERR_ARG, "tan_sim_grad_tiles: ragged columns (col_off) are only supported by tan_sim_grad_gemm");
  if (z == nullptr || g == nullptr || posbits == nullptr || col_valid == nullptr || ra == nullptr || rap == nullptr ||
      cb == nullptr || cbp == nullptr || G == nullptr || GT == nullptr)
    return set_error(TAN_ERR_ARG, "tan_sim_grad_tiles: null pointer");
  const int C = g->C;
  const int Cp = static_cast<int>(ldg);
  if (Rc <= 0 || Rc_pad < Rc || Rc_pad % 2 != 0 || C <= 0 || ldz < C + (C & 1) || ldz % 2 != 0 || ldg < C + (C & 1) ||
      ldg % 2 != 0 || ldgt < Rc_pad || ldgt % 2 != 0 || r0 < 0 || r0 + Rc > g->B_loc * g->T)
    return set_error(TAN_ERR_SHAPE, "tan_sim_grad_tiles: bad shape (Rc=%d Rc_pad=%d C=%d r0=%d)", Rc, Rc_pad, C, r0);
  dim3 grid((Rc_pad + 63) / 64, (Cp + 63) / 64);
  sim_grad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      z, ldz, Rc, Rc_pad, C, Cp, r0, g->T, g->N, (g->N + 31) / 32, g->b_off, posbits, col_valid, row_kill, ra, rap, cb,
      cbp, static_cast<bf16*>(G), ldg, static_cast<bf16*>(GT), ldgt);
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}

extern "C" int tan_attention_bwd_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                      const void* o, int64_t ldo, const void* d_out, int64_t lddo,
                                      const uint8_t* key_padding_mask, void* dq, int64_t lddq, void* dk, int64_t lddk,
                                      void* dv, int64_t lddv, float* lse, float* delta, int B, int H, int Lq, int Lk,
                                      void* stream) {
  TAN_CHECK(tan_device_check());
  if (q == nullptr || k == nullptr || v == nullptr || o == nullptr || d_out == nullptr || dq == nullptr ||
      dk == nullptr || dv == nullptr || lse == nullptr || delta == nullptr)
    return set_error(TAN_ERR_ARG, "tan_attention_bwd_bf16: null pointer");
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0 || B > 65535 || H > 65535)
    return set_error(TAN_ERR_SHAPE, "tan_attention_bwd_bf16: bad dims (B=%d H=%d Lq=%d Lk=%d)", B, H, Lq, Lk);
  const int64_t lds[8] = {ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv};
  const void* ptrs[8] = {q, k, v, o, d_out, dq, dk, dv};
  for (int i = 0; i < 8; ++i)
    if (lds[i] % 8 != 0 || lds[i] < static_cast<int64_t>(H) * 64 || (reinterpret_cast<uintptr_t>(ptrs[i]) & 15))
      return set_error(TAN_ERR_SHAPE, "tan_attention_bwd_bf16: operands need 16-byte aligned bases, pitches %% 8 == 0 and >= 64 H");
  AttnBwdArgs a;
  a.q = static_cast<const bf16*>(q); a.ldq = ldq;
  a.k = static_cast<const bf16*>(k); a.ldk = ldk;
  a.v = static_cast<const bf16*>(v); a.ldv = ldv;
  a.o = static_cast<const bf16*>(o); a.ldo = ldo;
  a.dO = static_cast<const bf16*>(d_out); a.lddo = lddo;
  a.kpm = key_padding_mask;
  a.dq = static_cast<bf16*>(dq); a.lddq = lddq;
  a.dk = static_cast<bf16*>(dk); a.lddk = lddk;
  a.dv = static_cast<bf16*>(dv); a.lddv = lddv;
  a.lse = lse; a.delta = delta;
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk;
  return attention_bwd_tc(a, static_cast<cudaStream_t>(stream));
}
