// Fused variants of backward.cu kernels.
//
// tan_transpose_colsum_bf16: tan_transpose_bf16 that also produces the column sums of its input -- the bias gradient
// of the same dY whose transpose feeds the weight-gradient GEMM -- so dY is read once instead of twice (48 tan_colsum
// calls, ~2.5 ms per training step at the bench shape, disappear).  Per 64-row tile a partial [tile][C] is written
// and finished in a fixed order (deterministic).
//
// STATUS: EXPERIMENTAL, NOT YET RUN ON A GPU (written after round 1's GPU budget was spent); train.py uses it only
// with TAN_FUSE_BIAS_SUM=1.
#include "common.cuh"

namespace tanb {

namespace {

__global__ void __launch_bounds__(256) transpose_colsum_kernel(const bf16* __restrict__ in, int64_t ldi,
                                                               bf16* __restrict__ out, int64_t ldo, int R, int C, int Rp,
                                                               float* __restrict__ partial) {
  __shared__ uint16_t tile[64][66];
  __shared__ float red[8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const uint16_t* src = reinterpret_cast<const uint16_t*>(in);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + 2 * tx;
    uint32_t v = 0;
    if (r < R && c < C) v = *reinterpret_cast<const uint32_t*>(src + static_cast<int64_t>(r) * ldi + c);
    tile[ty + 8 * i][2 * tx] = static_cast<uint16_t>(v & 0xffffu);
    tile[ty + 8 * i][2 * tx + 1] = static_cast<uint16_t>(v >> 16);
    const float2 f = unpack_bf16x2(v);           // zero beyond the matrix
    s0 += f.x;
    s1 += f.y;
  }
  red[ty][2 * tx] = s0;
  red[ty][2 * tx + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    const int c = c0 + threadIdx.x;
    if (c < C && r0 < R) partial[static_cast<int64_t>(blockIdx.x) * C + c] = s;
  }
  uint16_t* dst = reinterpret_cast<uint16_t*>(out);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + 2 * tx;
    if (c < C && r < Rp) {
      const uint32_t v = static_cast<uint32_t>(tile[2 * tx][ty + 8 * i]) |
                         (static_cast<uint32_t>(tile[2 * tx + 1][ty + 8 * i]) << 16);
      *reinterpret_cast<uint32_t*>(dst + static_cast<int64_t>(c) * ldo + r) = v;
    }
  }
}

// out[c] (+)= sum over the row tiles of partial[tile, c] (fixed order)
__global__ void colsum_tiles_finish_kernel(const float* __restrict__ partial, int tiles, int C, float* __restrict__ out,
                                           int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int p = 0;
  for (; p + 4 <= tiles; p += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] += partial[static_cast<int64_t>(p + u) * C + c];
  }
  for (; p < tiles; ++p) acc[0] += partial[static_cast<int64_t>(p) * C + c];
  const float s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  out[c] = accumulate ? out[c] + s : s;
}

}  // namespace

}  // namespace tanb

using namespace tanb;

extern "C" size_t tan_transpose_colsum_workspace_bytes(int R, int C) {
  return static_cast<size_t>((R + 63) / 64) * static_cast<size_t>(C) * sizeof(float);
}

extern "C" int tan_transpose_colsum_bf16(const void* in, int64_t ldi, void* out, int64_t ldo, int R, int C, int R_pad,
                                         float* colsum, int accumulate, void* workspace, size_t workspace_bytes,
                                         void* stream) {
  TAN_CHECK(tan_device_check());
  if (in == nullptr || out == nullptr || colsum == nullptr)
    return set_error(TAN_ERR_ARG, "tan_transpose_colsum_bf16: null pointer");
  if (R <= 0 || C <= 0 || C % 2 != 0 || R_pad < R || R_pad % 2 != 0 || ldi < C || ldo < R_pad || ldi % 2 != 0 ||
      ldo % 2 != 0)
    return set_error(TAN_ERR_SHAPE, "tan_transpose_colsum_bf16: need even C / R_pad / pitches, R_pad >= R (R=%d C=%d R_pad=%d)",
                     R, C, R_pad);
  if ((reinterpret_cast<uintptr_t>(in) & 3) || (reinterpret_cast<uintptr_t>(out) & 3))
    return set_error(TAN_ERR_SHAPE, "tan_transpose_colsum_bf16: pointers must be 4-byte aligned");
  if (workspace == nullptr || workspace_bytes < tan_transpose_colsum_workspace_bytes(R, C))
    return set_error(TAN_ERR_WORKSPACE, "tan_transpose_colsum_bf16: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the grid covers R_pad rows (zero padding of the transposed copy); only the tiles that hold real rows write partials
  dim3 grid((R_pad + 63) / 64, (C + 63) / 64);
  float* partial = static_cast<float*>(workspace);
  transpose_colsum_kernel<<<grid, 256, 0, st>>>(static_cast<const bf16*>(in), ldi, static_cast<bf16*>(out), ldo, R, C,
                                                R_pad, partial);
  TAN_CUDA(cudaGetLastError());
  colsum_tiles_finish_kernel<<<(C + 255) / 256, 256, 0, st>>>(partial, (R + 63) / 64, C, colsum, accumulate);
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}
