// tan_layernorm: warp-per-row LayerNorm with the fused extras the TAN forward needs between GEMMs
// (positional add, video|text concatenation by row scatter, per-stage raw and L2-normalised feature
// emission).  HBM/L2-bound: each row is read once with 16-byte loads, kept in registers, reduced by
// warp shuffles (two-pass mean / variance, as torch), and every output is written once.
#include "common.cuh"

namespace tanb {

template <int V>   // V = d / 128 float4 per lane
__global__ void __launch_bounds__(256) layernorm_kernel(const tan_ln_args a) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int d = a.d;
  pdl_launch_dependents();
  pdl_wait();
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < a.rows; r += gridDim.x * warps_per_block) {
    float x[V * 4];
    if (a.in_is_bf16) {
      const uint2* p = reinterpret_cast<const uint2*>(static_cast<const bf16*>(a.in) + static_cast<int64_t>(r) * d);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const uint2 u = p[i * 32 + lane];
        const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
        x[4 * i] = lo.x; x[4 * i + 1] = lo.y; x[4 * i + 2] = hi.x; x[4 * i + 3] = hi.y;
      }
    } else {
      const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(a.in) + static_cast<int64_t>(r) * d);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float4 u = p[i * 32 + lane];
        x[4 * i] = u.x; x[4 * i + 1] = u.y; x[4 * i + 2] = u.z; x[4 * i + 3] = u.w;
      }
    }
    const int b = r / a.L_in;
    const int l = r - b * a.L_in;
    if (a.gamma != nullptr) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < V * 4; ++i) s += x[i];
      const float mean = warp_sum(s) / d;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < V * 4; ++i) { const float t = x[i] - mean; q += t * t; }
      const float rstd = rsqrtf(warp_sum(q) / d + 1e-5f);
      const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
      const float4* b4 = reinterpret_cast<const float4*>(a.beta);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float4 g = __ldg(g4 + i * 32 + lane), be = __ldg(b4 + i * 32 + lane);
        x[4 * i] = (x[4 * i] - mean) * rstd * g.x + be.x;
        x[4 * i + 1] = (x[4 * i + 1] - mean) * rstd * g.y + be.y;
        x[4 * i + 2] = (x[4 * i + 2] - mean) * rstd * g.z + be.z;
        x[4 * i + 3] = (x[4 * i + 3] - mean) * rstd * g.w + be.w;
      }
    }
    if (a.add != nullptr) {
      const float4* p = reinterpret_cast<const float4*>(a.add + static_cast<int64_t>(l % a.add_rows) * d);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float4 u = __ldg(p + i * 32 + lane);
        x[4 * i] += u.x; x[4 * i + 1] += u.y; x[4 * i + 2] += u.z; x[4 * i + 3] += u.w;
      }
    }
    const int64_t dst = static_cast<int64_t>(b) * a.L_out + a.l_off + l;
    if (a.out_f32 != nullptr) {
      float4* p = reinterpret_cast<float4*>(a.out_f32 + dst * d);
#pragma unroll
      for (int i = 0; i < V; ++i) p[i * 32 + lane] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
    }
    if (a.out_bf16 != nullptr) {
      uint2* p = reinterpret_cast<uint2*>(static_cast<bf16*>(a.out_bf16) + dst * d);
#pragma unroll
      for (int i = 0; i < V; ++i)
        p[i * 32 + lane] = make_uint2(pack_bf16x2(x[4 * i], x[4 * i + 1]), pack_bf16x2(x[4 * i + 2], x[4 * i + 3]));
    }
    // stage emission
    const bool partA = l < a.l_split;
    float* raw = partA ? a.rawA_f32 : a.rawB_f32;
    bf16* nb = static_cast<bf16*>(partA ? a.nrmA_bf16 : a.nrmB_bf16);
    float* nf = partA ? a.nrmA_f32 : a.nrmB_f32;
    if (raw != nullptr || nb != nullptr || nf != nullptr) {
      const int64_t srow = partA ? (static_cast<int64_t>(b) * a.strideA + l)
                                 : (static_cast<int64_t>(b) * a.strideB + (l - a.l_split));
      if (raw != nullptr) {
        const int64_t rs = partA ? a.raw_strideA : a.raw_strideB;
        const int64_t rrow = rs == 0 ? srow : (static_cast<int64_t>(b) * rs + (partA ? l : l - a.l_split));
        float4* p = reinterpret_cast<float4*>(raw + rrow * d);
#pragma unroll
        for (int i = 0; i < V; ++i) p[i * 32 + lane] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
      }
      if (nb != nullptr || nf != nullptr) {
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < V * 4; ++i) q += x[i] * x[i];
        const float inv = 1.0f / sqrtf(warp_sum(q));      // x / x.norm(), no eps (model/tan_model.py:116)
        if (nf != nullptr) {
          float4* p = reinterpret_cast<float4*>(nf + srow * d);
#pragma unroll
          for (int i = 0; i < V; ++i)
            p[i * 32 + lane] = make_float4(x[4 * i] * inv, x[4 * i + 1] * inv, x[4 * i + 2] * inv, x[4 * i + 3] * inv);
        }
        if (nb != nullptr) {
          uint2* p = reinterpret_cast<uint2*>(nb + srow * d);
#pragma unroll
          for (int i = 0; i < V; ++i)
            p[i * 32 + lane] = make_uint2(pack_bf16x2(x[4 * i] * inv, x[4 * i + 1] * inv),
                                          pack_bf16x2(x[4 * i + 2] * inv, x[4 * i + 3] * inv));
        }
      }
    }
  }
}

}  // namespace tanb

using namespace tanb;

extern "C" int tan_layernorm(const tan_ln_args* args, void* stream) {
  TAN_CHECK(tan_device_check());
  if (args == nullptr || args->in == nullptr) return set_error(TAN_ERR_ARG, "tan_layernorm: null args/in");
  const tan_ln_args& a = *args;
  if (a.rows <= 0) return TAN_OK;
  if (a.d % 128 != 0 || a.d <= 0 || a.d > 1024)
    return set_error(TAN_ERR_SHAPE, "tan_layernorm: d must be a multiple of 128 and <= 1024 (d=%d)", a.d);
  if (a.L_in <= 0 || a.L_out < a.L_in + a.l_off || a.l_off < 0)
    return set_error(TAN_ERR_SHAPE, "tan_layernorm: bad row map (L_in=%d L_out=%d l_off=%d)", a.L_in, a.L_out, a.l_off);
  if ((a.gamma == nullptr) != (a.beta == nullptr)) return set_error(TAN_ERR_ARG, "tan_layernorm: gamma/beta mismatch");
  if (a.add != nullptr && a.add_rows <= 0) return set_error(TAN_ERR_ARG, "tan_layernorm: add_rows must be > 0");
  const int warps = 8;
  int blocks = (a.rows + warps - 1) / warps;
  const int cap = num_sms() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a.d / 128) {
    case 1: TAN_CHECK(launch_pdl(layernorm_kernel<1>, dim3(blocks), dim3(warps * 32), 0, st, 1, a)); break;
    case 2: TAN_CHECK(launch_pdl(layernorm_kernel<2>, dim3(blocks), dim3(warps * 32), 0, st, 1, a)); break;
    case 3: TAN_CHECK(launch_pdl(layernorm_kernel<3>, dim3(blocks), dim3(warps * 32), 0, st, 1, a)); break;
    case 4: TAN_CHECK(launch_pdl(layernorm_kernel<4>, dim3(blocks), dim3(warps * 32), 0, st, 1, a)); break;
    case 5: TAN_CHECK(launch_pdl(layernorm_kernel<5>, dim3(blocks), dim3(warps * 32), 0, st, 1, a)); break;
    case 6: TAN_CHECK(launch_pdl(layernorm_kernel<6>, dim3(blocks), dim3(warps * 32), 0, st, 1, a)); break;
    case 7: TAN_CHECK(launch_pdl(layernorm_kernel<7>, dim3(blocks), dim3(warps * 32), 0, st, 1, a)); break;
    default: TAN_CHECK(launch_pdl(layernorm_kernel<8>, dim3(blocks), dim3(warps * 32), 0, st, 1, a)); break;
  }
  TAN_CUDA(cudaGetLastError());
  return TAN_OK;
}
