// Microbenchmark: per-SM throughput of ex2.approx.ftz.f32, cvt.rn.bf16x2.f32 and fmax on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i * 0.1f;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]) - 1.0f;
      if (MODE == 1) { __nv_bfloat162 v = __floats2bfloat162_rn(a[i], a[(i + 1) & 7]); acc ^= *reinterpret_cast<unsigned*>(&v); a[i] += 1e-3f; }
      if (MODE == 2) a[i] = fmaxf(a[i], a[(i + 3) & 7] * 0.999f);
      if (MODE == 3) a[i] = fmaf(a[i], 0.999f, 0.001f);
      // the softmax inner pattern: 2 x (ffma, ex2, fadd) + one pack of the pair -- with cvt.rn.bf16x2 (MODE 4) or with
      // integer rounding + prmt (MODE 5); ops counted = ex2 (2 per i)
      if (MODE == 4 || MODE == 5) {
        const float p0 = ex2(fmaf(a[i], 0.999f, -0.5f)), p1 = ex2(fmaf(a[(i + 1) & 7], 0.998f, -0.25f));
        a[i] = (p0 + p1) - 1.5f;
        if (MODE == 4) { __nv_bfloat162 v = __floats2bfloat162_rn(p0, p1); acc ^= *reinterpret_cast<unsigned*>(&v); }
        else { acc ^= __byte_perm(__float_as_uint(p0) + 0x8000u, __float_as_uint(p1) + 0x8000u, 0x7632); }
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
}
template <int MODE>
void run(const char* name, int opsPerIter) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, sms * 8 * 1024 * sizeof(float));
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, warps * 32>>>(out, iters);
    cudaEventRecord(e0);
    k<MODE><<<sms, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double ops = double(sms) * warps * 32 * iters * 8 * opsPerIter;
    printf("%-10s warps/SM=%2d  %.3f ms  %.2f Gop/s  %.2f ops/clk/SM @%.0f MHz nominal\n", name, warps, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1e3);
  }
}
int main() { run<0>("ex2", 1); run<1>("cvt.bf16x2", 1); run<2>("fmax+fmul", 1); run<3>("ffma", 1);
  run<4>("2ex2+cvt", 2); run<5>("2ex2+prmt", 2); return 0; }
