// Library plumbing behind include/tan_b200.h: error strings, device check, TMA descriptor
// encoding, dtype cast.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace tanb {

__device__ long long* g_gemm_trace = nullptr;
static long long* g_trace_host = nullptr;      // the same pointer for kernels that take it as an argument
long long* debug_trace_ptr() { return g_trace_host; }

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return TAN_OK;
  return set_error(TAN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

struct DevInfo {
  int checked = 0;   // 0 = not yet, 1 = ok, -1 = wrong arch, -2 = cuda error
  int sms = 0;
  int major = 0, minor = 0;
};
static DevInfo g_dev[64];

static DevInfo* dev_info() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  DevInfo* d = &g_dev[dev];
  if (d->checked == 0) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
      d->checked = -2;
    } else {
      d->sms = p.multiProcessorCount;
      d->major = p.major;
      d->minor = p.minor;
      d->checked = (p.major == 10) ? 1 : -1;
    }
  }
  return d;
}

int num_sms() {
  DevInfo* d = dev_info();
  return (d && d->sms > 0) ? d->sms : 148;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: remember (device, kernel)
// pairs, so that one process driving several GPUs sets it on each of them (a per-process flag would leave the
// second device at the 48 KB default and its first launch would fail).
int set_max_dyn_smem(const void* kernel, int bytes) {
  struct Entry { int dev; const void* kernel; };
  static Entry done[512];
  static int n_done = 0;
  static std::mutex mu;
  int dev = 0;
  TAN_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n_done; ++i)
    if (done[i].dev == dev && done[i].kernel == kernel) return TAN_OK;
  TAN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (n_done < 512) done[n_done++] = Entry{dev, kernel};
  return TAN_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

// elem_bytes: 2 (bf16) or 4 (fp32); the box is always 128 bytes wide (the 128-byte swizzle span).
int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error(TAN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (elem_bytes != 2 && elem_bytes != 4) return set_error(TAN_ERR_ARG, "make_tmap_2d: element size must be 2 or 4");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * elem_bytes) % 16 != 0)
    return set_error(TAN_ERR_SHAPE, "TMA operand needs a 16-byte aligned base and row pitch");
  if (box_rows == 0 || box_rows > 256) return set_error(TAN_ERR_SHAPE, "TMA box must have 1..256 rows");
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {ld * elem_bytes};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elem_bytes), box_rows};
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(TAN_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  return TAN_OK;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  if (box_cols * 2 != 128) return set_error(TAN_ERR_SHAPE, "TMA box must be 64 bf16 wide");
  return make_tmap_2d(out, base, 2, rows, cols, ld, box_rows);
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t segs, uint64_t rows, uint64_t cols,
                      uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error(TAN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (cols * 2) % 16 != 0)
    return set_error(TAN_ERR_SHAPE, "TMA operand needs a 16-byte aligned base and row pitch");
  if (box_rows == 0 || box_rows > 256) return set_error(TAN_ERR_SHAPE, "TMA box must have 1..256 rows");
  const cuuint64_t gdim[3] = {cols, rows, segs};
  const cuuint64_t gstride[2] = {cols * 2, rows * cols * 2};
  const cuuint32_t box[3] = {64, box_rows, 1};
  const cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(TAN_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", int(r));
  return TAN_OK;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TAN_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

__global__ void cast_f32_bf16_kernel(const float4* __restrict__ in, uint4* __restrict__ out, size_t n8) {
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  pdl_launch_dependents();
  pdl_wait();
  for (; i < n8; i += stride) {
    const float4 a = __ldg(in + 2 * i);
    const float4 b = __ldg(in + 2 * i + 1);
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y);
    u.y = pack_bf16x2(a.z, a.w);
    u.z = pack_bf16x2(b.x, b.y);
    u.w = pack_bf16x2(b.z, b.w);
    out[i] = u;
  }
}

}  // namespace tanb

using namespace tanb;

extern "C" int tan_abi_version(void) { return 5; }

extern "C" const char* tan_last_error_string(void) { return g_err; }

extern "C" int tan_device_check(void) {
  DevInfo* d = dev_info();
  if (d == nullptr || d->checked == -2) return set_error(TAN_ERR_CUDA, "no usable CUDA device");
  if (d->checked != 1)
    return set_error(TAN_ERR_ARCH, "device is sm_%d%d; this library only contains sm_100a (B200) code", d->major,
                     d->minor);
  return TAN_OK;
}

extern "C" int tan_debug_set_trace(void* device_buffer) {
  TAN_CHECK(tan_device_check());
  long long* p = static_cast<long long*>(device_buffer);
  TAN_CUDA(cudaMemcpyToSymbol(g_gemm_trace, &p, sizeof(p)));
  g_trace_host = p;
  return TAN_OK;
}

extern "C" int tan_cast_f32_to_bf16(const float* in, void* out, size_t n, void* stream) {
  TAN_CHECK(tan_device_check());
  if (in == nullptr || out == nullptr) return set_error(TAN_ERR_ARG, "tan_cast_f32_to_bf16: null pointer");
  if (n % 8 != 0 || (reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(TAN_ERR_SHAPE, "tan_cast_f32_to_bf16: n %% 8 == 0 and 16-byte alignment required");
  if (n == 0) return TAN_OK;
  const size_t n8 = n / 8;
  size_t blocks = (n8 + 255) / 256;
  const size_t cap = static_cast<size_t>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  return launch_pdl(cast_f32_bf16_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0,
                    static_cast<cudaStream_t>(stream), 1, reinterpret_cast<const float4*>(in),
                    reinterpret_cast<uint4*>(out), n8);
}
