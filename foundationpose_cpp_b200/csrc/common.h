// Shared host-side helpers: error reporting for the C ABI, CUDA call checking.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only: ranges cost a pointer check unless a profiler injects itself
#include <stdexcept>
#include <string>

namespace fpb {

// Last error message of the calling thread (returned by fpb200_last_error()).
inline std::string& last_error() {
  static thread_local std::string e;
  return e;
}

inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define FPB_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      char _b[512];                                                                         \
      snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                \
               cudaGetErrorString(_e));                                                     \
      throw ::fpb::CudaError(_b);                                                           \
    }                                                                                       \
  } while (0)

#define FPB_REQUIRE(cond, ...)                                                              \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      char _b[512];                                                                         \
      snprintf(_b, sizeof(_b), __VA_ARGS__);                                                \
      throw std::invalid_argument(std::string(_b));                                         \
    }                                                                                       \
  } while (0)

// NVTX range around a host-side stage (Nsight Systems / ncu --nvtx): the reference has no tracing at all.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define FPB_NVTX_CAT2(a, b) a##b
#define FPB_NVTX_CAT(a, b) FPB_NVTX_CAT2(a, b)
#define FPB_NVTX(name) ::fpb::NvtxRange FPB_NVTX_CAT(_fpb_nvtx_, __LINE__)(name)

// The CUDA runtime loads a kernel lazily at its first launch; that load is a device-wide operation which waits for
// running kernels (seen as a 2 s stall of every stream when the first Register of one session met the spinning exchange
// wait of another).  Querying the attributes at create time loads the function then.
template <typename... K>
inline cudaError_t preload_kernels(K... kernels) {
  cudaError_t err = cudaSuccess;
  cudaFuncAttributes a;
  const cudaError_t rs[] = {cudaFuncGetAttributes(&a, kernels)...};
  for (cudaError_t r : rs)
    if (r != cudaSuccess) err = r;
  return err;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace fpb
