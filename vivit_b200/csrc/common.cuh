// Shared helpers for libvivit_b200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstring>

#include "../../include/vivit_b200.h"

namespace vvt {

// ---- error reporting ------------------------------------------------------
char* last_error_buffer();  // thread-local, 512 bytes (defined in api_misc.cu)
extern std::atomic<int64_t> g_launches;

inline int fail(int status, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(last_error_buffer(), 512, fmt, a, b);
  return status;
}

inline int check_cuda(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return VVT_OK;
  return fail(VVT_ERR_CUDA, "%s: %s", where, cudaGetErrorString(e));
}

// call after every kernel launch
inline int launched(const char* where) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), where);
}

#define VVT_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != VVT_OK) return _s; \
  } while (0)

#define VVT_REQUIRE(cond, msg) \
  do {                         \
    if (!(cond)) return ::vvt::fail(VVT_ERR_INVALID, "%s: %s", __func__, msg); \
  } while (0)

// dispatch on dtype enum: body sees `T`
#define VVT_DISPATCH(dtype, ...)                                          \
  do {                                                                    \
    if ((dtype) == VVT_F32) {                                             \
      using T = float;                                                    \
      __VA_ARGS__                                                         \
    } else if ((dtype) == VVT_F64) {                                      \
      using T = double;                                                   \
      __VA_ARGS__                                                         \
    } else {                                                              \
      return ::vvt::fail(VVT_ERR_INVALID, "%s: unknown dtype", __func__); \
    }                                                                     \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Function attributes and SM counts are per DEVICE: one process may drive several GPUs.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev < 0 || dev >= kMaxDevices) ? 0 : dev;
}

inline int num_sms() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = current_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// Largest dynamic shared-memory size configured so far for one kernel, per device.  Usage:
//   static SmemOptIn opt;  VVT_TRY(opt.ensure(kernel, bytes, "where"));
struct SmemOptIn {
  std::atomic<size_t> done[kMaxDevices];
  template <typename K>
  int ensure(K kern, size_t bytes, const char* where) {
    const int dev = current_device();
    if (done[dev].load(std::memory_order_acquire) >= bytes) return VVT_OK;
    const cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    if (e != cudaSuccess) return fail(VVT_ERR_CUDA, "%s: %s", where, cudaGetErrorString(e));
    size_t cur = done[dev].load(std::memory_order_relaxed);
    while (cur < bytes && !done[dev].compare_exchange_weak(cur, bytes, std::memory_order_release)) {}
    return VVT_OK;
  }
};

template <typename T>
__host__ __device__ constexpr T vmax(T a, T b) { return a > b ? a : b; }
template <typename T>
__host__ __device__ constexpr T vmin(T a, T b) { return a < b ? a : b; }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

template <typename T>
__device__ __forceinline__ T ldg(const T* p) {
  return __ldg(p);
}

// warp / block reductions
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace vvt
