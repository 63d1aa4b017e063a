// Shared helpers for the oscillink_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/oscillink_b200.h"

namespace osc {

// thread-local error text returned by osc_last_error()
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define OSC_CUDA(expr)                                        \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::osc::cuda_fail(_e, #expr); \
  } while (0)

#define OSC_LAUNCH_CHECK(name)                                 \
  do {                                                         \
    cudaError_t _e = cudaGetLastError();                       \
    if (_e != cudaSuccess) return ::osc::cuda_fail(_e, name);  \
  } while (0)

#define OSC_REQUIRE(cond, msg)                                          \
  do {                                                                  \
    if (!(cond)) return ::osc::fail(OSC_ERR_INVALID, std::string(msg)); \
  } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// bump allocator over a caller-owned workspace
struct Arena {
  char* base;
  size_t cap;
  size_t off = 0;
  bool ok = true;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), cap(n) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (off + bytes > cap) {
      ok = false;
      off += bytes;
      return nullptr;
    }
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// (similarity desc, column asc) ordering of graph.py:46-49
__device__ __forceinline__ bool better(float s, int j, float s2, int j2) {
  return (s > s2) || (s == s2 && j < j2);
}

int sm_count();

}  // namespace osc
