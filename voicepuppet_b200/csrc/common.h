// Host-side plumbing shared by the translation units of libvpb200: error reporting behind
// vp_last_error(), launch counting, grow-only device scratch.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/vpb200.h"

namespace vp {

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;

inline void count_launch(unsigned n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Grow-only device buffer.  Not thread safe on its own: owners serialise access.
struct DevBuf {
  void* ptr = nullptr;
  size_t cap = 0;
  int device = -1;
  cudaError_t reserve(size_t bytes, int dev) {
    if (bytes <= cap && dev == device) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    device = dev;
    if (bytes == 0) return cudaSuccess;
    size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      want = bytes;
      e = cudaMalloc(&ptr, want);
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return static_cast<T*>(ptr); }
};

}  // namespace vp

#define VP_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::vp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      (void)cudaGetLastError();                                                            \
      return VP_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define VP_LAUNCH_CHECK()                                                                  \
  do {                                                                                     \
    ::vp::count_launch();                                                                  \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ::vp::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VP_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define VP_REQUIRE(cond, msg)                                  \
  do {                                                         \
    if (!(cond)) {                                             \
      ::vp::set_error("invalid argument: %s (%s)", msg, #cond); \
      return VP_ERR_ARG;                                       \
    }                                                          \
  } while (0)

#define VP_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != VP_OK) return _rc; \
  } while (0)
