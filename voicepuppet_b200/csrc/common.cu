// Error reporting, launch counting and small utilities of libvpb200's C ABI.
#include "common.h"

#include <cstring>

namespace vp {

namespace {
thread_local char g_error[512] = "";
}

std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

}  // namespace vp

extern "C" const char* vp_last_error(void) { return vp::g_error; }

extern "C" int vp_version(void) { return 100; }

extern "C" int vp_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    vp::set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return -VP_ERR_CUDA;
  }
  return n;
}

extern "C" unsigned long long vp_launch_count(void) { return vp::g_launches.load(std::memory_order_relaxed); }

extern "C" int vp_host_alloc(void** out, size_t bytes) {
  VP_REQUIRE(out != nullptr, "null out pointer");
  *out = nullptr;
  if (bytes == 0) return VP_OK;
  VP_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
  return VP_OK;
}

extern "C" int vp_host_free(void* p) {
  if (p) VP_CUDA(cudaFreeHost(p));
  return VP_OK;
}
