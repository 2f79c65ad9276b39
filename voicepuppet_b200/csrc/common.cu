// Error reporting, launch counting and small utilities of libvpb200's C ABI.
#include "common.h"

#include <cstring>
#include <cstdint>

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

// ---- peer memory (CUDA IPC): lets every rank's resolve kernel store its frames straight into rank 0's
// buffer over NVLink, so the gather needs no separate copy --------------------------------------------
#include <cuda.h>

namespace {
typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
GetAddressRangeFn address_range_fn() {
  static GetAddressRangeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<GetAddressRangeFn>(p);
  }();
  return fn;
}
}  // namespace

extern "C" int vp_ipc_export(const void* dev_ptr, unsigned char* handle64, unsigned long long* offset) {
  VP_REQUIRE(dev_ptr && handle64 && offset, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  GetAddressRangeFn range = address_range_fn();
  VP_REQUIRE(range != nullptr, "cuMemGetAddressRange unavailable");
  CUdeviceptr base = 0;
  size_t size = 0;
  if (range(&base, &size, (CUdeviceptr)(uintptr_t)dev_ptr) != CUDA_SUCCESS) {
    vp::set_error("cuMemGetAddressRange failed for %p", dev_ptr);
    return VP_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  VP_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base)));
  std::memcpy(handle64, &h, 64);
  *offset = (unsigned long long)((uintptr_t)dev_ptr - (uintptr_t)base);
  return VP_OK;
}

extern "C" int vp_ipc_open(const unsigned char* handle64, int device, void** base_out) {
  VP_REQUIRE(handle64 && base_out, "null argument");
  VP_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  VP_CUDA(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
  return VP_OK;
}

extern "C" int vp_ipc_close(void* base) {
  if (base) VP_CUDA(cudaIpcCloseMemHandle(base));
  return VP_OK;
}
