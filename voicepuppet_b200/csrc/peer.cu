// Device-side completion flags for the peer-memory gather: after a rank's resolve kernel has stored
// its frames into rank 0's buffer over NVLink, a one-thread kernel publishes a step counter in rank 0's
// memory (system-scope release); rank 0 enqueues a one-warp kernel that waits until every rank's flag
// has reached the step (system-scope acquire).  Together they order rank 0's consumer behind every
// rank's stores without a host round trip or an NCCL launch.
#include <cstdlib>

#include "common.h"

namespace vp {

__device__ unsigned int g_peer_timeouts = 0;  // waits that gave up (vp_peer_timeouts)

__global__ void peer_signal_kernel(unsigned int* flag, unsigned int value) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// flags[i] >= value for all i < n (n <= 32).  Bounded spin: when a rank never signals (it died, or is more than the
// time-out behind) the wait gives up and counts it in g_peer_timeouts instead of hanging the GPU or trapping (a trap
// would poison the CUDA context of the whole process); the host checks vp_peer_timeouts() after synchronising.
__global__ void peer_wait_kernel(const unsigned int* flags, int n, unsigned int value, unsigned long long max_spins) {
  const int i = threadIdx.x;
  if (i < n) {
    unsigned int v;
    unsigned long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
      if ((int)(v - value) >= 0) break;
      __nanosleep(200);
    } while (++spins < max_spins);
    if ((int)(v - value) < 0) atomicAdd(&g_peer_timeouts, 1u);
  }
  __syncwarp();
  __threadfence_system();
}

}  // namespace vp

extern "C" int vp_peer_signal(unsigned int* flag_dev, unsigned int value, void* stream) {
  VP_REQUIRE(flag_dev != nullptr, "null flag");
  vp::peer_signal_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(flag_dev, value);
  VP_LAUNCH_CHECK();
  return VP_OK;
}

extern "C" int vp_peer_wait(const unsigned int* flags_dev, int n, unsigned int value, void* stream) {
  VP_REQUIRE(flags_dev != nullptr && n >= 1 && n <= 32, "1 <= n <= 32 flags");
  // time-out: VPB200_PEER_TIMEOUT_S seconds (default 60; a spin is a 200 ns sleep plus a system-scope load, ~1 us)
  static const unsigned long long max_spins = [] {
    const char* e = std::getenv("VPB200_PEER_TIMEOUT_S");
    const double s = e ? std::atof(e) : 60.0;
    return (unsigned long long)((s > 0.001 ? s : 60.0) * 1.0e6);
  }();
  vp::peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(flags_dev, n, value, max_spins);
  VP_LAUNCH_CHECK();
  return VP_OK;
}

// Number of vp_peer_wait calls on the current device that gave up since the library was loaded (call after
// synchronising the stream the wait was enqueued on); < 0 on error.
extern "C" int vp_peer_timeouts(void) {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, vp::g_peer_timeouts, sizeof(v)) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1;
  }
  return (int)v;
}

// Copy-engine push of finished frames into the peer-mapped buffer (device-to-device, asynchronous).
extern "C" int vp_copy_async(void* dst_dev, const void* src_dev, size_t bytes, void* stream) {
  VP_REQUIRE(bytes == 0 || (dst_dev && src_dev), "null pointer");
  if (bytes) VP_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return VP_OK;
}
