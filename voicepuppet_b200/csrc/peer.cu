// Device-side completion flags for the peer-memory gather: after a rank's resolve kernel has stored
// its frames into rank 0's buffer over NVLink, a one-thread kernel publishes a step counter in rank 0's
// memory (system-scope release); rank 0 enqueues a one-warp kernel that waits until every rank's flag
// has reached the step (system-scope acquire).  Together they order rank 0's consumer behind every
// rank's stores without a host round trip or an NCCL launch.
#include "common.h"

namespace vp {

__global__ void peer_signal_kernel(unsigned int* flag, unsigned int value) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// flags[i] >= value for all i < n (n <= 32).  Bounded spin: a missing rank traps instead of hanging the GPU.
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
    if ((int)(v - value) < 0) __trap();
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
  vp::peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(flags_dev, n, value, 20000000ull);  // ~ seconds
  VP_LAUNCH_CHECK();
  return VP_OK;
}

// Copy-engine push of finished frames into the peer-mapped buffer (device-to-device, asynchronous).
extern "C" int vp_copy_async(void* dst_dev, const void* src_dev, size_t bytes, void* stream) {
  VP_REQUIRE(bytes == 0 || (dst_dev && src_dev), "null pointer");
  if (bytes) VP_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return VP_OK;
}
