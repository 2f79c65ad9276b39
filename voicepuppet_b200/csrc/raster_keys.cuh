// Z-buffer key flavours and the packed flat colour, shared by the rasterizer kernels (raster.cuh) and the fused
// vertex + raster kernel (fused.cu).
#pragma once

#include <cuda_runtime.h>

#include "vp_math.cuh"

namespace vp {

// Key flavours.  FullKey: 32-bit depth code | 32-bit inverted triangle index, compared against keys
// initialised from the caller's depth buffer (the mesh_core_cython entry points).  EpochKey: the
// fused pipeline's variant -- [epoch | depth code | inverted index in `tri_bits` bits]; a key
// written by an earlier chunk carries a smaller epoch and loses every atomicMax, so the z-buffer
// never has to be cleared between chunks (the resolve pass treats a stale epoch as background).
struct FullKey {
  __device__ __forceinline__ unsigned long long make(float d, uint32_t tri) const { return make_key(d, tri); }
};
struct EpochKey {
  unsigned long long epoch_field;  // epoch << (32 + tri_bits)
  uint32_t tri_mask;               // (1 << tri_bits) - 1
  int tri_bits;
  __device__ __forceinline__ unsigned long long make(float d, uint32_t tri) const {
    return epoch_field | (static_cast<unsigned long long>(depth_code(d)) << tri_bits) |
           static_cast<unsigned long long>(tri_mask - tri);
  }
};

EpochKey make_epoch_key(int ntri, uint32_t epoch);  // raster.cu

// (a + b + c) / 3 per byte lane of three packed RGBx words; sums <= 765, so x * 0x5556 >> 16 == x / 3.
__device__ __forceinline__ uint32_t flat_color_packed(uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t rb = (a & 0x00FF00FFu) + (b & 0x00FF00FFu) + (c & 0x00FF00FFu);  // R | B << 16 (10 bits each)
  const uint32_t g = ((a >> 8) & 0xFFu) + ((b >> 8) & 0xFFu) + ((c >> 8) & 0xFFu);
  const uint32_t r3 = ((rb & 0xFFFFu) * 0x5556u) >> 16;
  const uint32_t b3 = ((rb >> 16) * 0x5556u) & 0xFFFF0000u;
  const uint32_t g3 = ((g * 0x5556u) >> 8) & 0xFF00u;
  return r3 | g3 | b3 | 0xFF000000u;
}

}  // namespace vp
