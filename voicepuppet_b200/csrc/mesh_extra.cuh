// The other two exports of the reference's native module:
//   render_texture_core  utils/cython/mesh_core_cython.pyx:80-99  -> mesh_core.cpp:234-333
//   get_normal_core      utils/cython/mesh_core_cython.pyx:40-47  -> mesh_core.cpp:85-105
// Both bit-exact.  render_texture shares rasterize_triangles' z-buffer decision (border rule, interpolated
// depth, strict '>'), so it reuses the 64-bit-key scatter and adds a texel resolve; get_normal is an ORDERED
// float32 accumulation (triangle order), done as a per-vertex gather over incidence lists that a stable
// radix sort builds on the device -- float atomics would not reproduce the reference's rounding.
// (Included at the end of raster.cu: the kernels of raster.cuh are not templates, so they live in one
// translation unit.)
#pragma once

#include <cub/device/device_radix_sort.cuh>

namespace vp {

namespace {
std::mutex g_extra_mutex;
DevBuf g_extra;

struct TextureArgs {
  const float* texture;        // [tex_h][tex_w][tex_c]
  const float* tex_coords;     // [*][3]
  const int* tex_triangles;    // [ntri][3]
  int c, tex_h, tex_w, tex_c, mapping_type;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Winner's depth and texel, recomputed with the reference's float32 expressions (mesh_core.cpp:290-324).
__global__ void resolve_texture_kernel(const unsigned long long* __restrict__ keys, GenericMesh mesh, TextureArgs ta,
                                       float* __restrict__ image, float* __restrict__ depth, int h, int w) {
  const size_t npix = (size_t)h * w;
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const int t = key_triangle(keys[p]);
  if (t < 0) return;
  int ia, ib, ic;
  uint32_t id;
  mesh.indices(t, ia, ib, ic, id);
  float x0, y0, z0, x1, y1, z1, x2, y2, z2;
  uint32_t r;
  mesh.vertex(0, ia, x0, y0, z0, r);
  mesh.vertex(0, ib, x1, y1, z1, r);
  mesh.vertex(0, ic, x2, y2, z2, r);
  TriSetup s;
  tri_edges(s, x0, y0, x1, y1, x2, y2);
  float u, v, w0, w1, w2;
  pixel_uv(s, (int)(p % w), (int)(p / w), u, v);
  depth[p] = weights_depth(u, v, z0, z1, z2, w0, w1, w2);
  // texture x from the texture triangle's vertices, y from the MESH triangle's vertices (:270-272)
  const int ta0 = __ldg(ta.tex_triangles + 3 * (size_t)t), ta1 = __ldg(ta.tex_triangles + 3 * (size_t)t + 1),
            ta2 = __ldg(ta.tex_triangles + 3 * (size_t)t + 2);
  const float* tc = ta.tex_coords;
  float tx = VP_ADD(VP_ADD(VP_MUL(w0, tc[3 * (size_t)ta0]), VP_MUL(w1, tc[3 * (size_t)ta1])), VP_MUL(w2, tc[3 * (size_t)ta2]));
  float ty = VP_ADD(VP_ADD(VP_MUL(w0, tc[3 * (size_t)ia + 1]), VP_MUL(w1, tc[3 * (size_t)ib + 1])),
                    VP_MUL(w2, tc[3 * (size_t)ic + 1]));
  tx = hi2(lo2(tx, static_cast<float>(ta.tex_w - 1)), 0.0f);   // :302-303, std::min / std::max
  ty = hi2(lo2(ty, static_cast<float>(ta.tex_h - 1)), 0.0f);
  const float yd = VP_SUB(ty, floorf(ty)), xd = VP_SUB(tx, floorf(tx));
  const size_t row = (size_t)ta.tex_w * ta.tex_c;
  float* out = image + p * ta.c;
  // the reference faults on NaN texture coordinates ((int)NaN indexes out of range); the index clamps
  // below only keep such a call memory-safe
  if (ta.mapping_type == 0) {
    const int yi = clampi(trunc_x86(roundf(ty)), 0, ta.tex_h - 1), xi = clampi(trunc_x86(roundf(tx)), 0, ta.tex_w - 1);
    for (int k = 0; k < ta.c; ++k) out[k] = ta.texture[yi * row + (size_t)xi * ta.tex_c + k];
  } else {
    const int yf = clampi(trunc_x86(floorf(ty)), 0, ta.tex_h - 1), yc = clampi(trunc_x86(ceilf(ty)), 0, ta.tex_h - 1);
    const int xf = clampi(trunc_x86(floorf(tx)), 0, ta.tex_w - 1), xc = clampi(trunc_x86(ceilf(tx)), 0, ta.tex_w - 1);
    const float ixd = VP_SUB(1.0f, xd), iyd = VP_SUB(1.0f, yd);
    for (int k = 0; k < ta.c; ++k) {
      const float ul = ta.texture[yf * row + (size_t)xf * ta.tex_c + k], ur = ta.texture[yf * row + (size_t)xc * ta.tex_c + k];
      const float dl = ta.texture[yc * row + (size_t)xf * ta.tex_c + k], dr = ta.texture[yc * row + (size_t)xc * ta.tex_c + k];
      const float a = VP_MUL(VP_MUL(ul, ixd), iyd), b = VP_MUL(VP_MUL(ur, xd), iyd);
      const float c2 = VP_MUL(VP_MUL(dl, ixd), yd), d = VP_MUL(VP_MUL(dr, xd), yd);
      out[k] = VP_ADD(VP_ADD(VP_ADD(a, b), c2), d);
    }
  }
}

__global__ void iota_kernel(int* __restrict__ v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// One thread per vertex: its incidences sit contiguously in the sorted arrays, in ascending (triangle,
// corner) order because the radix sort is stable; left-to-right float32 sum from the caller's value.
__global__ void ordered_normal_sum_kernel(const int* __restrict__ sorted_vertex, const int* __restrict__ sorted_inc,
                                          const float* __restrict__ tri_normal, float* __restrict__ normal, int nver,
                                          int ninc) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nver) return;
  int lo = 0, hi = ninc;  // lower_bound of v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(sorted_vertex + mid) < v) lo = mid + 1; else hi = mid;
  }
  float nx = normal[3 * (size_t)v], ny = normal[3 * (size_t)v + 1], nz = normal[3 * (size_t)v + 2];
  for (int j = lo; j < ninc && __ldg(sorted_vertex + j) == v; ++j) {
    const int tri = __ldg(sorted_inc + j) / 3;
    nx = VP_ADD(nx, __ldg(tri_normal + 3 * (size_t)tri));
    ny = VP_ADD(ny, __ldg(tri_normal + 3 * (size_t)tri + 1));
    nz = VP_ADD(nz, __ldg(tri_normal + 3 * (size_t)tri + 2));
  }
  normal[3 * (size_t)v] = nx;
  normal[3 * (size_t)v + 1] = ny;
  normal[3 * (size_t)v + 2] = nz;
}

}  // namespace
}  // namespace vp

using namespace vp;

extern "C" int vp_render_texture_core(float* image, const float* vertices, const int* triangles, const float* texture,
                                      const float* tex_coords, const int* tex_triangles, float* depth_buffer, int nver,
                                      int tex_nver, int ntri, int h, int w, int c, int tex_h, int tex_w, int tex_c,
                                      int mapping_type) {
  VP_REQUIRE(nver >= 0 && tex_nver >= 0 && ntri >= 0, "negative element count");
  VP_REQUIRE(h > 0 && w > 0 && h <= 16384 && w <= 16384, "image size must be in 1..16384");
  VP_REQUIRE(c >= 1 && tex_h >= 1 && tex_w >= 1 && tex_c >= c, "need c >= 1, a non-empty texture and tex_c >= c");
  VP_REQUIRE(image && depth_buffer, "null output buffer");
  VP_REQUIRE(ntri == 0 || (vertices && triangles && texture && tex_coords && tex_triangles), "null mesh buffer");
  // the reference indexes without checks (out-of-range reads); here a bad index is an error
  for (size_t i = 0; i < 3 * (size_t)ntri; ++i) {
    VP_REQUIRE(triangles[i] >= 0 && triangles[i] < nver, "triangle index out of range");
    VP_REQUIRE(tex_triangles[i] >= 0 && tex_triangles[i] < tex_nver, "texture triangle index out of range");
    // mesh_core.cpp:270-272 reads tex_coords[3 * MESH index + 1]
    VP_REQUIRE(triangles[i] < tex_nver, "tex_coords must have at least as many rows as the mesh indices used (reference reads y with the mesh index)");
  }
  int device = 0;
  VP_CUDA(cudaGetDevice(&device));
  std::lock_guard<std::mutex> lock(g_extra_mutex);
  const size_t npix = (size_t)h * w;
  Carver cv;
  const size_t o_keys = cv.take(npix * 8), o_depth = cv.take(npix * 4), o_img = cv.take(npix * c * 4),
               o_vert = cv.take((size_t)nver * 12), o_tri = cv.take((size_t)ntri * 12),
               o_ttri = cv.take((size_t)ntri * 12), o_tc = cv.take((size_t)tex_nver * 12),
               o_tex = cv.take((size_t)tex_h * tex_w * tex_c * 4);
  VP_CUDA(g_extra.reserve(cv.total, device));
  char* base = g_extra.as<char>();
  cudaStream_t st = nullptr;
  VP_CUDA(cudaMemcpyAsync(base + o_depth, depth_buffer, npix * 4, cudaMemcpyHostToDevice, st));
  VP_CUDA(cudaMemcpyAsync(base + o_img, image, npix * c * 4, cudaMemcpyHostToDevice, st));
  if (nver) VP_CUDA(cudaMemcpyAsync(base + o_vert, vertices, (size_t)nver * 12, cudaMemcpyHostToDevice, st));
  if (ntri) {
    VP_CUDA(cudaMemcpyAsync(base + o_tri, triangles, (size_t)ntri * 12, cudaMemcpyHostToDevice, st));
    VP_CUDA(cudaMemcpyAsync(base + o_ttri, tex_triangles, (size_t)ntri * 12, cudaMemcpyHostToDevice, st));
    VP_CUDA(cudaMemcpyAsync(base + o_tc, tex_coords, (size_t)tex_nver * 12, cudaMemcpyHostToDevice, st));
    VP_CUDA(cudaMemcpyAsync(base + o_tex, texture, (size_t)tex_h * tex_w * tex_c * 4, cudaMemcpyHostToDevice, st));
  }
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(base + o_keys);
  float* depth = reinterpret_cast<float*>(base + o_depth);
  VP_TRY(launch_keys_from_depth(depth, keys, npix, st));
  VP_TRY(launch_scatter_generic(kModeTriangles, reinterpret_cast<const float*>(base + o_vert), (size_t)3 * nver,
                                reinterpret_cast<const int*>(base + o_tri), keys, 1, ntri, h, w, st));
  if (ntri) {
    GenericMesh mesh{reinterpret_cast<const float*>(base + o_vert), reinterpret_cast<const int*>(base + o_tri),
                     (size_t)3 * nver};
    TextureArgs ta{reinterpret_cast<const float*>(base + o_tex), reinterpret_cast<const float*>(base + o_tc),
                   reinterpret_cast<const int*>(base + o_ttri), c, tex_h, tex_w, tex_c, mapping_type};
    resolve_texture_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(keys, mesh, ta,
                                                                          reinterpret_cast<float*>(base + o_img), depth, h, w);
    VP_LAUNCH_CHECK();
  }
  VP_CUDA(cudaMemcpyAsync(image, base + o_img, npix * c * 4, cudaMemcpyDeviceToHost, st));
  VP_CUDA(cudaMemcpyAsync(depth_buffer, base + o_depth, npix * 4, cudaMemcpyDeviceToHost, st));
  VP_CUDA(cudaStreamSynchronize(st));
  return VP_OK;
}

extern "C" int vp_get_normal_core(float* normal, const float* tri_normal, const int* triangles, int nver, int ntri) {
  VP_REQUIRE(nver >= 0 && ntri >= 0, "negative element count");
  VP_REQUIRE(nver == 0 || normal, "null normal buffer");
  VP_REQUIRE(ntri == 0 || (tri_normal && triangles), "null triangle buffer");
  if (ntri == 0 || nver == 0) {
    VP_REQUIRE(ntri == 0, "triangles but no vertices");
    return VP_OK;
  }
  VP_REQUIRE((size_t)ntri * 3 < (size_t)0x7FFFFFFF, "too many triangles");
  for (size_t i = 0; i < 3 * (size_t)ntri; ++i)
    VP_REQUIRE(triangles[i] >= 0 && triangles[i] < nver, "triangle index out of range");
  int device = 0;
  VP_CUDA(cudaGetDevice(&device));
  std::lock_guard<std::mutex> lock(g_extra_mutex);
  const int ninc = 3 * ntri;
  cudaStream_t st = nullptr;
  int end_bit = 1;
  while (end_bit < 31 && (1ll << end_bit) < (long long)nver) ++end_bit;
  size_t temp_bytes = 0;
  VP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const int*)nullptr, (int*)nullptr, (const int*)nullptr,
                                          (int*)nullptr, ninc, 0, end_bit, st));
  Carver cv;
  const size_t o_keys_in = cv.take((size_t)ninc * 4), o_keys_out = cv.take((size_t)ninc * 4),
               o_val_in = cv.take((size_t)ninc * 4), o_val_out = cv.take((size_t)ninc * 4),
               o_tn = cv.take((size_t)ntri * 12), o_n = cv.take((size_t)nver * 12), o_tmp = cv.take(temp_bytes);
  VP_CUDA(g_extra.reserve(cv.total, device));
  char* base = g_extra.as<char>();
  VP_CUDA(cudaMemcpyAsync(base + o_keys_in, triangles, (size_t)ninc * 4, cudaMemcpyHostToDevice, st));
  VP_CUDA(cudaMemcpyAsync(base + o_tn, tri_normal, (size_t)ntri * 12, cudaMemcpyHostToDevice, st));
  VP_CUDA(cudaMemcpyAsync(base + o_n, normal, (size_t)nver * 12, cudaMemcpyHostToDevice, st));
  iota_kernel<<<(ninc + 255) / 256, 256, 0, st>>>(reinterpret_cast<int*>(base + o_val_in), ninc);
  VP_LAUNCH_CHECK();
  VP_CUDA(cub::DeviceRadixSort::SortPairs(base + o_tmp, temp_bytes, reinterpret_cast<const int*>(base + o_keys_in),
                                          reinterpret_cast<int*>(base + o_keys_out),
                                          reinterpret_cast<const int*>(base + o_val_in),
                                          reinterpret_cast<int*>(base + o_val_out), ninc, 0, end_bit, st));
  count_launch(2);
  ordered_normal_sum_kernel<<<(nver + 127) / 128, 128, 0, st>>>(
      reinterpret_cast<const int*>(base + o_keys_out), reinterpret_cast<const int*>(base + o_val_out),
      reinterpret_cast<const float*>(base + o_tn), reinterpret_cast<float*>(base + o_n), nver, ninc);
  VP_LAUNCH_CHECK();
  VP_CUDA(cudaMemcpyAsync(normal, base + o_n, (size_t)nver * 12, cudaMemcpyDeviceToHost, st));
  VP_CUDA(cudaStreamSynchronize(st));
  return VP_OK;
}
