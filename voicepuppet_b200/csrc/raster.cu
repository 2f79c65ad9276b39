// C-ABI entry points that replace the reference's mesh_core_cython rasterizers
// (utils/cython/mesh_core_cython.pyx:49-78) and the launchers shared with the fused path.
#include "raster.cuh"

#include <cstdlib>
#include <mutex>

#include "common.h"
#include "launch.h"

namespace vp {

namespace {
std::mutex g_scratch_mutex;  // serialises the host-pointer entry points' scratch
DevBuf g_scratch;

struct Carver {  // carves 256-byte aligned sub-buffers out of one allocation
  size_t total = 0;
  size_t take(size_t bytes) {
    const size_t off = total;
    total += (bytes + 255) & ~size_t(255);
    return off;
  }
};
}  // namespace

int launch_keys_from_depth(const float* depth_dev, unsigned long long* keys_dev, size_t n, cudaStream_t st) {
  if (n == 0) return VP_OK;
  keys_from_depth_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(depth_dev, keys_dev, n);
  VP_LAUNCH_CHECK();
  return VP_OK;
}

int launch_scatter_generic(int mode, const float* vertices, size_t frame_stride, const int* triangles,
                           unsigned long long* keys, int nframes, int ntri, int h, int w, cudaStream_t st) {
  if (ntri == 0 || nframes == 0) return VP_OK;
  GenericMesh mesh{vertices, triangles, frame_stride};
  dim3 grid((ntri + kRasterBlock - 1) / kRasterBlock, nframes);
  if (mode == kModeColors)
    raster_scatter_kernel<kModeColors, GenericMesh, FullKey><<<grid, kRasterBlock, 0, st>>>(
        mesh, FullKey(), keys, nullptr, ntri, nframes, 1, h, w, 0);
  else
    raster_scatter_kernel<kModeTriangles, GenericMesh, FullKey><<<grid, kRasterBlock, 0, st>>>(
        mesh, FullKey(), keys, nullptr, ntri, nframes, 1, h, w, 0);
  VP_LAUNCH_CHECK();
  return VP_OK;
}

// Epoch keys need the inverted triangle index to fit next to a 32-bit depth code and >= 1 epoch bit.
int epoch_tri_bits(int ntri) {
  int bits = 1;
  while (bits < 31 && (1ll << bits) < (long long)ntri + 1) ++bits;
  return bits;
}
uint32_t epoch_limit(int ntri) { return (1u << (32 - epoch_tri_bits(ntri))) - 1u; }  // largest usable epoch

EpochKey make_epoch_key(int ntri, uint32_t epoch) {
  EpochKey km;
  km.tri_bits = epoch_tri_bits(ntri);
  km.tri_mask = (1u << km.tri_bits) - 1u;
  km.epoch_field = static_cast<unsigned long long>(epoch) << (32 + km.tri_bits);
  return km;
}

// Boxes up to this many pixels are walked by their own lane; larger ones are flattened over the warp (raster_walk.cuh).
int inline_box_pixels() {
  static const int v = [] { const char* e = std::getenv("VPB200_INLINE_BOX"); return e ? std::atoi(e) : 64; }();
  return v;
}

// Lanes per group of the scatter kernel's group walk (0 = every lane walks its own box), see raster_walk.cuh.
int walk_group_lanes() {
  static const int v = [] {
    const char* e = std::getenv("VPB200_WALK_GROUP");
    const int g = e ? std::atoi(e) : 4;
    return (g == 2 || g == 4 || g == 8 || g == 16 || g == 32) ? g : 0;
  }();
  return v;
}

int walk_group_min_pixels() {
  static const int v = [] { const char* e = std::getenv("VPB200_WALK_GROUP_MIN"); return e ? std::atoi(e) : 320; }();
  return v;
}

int launch_scatter_packed(const float4* vrec, size_t frame_stride, const int4* triangles,
                          unsigned long long* keys, uint32_t* tri_color, uint32_t epoch, int nframes, int ntri, int h,
                          int w, cudaStream_t st) {
  if (ntri == 0 || nframes == 0) return VP_OK;
  // From 768x768 the group walk (raster_walk.cuh; 1024x1024: 4.02 -> 3.52 us per frame) in a 64-register build, below it
  // every lane walks its own box in a 48-register build (40 warps/SM: +3..9 % there; the mere presence of the group path
  // costs 2 % at 256x256, hence two instantiations).  CTAs of 128 threads retire sooner than the 256 of round 1 (+1..3 %),
  // and the triangle indices are loaded once for 8 frames (4 at large frames, where 8 loses 1.5 %): profiles/r02o_group_walk.txt.
  static const long long group_res = [] { const char* e = std::getenv("VPB200_WALK_GROUP_RES"); return e ? std::atoll(e) : 768ll; }();
  const bool large = (long long)h * w >= group_res * group_res;
  static const int fpb_env = [] { const char* e = std::getenv("VPB200_SCATTER_FPB"); return e ? std::atoi(e) : 0; }();
  const int fpb = fpb_env > 0 ? fpb_env : (nframes >= 32 ? (large ? 4 : 8) : (nframes >= 16 ? 2 : 1));
  const bool fits32 = (unsigned long long)nframes * frame_stride < (1ull << 32) &&
                      (unsigned long long)nframes * (unsigned long long)ntri < (1ull << 32);
  if (!fits32) {  // the generic template on the packed records (64-bit offsets)
    dim3 grid((ntri + kRasterBlock - 1) / kRasterBlock, (nframes + fpb - 1) / fpb);
    PackedMesh mesh{vrec, triangles, frame_stride};
    raster_scatter_kernel<kModeColors, PackedMesh, EpochKey><<<grid, kRasterBlock, 0, st>>>(
        mesh, make_epoch_key(ntri, epoch), keys, tri_color, ntri, nframes, fpb, h, w, 1);
  } else {
    ScatterArgs a;
    a.vrec = vrec;
    a.tris = triangles;
    a.keys = keys;
    a.tri_color = tri_color;
    a.km = make_epoch_key(ntri, epoch);
    a.stride = (unsigned)frame_stride;
    a.ntri = ntri;
    a.nframes = nframes;
    a.frames_per_block = fpb;
    a.h = h;
    a.w = w;
    a.inline_max = inline_box_pixels();
    a.group = walk_group_lanes();
    a.group_min = walk_group_min_pixels();
    dim3 grid((ntri + kScatterBlock - 1) / kScatterBlock, (nframes + fpb - 1) / fpb);
    if (a.group > 0 && large)
      raster_scatter_packed_kernel<8, true><<<grid, kScatterBlock, 0, st>>>(a);
    else
      raster_scatter_packed_kernel<10, false><<<grid, kScatterBlock, 0, st>>>(a);
  }
  VP_LAUNCH_CHECK();
  return VP_OK;
}

int launch_resolve_packed(const unsigned long long* keys, const uint32_t* tri_color, uint32_t epoch, unsigned char* image, unsigned char* mask, int nframes, int ntri, int h, int w,
                          cudaStream_t st) {
  const size_t npix = (size_t)h * w;
  if (nframes == 0 || npix == 0) return VP_OK;
  // pixels per thread: 8 (two groups of 4, all key loads in flight before the first dependent gather) from 512x512 and in
  // long launches -- resolve -9 % at 1024x1024 (0.80 -> 0.88 of the HBM peak), -8 % at 512x512, step +1 % at 3000 x 256x256 --
  // 4 in the short launches of a 75-frame clip; 16 loses everywhere (profiles/r02o_group_walk.txt, calls 32-33)
  static const int px_env = [] { const char* e = std::getenv("VPB200_RESOLVE_PX"); return e ? std::atoi(e) : 0; }();
  const int px = px_env ? px_env : ((npix >= 512u * 512u || nframes >= 64) ? 8 : 4);
  const EpochKey km = make_epoch_key(ntri, epoch);
  if (px >= 8) {
    dim3 grid((unsigned)((npix / 4 + 511) / 512), nframes);
    resolve_packed_kernel<2><<<grid, 256, 0, st>>>(keys, km, tri_color, image, mask, ntri, npix);
  } else {
    dim3 grid((unsigned)((npix / 4 + 255) / 256), nframes);
    resolve_packed_kernel<1><<<grid, 256, 0, st>>>(keys, km, tri_color, image, mask, ntri, npix);
  }
  VP_LAUNCH_CHECK();
  return VP_OK;
}

// Host-pointer entry points: a triangle index outside the vertex buffer would make the kernels read out of bounds
// (the reference's C++ has no check either and reads whatever is there: mesh_core.cpp:186-188).
static int check_triangle_indices(const int* triangles, int ntri, int nver) {
  for (size_t i = 0; i < (size_t)ntri * 3; ++i)
    VP_REQUIRE(triangles[i] >= 0 && triangles[i] < nver, "triangle index outside the vertex buffer");
  return VP_OK;
}

static int check_raster_args(int nver, int ntri, int h, int w) {
  VP_REQUIRE(nver >= 0 && ntri >= 0, "negative element count");
  VP_REQUIRE(h > 0 && w > 0 && h <= 16384 && w <= 16384, "image size must be in 1..16384");
  return VP_OK;
}

int render_colors_batch_dev(unsigned char* image, unsigned char* mask, const float* vertices, const int* triangles,
                            const float* colors, float* depth, int* triangle_id, unsigned long long* keys,
                            int nframes, int nver, int ntri, int h, int w, int c, cudaStream_t st) {
  const size_t npix = (size_t)h * w;
  VP_TRY(launch_keys_from_depth(depth, keys, npix * nframes, st));
  VP_TRY(launch_scatter_generic(kModeColors, vertices, (size_t)3 * nver, triangles, keys, nframes, ntri, h, w, st));
  GenericMesh mesh{vertices, triangles, (size_t)3 * nver};
  dim3 grid((unsigned)((npix + 255) / 256), nframes);
  resolve_colors_generic_kernel<<<grid, 256, 0, st>>>(keys, mesh, colors, (size_t)c * nver, image, mask, depth,
                                                      triangle_id, h, w, c);
  VP_LAUNCH_CHECK();
  return VP_OK;
}

}  // namespace vp

using namespace vp;

extern "C" int vp_render_colors_batch_dev(unsigned char* image, unsigned char* face_mask, const float* vertices,
                                          const int* triangles, const float* colors, float* depth_buffer,
                                          int* triangle_id, int nframes, int nver, int ntri, int h, int w, int c,
                                          int device, void* stream) {
  VP_TRY(check_raster_args(nver, ntri, h, w));
  VP_REQUIRE(nframes >= 0 && c >= 1, "nframes >= 0 and c >= 1");
  VP_REQUIRE(image && face_mask && depth_buffer, "null output buffer");
  VP_REQUIRE(ntri == 0 || (vertices && triangles && colors), "null mesh buffer");
  if (nframes == 0) return VP_OK;
  VP_CUDA(cudaSetDevice(device));
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  const size_t npix = (size_t)h * w;
  VP_CUDA(g_scratch.reserve(npix * nframes * sizeof(unsigned long long), device));
  return render_colors_batch_dev(image, face_mask, vertices, triangles, colors, depth_buffer, triangle_id,
                                 g_scratch.as<unsigned long long>(), nframes, nver, ntri, h, w, c,
                                 static_cast<cudaStream_t>(stream));
}

extern "C" int vp_render_colors_core(unsigned char* image, unsigned char* face_mask, const float* vertices,
                                     const int* triangles, const float* colors, float* depth_buffer,
                                     int* triangle_id, int nver, int ntri, int h, int w, int c) {
  VP_TRY(check_raster_args(nver, ntri, h, w));
  VP_REQUIRE(c >= 1, "c >= 1");
  VP_REQUIRE(image && face_mask && depth_buffer, "null output buffer");
  VP_REQUIRE(ntri == 0 || (vertices && triangles && colors), "null mesh buffer");
  VP_TRY(check_triangle_indices(triangles, ntri, nver));
  int device = 0;
  VP_CUDA(cudaGetDevice(&device));
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  const size_t npix = (size_t)h * w;
  Carver cv;
  const size_t o_keys = cv.take(npix * 8), o_img = cv.take(npix * c), o_mask = cv.take(npix),
               o_depth = cv.take(npix * 4), o_tid = cv.take(npix * 4), o_vert = cv.take((size_t)nver * 12),
               o_tri = cv.take((size_t)ntri * 12), o_col = cv.take((size_t)nver * c * 4);
  VP_CUDA(g_scratch.reserve(cv.total, device));
  char* base = g_scratch.as<char>();
  cudaStream_t st = nullptr;
  VP_CUDA(cudaMemcpyAsync(base + o_img, image, npix * c, cudaMemcpyHostToDevice, st));
  VP_CUDA(cudaMemcpyAsync(base + o_mask, face_mask, npix, cudaMemcpyHostToDevice, st));
  VP_CUDA(cudaMemcpyAsync(base + o_depth, depth_buffer, npix * 4, cudaMemcpyHostToDevice, st));
  if (nver) VP_CUDA(cudaMemcpyAsync(base + o_vert, vertices, (size_t)nver * 12, cudaMemcpyHostToDevice, st));
  if (ntri) VP_CUDA(cudaMemcpyAsync(base + o_tri, triangles, (size_t)ntri * 12, cudaMemcpyHostToDevice, st));
  if (nver) VP_CUDA(cudaMemcpyAsync(base + o_col, colors, (size_t)nver * c * 4, cudaMemcpyHostToDevice, st));
  VP_TRY(render_colors_batch_dev(reinterpret_cast<unsigned char*>(base + o_img),
                                 reinterpret_cast<unsigned char*>(base + o_mask),
                                 reinterpret_cast<const float*>(base + o_vert),
                                 reinterpret_cast<const int*>(base + o_tri),
                                 reinterpret_cast<const float*>(base + o_col),
                                 reinterpret_cast<float*>(base + o_depth),
                                 triangle_id ? reinterpret_cast<int*>(base + o_tid) : nullptr,
                                 reinterpret_cast<unsigned long long*>(base + o_keys), 1, nver, ntri, h, w, c, st));
  VP_CUDA(cudaMemcpyAsync(image, base + o_img, npix * c, cudaMemcpyDeviceToHost, st));
  VP_CUDA(cudaMemcpyAsync(face_mask, base + o_mask, npix, cudaMemcpyDeviceToHost, st));
  VP_CUDA(cudaMemcpyAsync(depth_buffer, base + o_depth, npix * 4, cudaMemcpyDeviceToHost, st));
  if (triangle_id) VP_CUDA(cudaMemcpyAsync(triangle_id, base + o_tid, npix * 4, cudaMemcpyDeviceToHost, st));
  VP_CUDA(cudaStreamSynchronize(st));
  return VP_OK;
}

extern "C" int vp_rasterize_triangles_core(const float* vertices, const int* triangles, float* depth_buffer,
                                           int* triangle_buffer, float* barycentric_weight, int nver, int ntri,
                                           int h, int w) {
  VP_TRY(check_raster_args(nver, ntri, h, w));
  VP_REQUIRE(depth_buffer && triangle_buffer && barycentric_weight, "null output buffer");
  VP_REQUIRE(ntri == 0 || (vertices && triangles), "null mesh buffer");
  VP_TRY(check_triangle_indices(triangles, ntri, nver));
  int device = 0;
  VP_CUDA(cudaGetDevice(&device));
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  const size_t npix = (size_t)h * w;
  Carver cv;
  const size_t o_keys = cv.take(npix * 8), o_depth = cv.take(npix * 4), o_tid = cv.take(npix * 4),
               o_wgt = cv.take(npix * 12), o_vert = cv.take((size_t)nver * 12), o_tri = cv.take((size_t)ntri * 12);
  VP_CUDA(g_scratch.reserve(cv.total, device));
  char* base = g_scratch.as<char>();
  cudaStream_t st = nullptr;
  VP_CUDA(cudaMemcpyAsync(base + o_depth, depth_buffer, npix * 4, cudaMemcpyHostToDevice, st));
  VP_CUDA(cudaMemcpyAsync(base + o_tid, triangle_buffer, npix * 4, cudaMemcpyHostToDevice, st));
  VP_CUDA(cudaMemcpyAsync(base + o_wgt, barycentric_weight, npix * 12, cudaMemcpyHostToDevice, st));
  if (nver) VP_CUDA(cudaMemcpyAsync(base + o_vert, vertices, (size_t)nver * 12, cudaMemcpyHostToDevice, st));
  if (ntri) VP_CUDA(cudaMemcpyAsync(base + o_tri, triangles, (size_t)ntri * 12, cudaMemcpyHostToDevice, st));
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(base + o_keys);
  float* depth = reinterpret_cast<float*>(base + o_depth);
  VP_TRY(launch_keys_from_depth(depth, keys, npix, st));
  VP_TRY(launch_scatter_generic(kModeTriangles, reinterpret_cast<const float*>(base + o_vert), (size_t)3 * nver,
                                reinterpret_cast<const int*>(base + o_tri), keys, 1, ntri, h, w, st));
  GenericMesh mesh{reinterpret_cast<const float*>(base + o_vert), reinterpret_cast<const int*>(base + o_tri),
                   (size_t)3 * nver};
  resolve_triangles_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(
      keys, mesh, depth, reinterpret_cast<int*>(base + o_tid), reinterpret_cast<float*>(base + o_wgt), h, w);
  VP_LAUNCH_CHECK();
  VP_CUDA(cudaMemcpyAsync(depth_buffer, base + o_depth, npix * 4, cudaMemcpyDeviceToHost, st));
  VP_CUDA(cudaMemcpyAsync(triangle_buffer, base + o_tid, npix * 4, cudaMemcpyDeviceToHost, st));
  VP_CUDA(cudaMemcpyAsync(barycentric_weight, base + o_wgt, npix * 12, cudaMemcpyDeviceToHost, st));
  VP_CUDA(cudaStreamSynchronize(st));
  return VP_OK;
}

#include "mesh_extra.cuh"
