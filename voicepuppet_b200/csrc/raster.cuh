// Z-buffer rasterizer kernels (K3 scatter + K4 resolve), order independent.
//
// The reference walks triangles sequentially and keeps a pixel when "d > depth[p]"
// (utils/cython/mesh_core.cpp:153,211).  That fixed point is: the largest candidate depth
// strictly above the caller's initial depth, ties to the lowest triangle index.  We get it
// with one 64-bit atomicMax per (triangle, pixel) candidate on key = (depth code, ~index)
// (vp_math.cuh), then a resolve pass that recomputes the winner's outputs with the same
// float32 expressions -- so image/mask/depth/triangle/weights are bit-identical.
#pragma once

#include <cuda_runtime.h>

#include "raster_keys.cuh"
#include "raster_walk.cuh"
#include "vp_math.cuh"

namespace vp {

constexpr int kRasterBlock = 256;
constexpr int kScatterBlock = 128;  // CTA of the packed scatter kernel
constexpr int kSmallBox = 12;  // boxes up to this many pixels are walked by the owning lane

enum RasterMode { kModeColors = 0, kModeTriangles = 1 };

// ---- vertex / triangle fetch policies -------------------------------------------------
// Generic: the reference's flat arrays (float xyz per vertex, int32 index triples).
struct GenericMesh {
  const float* vertices;   // [frames][3*nver]
  const int* triangles;    // [3*ntri]
  size_t frame_stride;     // floats
  __device__ __forceinline__ void indices(int f, int& a, int& b, int& c, uint32_t& id) const {
    a = __ldg(triangles + 3 * (size_t)f);
    b = __ldg(triangles + 3 * (size_t)f + 1);
    c = __ldg(triangles + 3 * (size_t)f + 2);
    id = (uint32_t)f;
  }
  __device__ __forceinline__ void vertex(int frame, int i, float& x, float& y, float& z, uint32_t& rgba) const {
    const float* p = vertices + (size_t)frame * frame_stride + 3 * (size_t)i;
    x = __ldg(p);
    y = __ldg(p + 1);
    z = __ldg(p + 2);
    rgba = 0;
  }
};

// Packed: the fused pipeline's records.  Vertex = float4 (x, y, z, rgba bits), triangle =
// int4 (internal vertex ids, original triangle index for the tie-break).
struct PackedMesh {
  const float4* vertices;  // [frames][nver_pad]
  const int4* triangles;   // [ntri]
  size_t frame_stride;     // float4 elements
  __device__ __forceinline__ void indices(int f, int& a, int& b, int& c, uint32_t& id) const {
    const int4 t = __ldg(triangles + f);
    a = t.x;
    b = t.y;
    c = t.z;
    id = (uint32_t)t.w;
  }
  __device__ __forceinline__ void vertex(int frame, int i, float& x, float& y, float& z, uint32_t& rgba) const {
    const float4 v = __ldg(vertices + (size_t)frame * frame_stride + i);
    x = v.x;
    y = v.y;
    z = v.z;
    rgba = __float_as_uint(v.w);
  }
};

struct Candidate {  // what one lane knows about its triangle
  TriSetup s;
  float z0, z1, z2;             // kModeTriangles: per-corner depth
  unsigned long long key;       // kModeColors: the triangle's key (flat depth, index)
  uint32_t id;
  int n;                        // bbox pixel count (0 = nothing to do)
};

// One pixel row of the bounding box.  The row terms e0y*py / e1y*py are hoisted: they are the
// same individually rounded products pixel_uv() forms (mesh_core.cpp:29,34,36).
template <int MODE, typename KeyMaker>
__device__ __forceinline__ void offer_span(const Candidate& c, const KeyMaker& km, int y, int x_begin, int x_end,
                                           int x_step, unsigned long long* keys, int h, int w) {
  const float py = VP_SUB(static_cast<float>(y), c.s.ay);
  const float m0y = VP_MUL(c.s.e0y, py), m1y = VP_MUL(c.s.e1y, py);
  unsigned long long* row = keys + (size_t)y * w;
  for (int x = x_begin; x <= x_end; x += x_step) {
    const float px = VP_SUB(static_cast<float>(x), c.s.ax);
    const float d02 = VP_ADD(VP_MUL(c.s.e0x, px), m0y);
    const float d12 = VP_ADD(VP_MUL(c.s.e1x, px), m1y);
    const float u = VP_MUL(VP_SUB(VP_MUL(c.s.d11, d02), VP_MUL(c.s.d01, d12)), c.s.inv);
    const float v = VP_MUL(VP_SUB(VP_MUL(c.s.d00, d12), VP_MUL(c.s.d01, d02)), c.s.inv);
    if (MODE == kModeColors) {
      if (uv_inside(u, v)) atomicMax(row + x, c.key);
    } else {
      if (in_border(x, y, h, w) || uv_inside(u, v)) {
        float w0, w1, w2;
        const float d = weights_depth(u, v, c.z0, c.z1, c.z2, w0, w1, w2);
        if (d == d) atomicMax(row + x, km.make(d, c.id));
      }
    }
  }
}

// One lane per triangle.  Boxes of up to kSmallBox pixels are walked by the owning lane; larger
// ones are broadcast through shared memory and walked by the whole warp, one row per iteration.
// tri_color (may be NULL, kModeColors + PackedMesh only): flat colour per triangle, written here so
// that the resolve pass needs one colour gather per covered pixel.
// const_init: the initial depth is the constant kInitDepth (infer_bfmvid.py:106), so the strict
// "d > initial" test (mesh_core.cpp:211) is done here per triangle instead of through init keys.
template <int MODE, typename Mesh, typename KeyMaker>
__global__ void __launch_bounds__(kRasterBlock)
raster_scatter_kernel(Mesh mesh, KeyMaker km, unsigned long long* __restrict__ keys_all,
                      uint32_t* __restrict__ tri_color, int ntri, int nframes, int frames_per_block, int h, int w,
                      int const_init) {
  __shared__ Candidate s_big[kRasterBlock / 32];
  const int f = blockIdx.x * kRasterBlock + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  int ia = 0, ib = 0, ic = 0;
  uint32_t id = 0;
  if (f < ntri) mesh.indices(f, ia, ib, ic, id);  // shared by all the frames this block walks

  const int frame_begin = blockIdx.y * frames_per_block;
  const int frame_end = min(nframes, frame_begin + frames_per_block);
  for (int frame = frame_begin; frame < frame_end; ++frame) {
    unsigned long long* keys = keys_all + (size_t)frame * h * w;
    Candidate c;
    c.n = 0;
    c.id = id;
    c.key = 0ull;
    c.z0 = c.z1 = c.z2 = 0.f;
    if (f < ntri) {
      float x0, y0, z0, x1, y1, z1, x2, y2, z2;
      uint32_t r0, r1, r2;
      mesh.vertex(frame, ia, x0, y0, z0, r0);
      mesh.vertex(frame, ib, x1, y1, z1, r1);
      mesh.vertex(frame, ic, x2, y2, z2, r2);
      if (tri_bbox(c.s, x0, y0, x1, y1, x2, y2, h, w)) {
        c.n = (c.s.x_hi - c.s.x_lo + 1) * (c.s.y_hi - c.s.y_lo + 1);
        if (MODE == kModeColors) {
          const float d = flat_depth(z0, z1, z2);
          if (!(d == d)) c.n = 0;                          // NaN never wins a '>' test
          if (const_init && !(d > kInitDepth)) c.n = 0;    // mesh_core.cpp:211 against -99999
          c.key = km.make(d, c.id);
        } else {
          c.z0 = z0;
          c.z1 = z1;
          c.z2 = z2;
        }
        if (c.n > 0) tri_edges(c.s, x0, y0, x1, y1, x2, y2);
      }
      if (MODE == kModeColors && tri_color != nullptr) {
        // integral colours in [0,255] packed by the vertex kernel: sum <= 765 is exact in float, so
        // integer arithmetic equals mesh_core.cpp:219.  Indexed by the triangle's id (the ORIGINAL index the
        // z-buffer key carries), so that the resolve pass needs one gather.
        const uint32_t r = ((r0 & 255u) + (r1 & 255u) + (r2 & 255u)) / 3u;
        const uint32_t g = (((r0 >> 8) & 255u) + ((r1 >> 8) & 255u) + ((r2 >> 8) & 255u)) / 3u;
        const uint32_t b = (((r0 >> 16) & 255u) + ((r1 >> 16) & 255u) + ((r2 >> 16) & 255u)) / 3u;
        tri_color[(size_t)frame * ntri + id] = r | (g << 8) | (b << 16) | 0xFF000000u;
      }
    }

    if (c.n > 0 && c.n <= kSmallBox) {
      for (int y = c.s.y_lo; y <= c.s.y_hi; ++y) offer_span<MODE>(c, km, y, c.s.x_lo, c.s.x_hi, 1, keys, h, w);
    }
    unsigned big = __ballot_sync(0xFFFFFFFFu, c.n > kSmallBox);
    Candidate* slot = &s_big[threadIdx.x >> 5];
    while (big) {
      const int src = __ffs(big) - 1;
      big &= big - 1;
      __syncwarp();
      if ((int)lane == src) *slot = c;
      __syncwarp();
      const Candidate o = *slot;
      const int bw = o.s.x_hi - o.s.x_lo + 1;
      if (bw >= 16) {  // wide box: the warp strides along x, row by row
        for (int y = o.s.y_lo; y <= o.s.y_hi; ++y)
          offer_span<MODE>(o, km, y, o.s.x_lo + (int)lane, o.s.x_hi, 32, keys, h, w);
      } else {         // narrow box: 32 / bw' rows at a time (bw' = bw rounded up to a power of two)
        const int bwp = bw <= 1 ? 1 : (bw <= 2 ? 2 : (bw <= 4 ? 4 : (bw <= 8 ? 8 : 16)));
        const int rows_per_iter = 32 / bwp;
        const int dx = (int)lane & (bwp - 1), dy = (int)lane / bwp;
        for (int y = o.s.y_lo + dy; y <= o.s.y_hi; y += rows_per_iter)
          if (dx < bw) offer_span<MODE>(o, km, y, o.s.x_lo + dx, o.s.x_lo + dx, 1, keys, h, w);
      }
    }
  }
}

// ---- fused-path scatter (render_colors semantics, packed records, epoch keys) ----------------
// Same arithmetic as raster_scatter_kernel<kModeColors, PackedMesh, EpochKey> with const_init, written
// for the instruction-issue limit that bounds it (ncu r01d: 75 % issue-active, 365 warp-instructions per
// 32 triangles): uniform per-frame base pointers, a bounding box without the x86-cast emulation when
// every coordinate is finite and small (the emulation only matters for NaN / |x| >= 2^31), byte-lane
// arithmetic for the flat colour, and one flat pixel loop per lane.
struct ScatterArgs {
  const float4* vrec;           // [frames][stride]
  const int4* tris;             // [ntri]
  unsigned long long* keys;     // [frames][h*w]
  uint32_t* tri_color;          // [frames][ntri]
  EpochKey km;
  unsigned stride;              // float4 per frame
  int ntri, nframes, frames_per_block, h, w;
  int inline_max;               // boxes up to this many pixels are walked by their own lane, larger ones flattened
  int group, group_min;         // lanes (0 / 4 / 8) that walk each other's boxes together, and the pixels a warp's boxes must hold for it (raster_walk.cuh)
};

template <int MIN_BLOCKS, bool GROUP>
__global__ void __launch_bounds__(kScatterBlock, MIN_BLOCKS)
raster_scatter_packed_kernel(const ScatterArgs a) {
  __shared__ float4 s_rec[kScatterBlock / 32][4][32];  // per-warp staging of the group walk / the flattened large boxes (raster_walk.cuh)
  const int f = blockIdx.x * kScatterBlock + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  const bool valid = f < a.ntri;
  int4 t = make_int4(0, 0, 0, 0);
  if (valid) t = __ldg(a.tris + f);  // shared by all the frames this block walks
  const unsigned long long low = static_cast<unsigned long long>(a.km.tri_mask - (uint32_t)t.w);
  const int wm1 = a.w - 1, hm1 = a.h - 1;
  const unsigned npix = (unsigned)a.h * (unsigned)a.w;

  const int frame_begin = blockIdx.y * a.frames_per_block;
  const int frame_end = min(a.nframes, frame_begin + a.frames_per_block);
  for (int frame = frame_begin; frame < frame_end; ++frame) {
    // 32-bit element offsets (the launcher checks nframes * stride and nframes * ntri fit)
    const unsigned vbase = (unsigned)frame * a.stride;
    unsigned long long* keys = a.keys + (size_t)(unsigned)frame * npix;
    Candidate c;
    c.n = 0;
    c.key = 0ull;
    c.id = 0;
    c.z0 = c.z1 = c.z2 = 0.f;
    if (valid) {
      const float4 v0 = __ldg(a.vrec + (vbase + (unsigned)t.x)), v1 = __ldg(a.vrec + (vbase + (unsigned)t.y)),
                   v2 = __ldg(a.vrec + (vbase + (unsigned)t.z));
      // indexed by the ORIGINAL triangle index (what the z-buffer key carries): a scattered 4-byte store here, and the
      // resolve pass needs ONE gather per pixel instead of two dependent ones (measured: resolve -9..-12 %)
      a.tri_color[(unsigned)frame * (unsigned)a.ntri + (unsigned)t.w] =
          flat_color_packed(__float_as_uint(v0.w), __float_as_uint(v1.w), __float_as_uint(v2.w));
      const float kBig = 1073741824.0f;  // 2^30: below it ceil/floor and the int casts are exact and in range
      const bool tame = fabsf(v0.x) < kBig && fabsf(v1.x) < kBig && fabsf(v2.x) < kBig && fabsf(v0.y) < kBig &&
                        fabsf(v1.y) < kBig && fabsf(v2.y) < kBig;  // false for NaN / inf
      bool nonempty;
      if (tame) {
        // mesh_core.cpp:194-203 for finite coordinates: min / max are order independent, the casts exact
        c.s.x_lo = max(__float2int_ru(fminf(v0.x, fminf(v1.x, v2.x))), 0);
        c.s.x_hi = min(__float2int_rd(fmaxf(v0.x, fmaxf(v1.x, v2.x))), wm1);
        c.s.y_lo = max(__float2int_ru(fminf(v0.y, fminf(v1.y, v2.y))), 0);
        c.s.y_hi = min(__float2int_rd(fmaxf(v0.y, fmaxf(v1.y, v2.y))), hm1);
        nonempty = c.s.x_hi >= c.s.x_lo && c.s.y_hi >= c.s.y_lo;
      } else {
        nonempty = tri_bbox(c.s, v0.x, v0.y, v1.x, v1.y, v2.x, v2.y, a.h, a.w);
      }
      if (nonempty) {
        const float d = flat_depth(v0.z, v1.z, v2.z);
        if (d > kInitDepth) {  // mesh_core.cpp:211 against the constant initial depth; false for NaN
          c.n = (c.s.x_hi - c.s.x_lo + 1) * (c.s.y_hi - c.s.y_lo + 1);
          const uint32_t b = __float_as_uint(__fadd_rn(d, 0.0f));  // -0 -> +0
          const uint32_t code = b ^ (static_cast<uint32_t>(static_cast<int>(b) >> 31) | 0x80000000u);
          c.key = a.km.epoch_field | (static_cast<unsigned long long>(code) << a.km.tri_bits) | low;
          tri_edges(c.s, v0.x, v0.y, v1.x, v1.y, v2.x, v2.y);
        }
      }
    }

    walk_boxes(c.s, c.key, c.n, a.inline_max, GROUP ? a.group : 0, a.group_min, s_rec[threadIdx.x >> 5], keys, a.w, lane);
  }
}

// keys[p] = init_key(depth[p]) for the caller-initialised depth buffer.
__global__ void keys_from_depth_kernel(const float* __restrict__ depth, unsigned long long* __restrict__ keys,
                                       size_t n) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) keys[p] = init_key(depth[p]);
}

// Generic resolve for render_colors_core: any channel count, float colours, in-place buffers.
__global__ void resolve_colors_generic_kernel(const unsigned long long* __restrict__ keys, GenericMesh mesh,
                                              const float* __restrict__ colors, size_t color_stride,
                                              unsigned char* __restrict__ image, unsigned char* __restrict__ mask,
                                              float* __restrict__ depth, int* __restrict__ tri_out, int h, int w,
                                              int c) {
  const int frame = blockIdx.y;
  const size_t npix = (size_t)h * w;
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const size_t gp = (size_t)frame * npix + p;
  const int t = key_triangle(keys[gp]);
  if (tri_out) tri_out[gp] = t;
  if (t < 0) return;
  int ia, ib, ic;
  uint32_t id;
  mesh.indices(t, ia, ib, ic, id);
  const float* vb = mesh.vertices + (size_t)frame * mesh.frame_stride;
  depth[gp] = flat_depth(vb[3 * (size_t)ia + 2], vb[3 * (size_t)ib + 2], vb[3 * (size_t)ic + 2]);
  mask[gp] = 255;
  const float* cb = colors + (size_t)frame * color_stride;
  for (int k = 0; k < c; ++k)
    image[gp * c + k] = flat_color(cb[(size_t)c * ia + k], cb[(size_t)c * ib + k], cb[(size_t)c * ic + k]);
}

// Resolve for rasterize_triangles_core: recompute the winner's weights and depth.
__global__ void resolve_triangles_kernel(const unsigned long long* __restrict__ keys, GenericMesh mesh,
                                         float* __restrict__ depth, int* __restrict__ tri_buf,
                                         float* __restrict__ weights, int h, int w) {
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
  tri_buf[p] = t;
  weights[3 * p + 0] = w0;
  weights[3 * p + 1] = w1;
  weights[3 * p + 2] = w2;
}

// Fused-path resolve: NG (1 or 2) groups of 4 consecutive pixels per thread (the groups of a thread lie 256 groups apart, so
// that every load and store instruction of a warp stays contiguous), one tri_color gather per covered pixel (by original
// triangle index), every pixel written (uncovered or stale-epoch key -> 0), so neither the image nor the z-buffer needs a
// clear.  Requires (h*w) % 4 == 0.
template <int NG>
__global__ void __launch_bounds__(256)
resolve_packed_kernel(const unsigned long long* __restrict__ keys, EpochKey km,
                      const uint32_t* __restrict__ tri_color, unsigned char* __restrict__ image,
                      unsigned char* __restrict__ mask, int ntri, size_t npix) {
  const int frame = blockIdx.y;
  const size_t q0 = (size_t)blockIdx.x * (256 * NG) + threadIdx.x;  // first group of 4 pixels
  const uint32_t* tc = tri_color + (size_t)frame * ntri;
  unsigned long long k[NG][4];
#pragma unroll
  for (int g = 0; g < NG; ++g) {  // all key loads in flight before the first dependent gather
    const size_t q = q0 + (size_t)g * 256;
    if (q * 4 < npix) {
      const size_t base = (size_t)frame * npix + q * 4;
      const ulonglong2 k01 = __ldcs(reinterpret_cast<const ulonglong2*>(keys + base));
      const ulonglong2 k23 = __ldcs(reinterpret_cast<const ulonglong2*>(keys + base + 2));
      k[g][0] = k01.x;
      k[g][1] = k01.y;
      k[g][2] = k23.x;
      k[g][3] = k23.y;
    }
  }
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const size_t q = q0 + (size_t)g * 256;
    if (q * 4 >= npix) break;
    const size_t base = (size_t)frame * npix + q * 4;
    uint32_t col[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool live = k[g][i] >= km.epoch_field;  // written during this chunk
      const uint32_t t = km.tri_mask - (static_cast<uint32_t>(k[g][i]) & km.tri_mask);  // winner's ORIGINAL index
      // neighbouring pixels of one triangle carry the same key (flat depth): one gather chain serves them all
      // (at 1024x1024 a triangle covers ~8 pixels, so this removes about half of the dependent gathers)
      if (i > 0 && k[g][i] == k[g][i - 1])
        col[i] = col[i - 1];
      else
        col[i] = live ? __ldg(tc + t) : 0u;
    }
    // 12 bytes of RGB for 4 pixels as three 32-bit words
    const uint32_t w0 = (col[0] & 0xFFFFFFu) | ((col[1] & 0xFFu) << 24);
    const uint32_t w1 = ((col[1] >> 8) & 0xFFFFu) | ((col[2] & 0xFFFFu) << 16);
    const uint32_t w2 = ((col[2] >> 16) & 0xFFu) | ((col[3] & 0xFFFFFFu) << 8);
    uint32_t* out = reinterpret_cast<uint32_t*>(image + base * 3);
    __stcs(out, w0);
    __stcs(out + 1, w1);
    __stcs(out + 2, w2);
    if (mask != nullptr) {
      const uint32_t m = (col[0] >> 24) | ((col[1] >> 24) << 8) | ((col[2] >> 24) << 16) | ((col[3] >> 24) << 24);
      __stcs(reinterpret_cast<uint32_t*>(mask + base), m);
    }
  }
}

}  // namespace vp
