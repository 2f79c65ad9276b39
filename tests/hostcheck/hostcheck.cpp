// TEST INFRASTRUCTURE -- runs the product's shared rasterizer arithmetic (voicepuppet_b200/csrc/vp_math.cuh, the
// very header the CUDA kernels include) on the host, so that its float32 expressions, z-buffer keys and bounding
// boxes can be compared with the oracle and the reference golden vectors without a GPU.  The structure mirrors the
// kernels: scatter = every (triangle, pixel) candidate offers a 64-bit key to a max; resolve = the winner's outputs
// are recomputed.  Build: g++ -O2 -ffp-contract=off (every operation individually rounded, as on the device
// where the header spells them with __f*_rn intrinsics).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../voicepuppet_b200/csrc/vp_math.cuh"

using namespace vp;

namespace {
struct Tri {
  float x0, y0, z0, x1, y1, z1, x2, y2, z2;
};
Tri fetch(const float* v, const int* t) {
  return Tri{v[3 * t[0]], v[3 * t[0] + 1], v[3 * t[0] + 2], v[3 * t[1]], v[3 * t[1] + 1], v[3 * t[1] + 2],
             v[3 * t[2]], v[3 * t[2] + 1], v[3 * t[2] + 2]};
}
}  // namespace

extern "C" int hc_render_colors(unsigned char* image, unsigned char* face_mask, const float* vertices,
                                const int* triangles, const float* colors, float* depth, int* tri_out, int ntri, int h,
                                int w, int c) {
  const size_t npix = (size_t)h * w;
  std::vector<unsigned long long> keys(npix);
  for (size_t p = 0; p < npix; ++p) keys[p] = init_key(depth[p]);
  for (int i = 0; i < ntri; ++i) {
    const Tri t = fetch(vertices, triangles + 3 * (size_t)i);
    TriSetup s;
    if (!tri_bbox(s, t.x0, t.y0, t.x1, t.y1, t.x2, t.y2, h, w)) continue;
    const float d = flat_depth(t.z0, t.z1, t.z2);
    if (!(d == d)) continue;
    tri_edges(s, t.x0, t.y0, t.x1, t.y1, t.x2, t.y2);
    const unsigned long long key = make_key(d, (uint32_t)i);
    for (int y = s.y_lo; y <= s.y_hi; ++y)
      for (int x = s.x_lo; x <= s.x_hi; ++x) {
        float u, v;
        pixel_uv(s, x, y, u, v);
        if (uv_inside(u, v) && key > keys[(size_t)y * w + x]) keys[(size_t)y * w + x] = key;
      }
  }
  for (size_t p = 0; p < npix; ++p) {
    const int i = key_triangle(keys[p]);
    if (tri_out) tri_out[p] = i;
    if (i < 0) continue;
    const int* t = triangles + 3 * (size_t)i;
    const Tri q = fetch(vertices, t);
    depth[p] = flat_depth(q.z0, q.z1, q.z2);
    face_mask[p] = 255;
    for (int k = 0; k < c; ++k)
      image[p * c + k] = flat_color(colors[(size_t)c * t[0] + k], colors[(size_t)c * t[1] + k], colors[(size_t)c * t[2] + k]);
  }
  return 0;
}

extern "C" int hc_rasterize_triangles(const float* vertices, const int* triangles, float* depth, int* tri_buf,
                                      float* weights, int ntri, int h, int w) {
  const size_t npix = (size_t)h * w;
  std::vector<unsigned long long> keys(npix);
  for (size_t p = 0; p < npix; ++p) keys[p] = init_key(depth[p]);
  for (int i = 0; i < ntri; ++i) {
    const Tri t = fetch(vertices, triangles + 3 * (size_t)i);
    TriSetup s;
    if (!tri_bbox(s, t.x0, t.y0, t.x1, t.y1, t.x2, t.y2, h, w)) continue;
    tri_edges(s, t.x0, t.y0, t.x1, t.y1, t.x2, t.y2);
    for (int y = s.y_lo; y <= s.y_hi; ++y)
      for (int x = s.x_lo; x <= s.x_hi; ++x) {
        float u, v, w0, w1, w2;
        pixel_uv(s, x, y, u, v);
        if (!(in_border(x, y, h, w) || uv_inside(u, v))) continue;
        const float d = weights_depth(u, v, t.z0, t.z1, t.z2, w0, w1, w2);
        if (!(d == d)) continue;
        const unsigned long long key = make_key(d, (uint32_t)i);
        if (key > keys[(size_t)y * w + x]) keys[(size_t)y * w + x] = key;
      }
  }
  for (size_t p = 0; p < npix; ++p) {
    const int i = key_triangle(keys[p]);
    if (i < 0) continue;
    const Tri t = fetch(vertices, triangles + 3 * (size_t)i);
    TriSetup s;
    tri_edges(s, t.x0, t.y0, t.x1, t.y1, t.x2, t.y2);
    float u, v, w0, w1, w2;
    pixel_uv(s, (int)(p % (size_t)w), (int)(p / (size_t)w), u, v);
    depth[p] = weights_depth(u, v, t.z0, t.z1, t.z2, w0, w1, w2);
    tri_buf[p] = i;
    weights[3 * p] = w0;
    weights[3 * p + 1] = w1;
    weights[3 * p + 2] = w2;
  }
  return 0;
}

// clip_trunc_byte (vertex stage -> colour bytes) and depth_code ordering, for property checks
extern "C" unsigned int hc_clip_trunc_byte(float c) { return clip_trunc_byte(c); }
extern "C" unsigned int hc_depth_code(float d) { return depth_code(d); }
