// Shared arithmetic of the rasterizer and the vertex stage.
//
// Everything here is __host__ __device__ so that tests/hostcheck can run the very same
// expressions on the CPU and compare them with the oracle before any GPU time is spent.
//
// Bit-exactness contract for the rasterizer (reference utils/cython/mesh_core.cpp:23-82,
// 132-136, 204, 219; mesh_core.h:19-30): the reference is built without FMA, so every
// float32 operation is individually rounded in source association order.  On the device we
// therefore spell every operation with __f*_rn intrinsics, which nvcc never contracts into
// FMA regardless of -fmad.  On the host the file must be compiled with -ffp-contract=off.
#pragma once

#include <cstdint>
#include <cmath>
#include <climits>

#if defined(__CUDA_ARCH__)
#define VP_MUL(a, b) __fmul_rn((a), (b))
#define VP_ADD(a, b) __fadd_rn((a), (b))
#define VP_SUB(a, b) __fsub_rn((a), (b))
#define VP_DIV(a, b) __fdiv_rn((a), (b))
#else
#define VP_MUL(a, b) ((a) * (b))
#define VP_ADD(a, b) ((a) + (b))
#define VP_SUB(a, b) ((a) - (b))
#define VP_DIV(a, b) ((a) / (b))
#endif

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

#define VP_HD __host__ __device__ __forceinline__

namespace vp {

constexpr uint32_t kNoTri = 0xFFFFFFFFu;
constexpr float kInitDepth = -99999.0f;  // infer_bfmvid.py:106

// float -> int the way x86 cvttss2si does it (what the compiled reference's (int) casts do):
// NaN and out-of-range give INT_MIN.
VP_HD int trunc_x86(float f) {
  return (f >= -2147483648.0f && f < 2147483648.0f) ? static_cast<int>(f) : INT_MIN;
}

// libstdc++ std::min / std::max on floats: (b < a) ? b : a  and  (a < b) ? b : a
VP_HD float lo2(float a, float b) { return (b < a) ? b : a; }
VP_HD float hi2(float a, float b) { return (a < b) ? b : a; }

struct TriSetup {
  float ax, ay;         // corner 0
  float e0x, e0y;       // corner 2 - corner 0   (mesh_core.cpp:27)
  float e1x, e1y;       // corner 1 - corner 0   (mesh_core.cpp:28)
  float d00, d01, d11;  // mesh_core.cpp:32-35
  float inv;            // mesh_core.cpp:39-43
  int x_lo, x_hi, y_lo, y_hi;
};

// Bounding box clamp of mesh_core.cpp:132-141 / 194-203.  Returns false when empty.
VP_HD bool tri_bbox(TriSetup& s, float x0, float y0, float x1, float y1, float x2, float y2, int h,
                    int w) {
  int v;
  v = trunc_x86(ceilf(lo2(x0, lo2(x1, x2))));
  s.x_lo = v > 0 ? v : 0;
  v = trunc_x86(floorf(hi2(x0, hi2(x1, x2))));
  s.x_hi = v < w - 1 ? v : w - 1;
  v = trunc_x86(ceilf(lo2(y0, lo2(y1, y2))));
  s.y_lo = v > 0 ? v : 0;
  v = trunc_x86(floorf(hi2(y0, hi2(y1, y2))));
  s.y_hi = v < h - 1 ? v : h - 1;
  return !(s.x_hi < s.x_lo || s.y_hi < s.y_lo);
}

// Pixel-independent part of isPointInTri / get_point_weight.
VP_HD void tri_edges(TriSetup& s, float x0, float y0, float x1, float y1, float x2, float y2) {
  s.ax = x0;
  s.ay = y0;
  s.e0x = VP_SUB(x2, x0);
  s.e0y = VP_SUB(y2, y0);
  s.e1x = VP_SUB(x1, x0);
  s.e1y = VP_SUB(y1, y0);
  s.d00 = VP_ADD(VP_MUL(s.e0x, s.e0x), VP_MUL(s.e0y, s.e0y));
  s.d01 = VP_ADD(VP_MUL(s.e0x, s.e1x), VP_MUL(s.e0y, s.e1y));
  s.d11 = VP_ADD(VP_MUL(s.e1x, s.e1x), VP_MUL(s.e1y, s.e1y));
  const float den = VP_SUB(VP_MUL(s.d00, s.d11), VP_MUL(s.d01, s.d01));
  s.inv = (den == 0.0f) ? 0.0f : VP_DIV(1.0f, den);
}

// Barycentric (u, v) of the integer pixel (x, y): mesh_core.cpp:29,34,36,45-46.
VP_HD void pixel_uv(const TriSetup& s, int x, int y, float& u, float& v) {
  const float px = VP_SUB(static_cast<float>(x), s.ax);
  const float py = VP_SUB(static_cast<float>(y), s.ay);
  const float d02 = VP_ADD(VP_MUL(s.e0x, px), VP_MUL(s.e0y, py));
  const float d12 = VP_ADD(VP_MUL(s.e1x, px), VP_MUL(s.e1y, py));
  u = VP_MUL(VP_SUB(VP_MUL(s.d11, d02), VP_MUL(s.d01, d12)), s.inv);
  v = VP_MUL(VP_SUB(VP_MUL(s.d00, d12), VP_MUL(s.d01, d02)), s.inv);
}

VP_HD bool uv_inside(float u, float v) { return (u >= 0.0f) && (v >= 0.0f) && (VP_ADD(u, v) < 1.0f); }

// mesh_core.cpp:204: (z0 + z1 + z2) / 3, left to right.
VP_HD float flat_depth(float z0, float z1, float z2) { return VP_DIV(VP_ADD(VP_ADD(z0, z1), z2), 3.0f); }

// mesh_core.cpp:79-81 and :151.
VP_HD float weights_depth(float u, float v, float z0, float z1, float z2, float& w0, float& w1,
                          float& w2) {
  w0 = VP_SUB(VP_SUB(1.0f, u), v);
  w1 = v;
  w2 = u;
  return VP_ADD(VP_ADD(VP_MUL(w0, z0), VP_MUL(w1, z1)), VP_MUL(w2, z2));
}

// mesh_core.cpp:148: the two-pixel frame qualifies without the inside test.
VP_HD bool in_border(int x, int y, int h, int w) {
  const float fx = static_cast<float>(x), fy = static_cast<float>(y);
  return fx < 2.0f || fx > static_cast<float>(w - 3) || fy < 2.0f || fy > static_cast<float>(h - 3);
}

// mesh_core.cpp:219-221: float sum -> (int) -> integer /3 -> float -> unsigned char.
VP_HD unsigned char flat_color(float c0, float c1, float c2) {
  const float sum = VP_ADD(VP_ADD(c0, c1), c2);
  const float pc = static_cast<float>(trunc_x86(sum) / 3);
  return static_cast<unsigned char>(trunc_x86(pc) & 0xFF);
}

// Order-preserving float -> uint32 (larger depth -> larger code), -0 == +0.  Caller filters NaN.
VP_HD uint32_t depth_code(float d) {
  if (d == 0.0f) d = 0.0f;
#if defined(__CUDA_ARCH__)
  const uint32_t b = __float_as_uint(d);
#else
  uint32_t b;
  __builtin_memcpy(&b, &d, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// 64-bit z-buffer key: larger wins under atomicMax.  High word = depth code, low word orders
// equal depths by ascending triangle index (the sequential loop's strict '>' keeps the first).
// Low word kNoTri marks "the caller's initial depth": it beats every triangle at equal depth.
VP_HD unsigned long long make_key(float d, uint32_t tri) {
  return (static_cast<unsigned long long>(depth_code(d)) << 32) |
         static_cast<unsigned long long>(kNoTri - 1u - tri);
}
VP_HD unsigned long long init_key(float d) {
  if (d != d) return ~0ull;  // NaN in the caller's buffer: "cand > depth" is never true
  return (static_cast<unsigned long long>(depth_code(d)) << 32) | kNoTri;
}
// -1 when the pixel keeps the caller's value (or key == 0: never touched).
VP_HD int key_triangle(unsigned long long k) {
  const uint32_t lo = static_cast<uint32_t>(k);
  if (k == 0ull || lo == kNoTri) return -1;
  return static_cast<int>(kNoTri - 1u - lo);
}

// ---------------------------------------------------------------------------------------
// Vertex stage (reference utils/reconstruct_mesh.py).  Tolerance contract, not bit-exact:
// geometry in float64 like the oracle, lighting in float32.
// ---------------------------------------------------------------------------------------

struct FrameParams {  // == vp_frame_params in include/vpb200.h
  double rot[9];
  float trans[3];
  float gamma[27];
};

// v @ R for a row vector (np.matmul(face_shape, rotation), reconstruct_mesh.py:111,211)
VP_HD void rotate_row(const double* R, double x, double y, double z, double& ox, double& oy,
                      double& oz) {
  ox = x * R[0] + y * R[3] + z * R[6];
  oy = x * R[1] + y * R[4] + z * R[7];
  oz = x * R[2] + y * R[5] + z * R[8];
}

// Projection_layer, reconstruct_mesh.py:100-120: returns image-plane (x, y) and z_buffer = -z'.
VP_HD void project(const double* R, const float* t, double focal, double center, double sx, double sy,
                   double sz, double& px, double& py, double& zbuf) {
  double rx, ry, rz;
  rotate_row(R, sx, sy, sz, rx, ry, rz);
  rx += static_cast<double>(t[0]);
  ry += static_cast<double>(t[1]);
  rz += static_cast<double>(t[2]);
  const double zc = 10.0 - rz;  // reverse_z then + camera_pos
  const double inv = 1.0 / zc;  // (f*x + c*z) / z with one reciprocal: differs from the division by <= 1 ulp of float64
  px = (focal * rx + center * zc) * inv;
  py = (focal * ry + center * zc) * inv;
  zbuf = -zc;
}

// Illumination_layer, reconstruct_mesh.py:129-168, for one vertex; n = rotated unit normal.
// lit[c] = sum_k Y_k * (gamma[c][k] + 0.8 * (k == 0)).
template <typename T>
VP_HD void sh_lighting(const float* gamma, T nx, T ny, T nz, T* lit) {
  // products of a0=pi, a1=2pi/sqrt3, a2=2pi/sqrt8, c0=1/sqrt(4pi), c1=sqrt3/sqrt(4pi),
  // c2=3sqrt5/sqrt(12pi), evaluated in float64 as the reference does (:138-153)
  const T k0 = static_cast<T>(0.8862269254527579);   // a0*c0
  const T k1 = static_cast<T>(1.772453850905516);    // a1*c1
  const T k2 = static_cast<T>(2.4270323906946243);   // a2*c2
  const T k6 = static_cast<T>(0.7006239020497412);   // a2*c2*0.5/sqrt(3)
  const T k8 = static_cast<T>(1.2135161953473121);   // a2*c2*0.5
  T Y[9];
  Y[0] = k0;
  Y[1] = -k1 * ny;
  Y[2] = k1 * nz;
  Y[3] = -k1 * nx;
  Y[4] = k2 * nx * ny;
  Y[5] = -k2 * ny * nz;
  Y[6] = k6 * (static_cast<T>(3) * nz * nz - static_cast<T>(1));
  Y[7] = -k2 * nx * nz;
  Y[8] = k8 * (nx * nx - ny * ny);
  for (int c = 0; c < 3; ++c) {
    T acc = Y[0] * (static_cast<T>(gamma[9 * c]) + static_cast<T>(0.8));
    for (int k = 1; k < 9; ++k) acc += Y[k] * static_cast<T>(gamma[9 * c + k]);
    lit[c] = acc;
  }
}

// np.clip(color, 0, 255).astype(np.int32) of infer_bfmvid.py:98, as a byte.
VP_HD unsigned int clip_trunc_byte(float c) {
  if (!(c > 0.0f)) return 0u;  // also NaN
  if (c >= 255.0f) return 255u;
  return static_cast<unsigned int>(c);
}

}  // namespace vp
