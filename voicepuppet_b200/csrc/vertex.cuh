// Pieces of the vertex stage shared by the vertex kernels (reconstruct.cu) and the fused vertex + raster kernel
// (fused.cu): per-frame constants, the per-thread staging state, the fan normal sum, lighting and projection of
// the raster-record-only path.
#pragma once

#include "launch.h"

namespace vp {

struct __align__(16) FrameShared {   // per-frame constants staged in shared memory
  FrameParams par;     // 192 B: rotation (float64), translation, gamma
  float rot[12];       // rotation as float32
  float sh[28];        // gamma with the SH band constants (and the 0.8 ambient) folded in: [3][9]
};

// The same frame folded further, for the raster-record-only path (no per-vertex outputs requested):
//  * geometry: rotation(s), translation, camera, focal / centre, the y flip and the raster scale are one
//    projective map of the unrotated vertex v (float64): X = v.ax + bx, Y = v.ay + by, zc = v.az + bz,
//    record = (X / zc, Y / zc, -zc)   [Reconstruction_rotation :211 + Projection_layer :100-120 + :215 + the
//    xy * res/224 convention]
//  * lighting: rotating the normal and evaluating the 9 SH bands is a quadratic form of the UNROTATED unit
//    normal n per channel: lit = c + b.n + n'Qn with b = R b_r, Q = R Q_r R' (Illumination_layer :129-168)
struct __align__(16) FrameFast {
  double lin[12];      // ax[3], bx, ay[3], by, az[3], bz
  float shq[32];       // per channel 10 values: c, bx, by, bz, qxx, qyy, qzz, 2qxy, 2qxz, 2qyz (30 used)
};

struct __align__(16) FrameConst {
  FrameShared slow;
  FrameFast fast;
};
static_assert(sizeof(FrameFast) == 224 && sizeof(FrameConst) == 576, "frame constant layout");

struct VertexArgs {
  const TileDesc* tiles;
  const int* tile_list;        // blockIdx.x -> tile id (this launch's slice of vp_model::tile_list)
  const uint32_t* fan;
  const int* slot_off;         // optional slot tables (SLOTS flavour of the fan kernel), see Topology
  const uint16_t* slot_tab;
  const uint32_t* fan_slot;
  const uint32_t* ltri;
  const int* halo;
  const uint16_t* ring;
  const int* v_int2orig;
  const double* base;
  const float* tex;
  const float* disp;
  size_t disp_stride;
  const FrameConst* fshared;   // per-frame constants prepared by frame_prep_kernel
  int nframes;
  int frames_per_block;
  int rotate_first;
  int has_out;
  double focal, center, image_size, raster_scale;
  float4* vrec;
  size_t vrec_stride;
  ReconOut out;
  int nver;
};


static_assert(sizeof(FrameShared) == 352, "FrameShared layout");

constexpr int kSlotsV = kTileLV / kTileV;  // position-staging slots per thread (2)
constexpr int kSlotsT = kTileLT / kTileV;  // triangle slots per thread of the generic kernel (4)

// Per-thread state shared by both kernels: the thread's local vertices (own first) and how their
// positions are staged.  Positions are float32 relative to the tile's first vertex: (base - origin) is
// rounded once per tile (|.| ~ tile extent, so its float32 error is ~1e-8 of the mesh scale) and the float32
// expression displacement is added; the edges the normals are made of are differences of nearby points,
// so this keeps their cancellation error at the level of the displacement's own float32 rounding.
struct LocalVerts {
  int gv[kSlotsV];
  float rx[kSlotsV], ry[kSlotsV], rz[kSlotsV];  // base - origin
  float dx[kSlotsV], dy[kSlotsV], dz[kSlotsV];  // displacement of the frame staged next
  double bx, by, bz;                            // own vertex, float64

  __device__ __forceinline__ void load(const VertexArgs& a, const TileDesc& td, int tid) {
    const double ox = __ldg(a.base + 3 * (size_t)td.v_begin), oy = __ldg(a.base + 3 * (size_t)td.v_begin + 1),
                 oz = __ldg(a.base + 3 * (size_t)td.v_begin + 2);
    bx = by = bz = 0.0;
#pragma unroll
    for (int q = 0; q < kSlotsV; ++q) {
      const int i = tid + q * kTileV;
      gv[q] = td.v_begin;
      rx[q] = ry[q] = rz[q] = 0.f;
      dx[q] = dy[q] = dz[q] = 0.f;
      if (i < td.nlv) {
        gv[q] = (i < td.nv) ? td.v_begin + i : __ldg(a.halo + td.halo_off + i - td.nv);
        const double x = __ldg(a.base + 3 * (size_t)gv[q]), y = __ldg(a.base + 3 * (size_t)gv[q] + 1),
                     z = __ldg(a.base + 3 * (size_t)gv[q] + 2);
        rx[q] = (float)(x - ox);
        ry[q] = (float)(y - oy);
        rz[q] = (float)(z - oz);
        if (q == 0) {
          bx = x;
          by = y;
          bz = z;
        }
      }
    }
  }
  __device__ __forceinline__ void fetch(const VertexArgs& a, int f, int nq_v) {  // displacement of frame f
    if (a.disp == nullptr) return;
#pragma unroll
    for (int q = 0; q < kSlotsV; ++q)
      if (q < nq_v) {
        const float* d = a.disp + (size_t)f * a.disp_stride + 3 * (size_t)gv[q];
        dx[q] = __ldg(d);
        dy[q] = __ldg(d + 1);
        dz[q] = __ldg(d + 2);
      }
  }
  __device__ __forceinline__ void stage(float4* pos, int tid, int nq_v) const {
#pragma unroll
    for (int q = 0; q < kSlotsV; ++q)
      if (q < nq_v) pos[tid + q * kTileV] = make_float4(rx[q] + dx[q], ry[q] + dy[q], rz[q] + dz[q], 0.f);
  }
  // same, at explicit byte offsets (two 16-bit halves of `offs`): the bank-conflict-aware slot placement
  __device__ __forceinline__ void stage_at(char* pos, uint32_t offs, int nq_v) const {
#pragma unroll
    for (int q = 0; q < kSlotsV; ++q)
      if (q < nq_v)
        *reinterpret_cast<float4*>(pos + ((offs >> (16 * q)) & 0xFFFFu)) =
            make_float4(rx[q] + dx[q], ry[q] + dy[q], rz[q] + dz[q], 0.f);
  }
  // the staged position of the thread's first local vertex (its own vertex), kept in registers by the fan kernel
  __device__ __forceinline__ float3 own_staged() const { return make_float3(rx[0] + dx[0], ry[0] + dy[0], rz[0] + dz[0]); }
};

__device__ __forceinline__ void stage_frame_constants(const VertexArgs& a, FrameShared* dst, int f, int tid) {
  if (tid < (int)(sizeof(FrameShared) / 4))
    reinterpret_cast<uint32_t*>(dst)[tid] = __ldg(reinterpret_cast<const uint32_t*>(&a.fshared[f].slow) + tid);
}
__device__ __forceinline__ void stage_frame_constants(const VertexArgs& a, FrameFast* dst, int f, int tid) {
  if (tid < (int)(sizeof(FrameFast) / 4))
    reinterpret_cast<uint32_t*>(dst)[tid] = __ldg(reinterpret_cast<const uint32_t*>(&a.fshared[f].fast) + tid);
}

// Lighting of the raster-record-only path: normalise the summed face normal, evaluate the folded quadratic form
// per channel (FrameFast::shq), multiply by the texture, clip / truncate to bytes (infer_bfmvid.py:98,105).
// 6 + 27 float32 operations, spelled with explicit fmaf so that every kernel using it rounds identically.
__device__ __forceinline__ uint32_t light_fast(const FrameFast& ff, float nx, float ny, float nz, float tr, float tg,
                                               float tb) {
  {
    const float inv = rsqrtf(fmaf(nx, nx, fmaf(ny, ny, nz * nz)));  // 0 * inf -> NaN for a vertex without faces
    nx *= inv;
    ny *= inv;
    nz *= inv;
  }
  const float xx = nx * nx, yy = ny * ny, zz = nz * nz, xy = nx * ny, xz = nx * nz, yz = ny * nz;
  float lit[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* g = ff.shq + 10 * c;
    float acc = fmaf(g[1], nx, g[0]);
    acc = fmaf(g[2], ny, acc);
    acc = fmaf(g[3], nz, acc);
    acc = fmaf(g[4], xx, acc);
    acc = fmaf(g[5], yy, acc);
    acc = fmaf(g[6], zz, acc);
    acc = fmaf(g[7], xy, acc);
    acc = fmaf(g[8], xz, acc);
    lit[c] = fmaf(g[9], yz, acc);
  }
  return clip_trunc_byte(lit[0] * tr) | (clip_trunc_byte(lit[1] * tg) << 8) | (clip_trunc_byte(lit[2] * tb) << 16);
}

// Geometry of the raster-record-only path: the folded projective map (FrameFast::lin) of the unrotated vertex in
// float64 (9 DFMA + one reciprocal + 2 DMUL), then (x, S - y, -z) -> float32 as infer_bfmvid.py:93-103 casts them.
// Explicit fma: a halo vertex projected by a neighbouring tile gets bit-identical coordinates.
__device__ __forceinline__ float3 project_fast(const FrameFast& ff, double vx, double vy, double vz) {
  const double* L = ff.lin;
  const double X = fma(vx, L[0], fma(vy, L[1], fma(vz, L[2], L[3])));
  const double Y = fma(vx, L[4], fma(vy, L[5], fma(vz, L[6], L[7])));
  const double zc = fma(vx, L[8], fma(vy, L[9], fma(vz, L[10], L[11])));
  const double inv = 1.0 / zc;
  return make_float3((float)(X * inv), (float)(Y * inv), (float)(-zc));
}

// Summed face normal of an own vertex from its fan record (launch.h): 9 gathers of staged positions,
// sum over the marked pairs of (u_i - v) x (u_i+1 - v).  pos: the frame's staged positions (bytes), pv: the own vertex.
__device__ __forceinline__ void fan_normal_sum(const char* pos, const uint32_t (&fan)[kFanWords], const float3 pv,
                                               float& nx, float& ny, float& nz) {
  nx = ny = nz = 0.f;
  float4 p = *reinterpret_cast<const float4*>(pos + (fan[0] & 0xFFFFu));
  float ex = p.x - pv.x, ey = p.y - pv.y, ez = p.z - pv.z;
  const uint32_t mask = fan[4] >> 16;
#pragma unroll
  for (int i = 0; i < kFanEntries - 1; ++i) {
    const uint32_t off = ((i + 1) & 1) ? (fan[(i + 1) >> 1] >> 16) : (fan[(i + 1) >> 1] & 0xFFFFu);
    p = *reinterpret_cast<const float4*>(pos + off);
    const float gx = p.x - pv.x, gy = p.y - pv.y, gz = p.z - pv.z;
    if (mask & (1u << i)) {
      nx += ey * gz - ez * gy;
      ny += ez * gx - ex * gz;
      nz += ex * gy - ey * gx;
    }
    ex = gx;
    ey = gy;
    ez = gz;
  }
}

}  // namespace vp
