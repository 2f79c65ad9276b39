// Warp-collective walk of one triangle bounding box per lane against the 64-bit z-buffer (render_colors semantics:
// flat depth key, mesh_core.cpp:205-229), shared by the packed scatter kernel (raster.cuh) and the fused vertex +
// raster kernel (fused.cu).
//   * boxes of up to `inline_max` pixels are walked by their own lane in one flat loop (at 256x256 the mean box is
//     1.1 pixels: anything cleverer costs more than it saves);
//   * larger boxes are FLATTENED: the warp's (triangle, box row) pairs -- rows wider than 16 pixels cut into
//     segments -- are numbered by a warp scan and dealt out one per lane, the triangle set-ups being staged in shared
//     memory, so a lane walks at most 16 pixels of one row whatever the triangle sizes are.  This replaces the
//     round-1 scheme (large boxes broadcast one at a time and walked by the whole warp), which ran the 1024x1024
//     configuration (mean box 17 pixels, max 64) at 11 us per frame.
// Every float operation of the inside test is individually rounded in the reference's order (vp_math.cuh).
#pragma once

#include <cuda_runtime.h>

#include "vp_math.cuh"

namespace vp {

constexpr int kSegPixels = 16;  // a (triangle, row) item wider than this is cut into segments
constexpr unsigned kFullWarp = 0xFFFFFFFFu;

// s: bounding box + edge set-up (valid when n > 0); key: the triangle's z-buffer key; n: box pixels (0 = nothing to
// do); rec: this warp's staging, float4[4][32]; keys: the frame's z-buffer.  All 32 lanes must call.
// (Testing two adjacent pixels per iteration of the lane's own loop -- 52 instead of 2 x 38 instructions -- was
// measured in round 2 and changed nothing at 512x512 / 1024x1024, profiles/r02f_pairs_chunks.txt: not kept.)
__device__ __forceinline__ void walk_boxes(const TriSetup& s, unsigned long long key, int n, int inline_max,
                                           float4 (*rec)[32], unsigned long long* __restrict__ keys, int w,
                                           unsigned lane) {
  if (n > 0 && n <= inline_max) {
    // one flat loop over the box: row terms are refreshed when x wraps
    int x = s.x_lo, y = s.y_lo;
    float py = VP_SUB(static_cast<float>(y), s.ay);
    float m0y = VP_MUL(s.e0y, py), m1y = VP_MUL(s.e1y, py);
    unsigned long long* row = keys + (unsigned)y * (unsigned)w;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      const float px = VP_SUB(static_cast<float>(x), s.ax);
      const float d02 = VP_ADD(VP_MUL(s.e0x, px), m0y);
      const float d12 = VP_ADD(VP_MUL(s.e1x, px), m1y);
      const float u = VP_MUL(VP_SUB(VP_MUL(s.d11, d02), VP_MUL(s.d01, d12)), s.inv);
      const float v = VP_MUL(VP_SUB(VP_MUL(s.d00, d12), VP_MUL(s.d01, d02)), s.inv);
      if (uv_inside(u, v)) atomicMax(row + x, key);
      if (++x > s.x_hi) {
        x = s.x_lo;
        ++y;
        py = VP_SUB(static_cast<float>(y), s.ay);
        m0y = VP_MUL(s.e0y, py);
        m1y = VP_MUL(s.e1y, py);
        row += w;
      }
    }
  }
  int nitems = 0, bw = 0, nseg = 1;
  if (n > inline_max) {
    bw = s.x_hi - s.x_lo + 1;
    nseg = (bw + kSegPixels - 1) / kSegPixels;
    nitems = (s.y_hi - s.y_lo + 1) * nseg;
  }
  const unsigned live = __ballot_sync(kFullWarp, nitems > 0);
  if (live == 0u) return;
  const unsigned lt_mask = (1u << lane) - 1u, le_mask = kFullWarp >> (31u - lane);
  int incl = nitems;  // inclusive scan of the item counts
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(kFullWarp, incl, d);
    if ((int)lane >= d) incl += t;
  }
  const int excl = incl - nitems;
  const int total = __shfl_sync(kFullWarp, incl, 31);
  __syncwarp();  // the previous call's readers are done with rec
  if (nitems > 0) {
    const int slot = __popc(live & lt_mask);  // live triangles are compacted: slot order == lane order
    rec[0][slot] = make_float4(s.ax, s.ay, s.e0x, s.e0y);
    rec[1][slot] = make_float4(s.e1x, s.e1y, s.d00, s.d01);
    rec[2][slot] = make_float4(s.d11, s.inv, __uint_as_float(static_cast<uint32_t>(key)),
                               __uint_as_float(static_cast<uint32_t>(key >> 32)));
    rec[3][slot] = make_float4(__int_as_float(s.x_lo), __int_as_float(s.y_lo), __int_as_float(bw | (nseg << 16)),
                               __int_as_float(excl));
  }
  __syncwarp();
  const bool has = nitems > 0;
  for (int base = 0; base < total; base += 32) {
    // slot of item base + lane: the last live triangle starting at or before `base`, plus the number of
    // triangles starting inside (base, base + lane]
    const unsigned le = __ballot_sync(kFullWarp, has && excl <= base);
    const unsigned starts =
        __reduce_or_sync(kFullWarp, (has && excl > base && excl < base + 32) ? (1u << (excl - base)) : 0u);
    const int item = base + (int)lane;
    if (item < total) {
      const int sl = __popc(le) - 1 + __popc(starts & le_mask);
      const float4 q3 = rec[3][sl];
      const int x_lo = __float_as_int(q3.x), y_lo = __float_as_int(q3.y), bwn = __float_as_int(q3.z);
      const int local = item - __float_as_int(q3.w);
      const int tbw = bwn & 0xFFFF, tns = bwn >> 16;
      int r = local, seg = 0;
      if (tns > 1) {
        r = local / tns;
        seg = local - r * tns;
      }
      const int y = y_lo + r;
      int x = x_lo + seg * kSegPixels;
      const int xe = min(x_lo + tbw - 1, x + kSegPixels - 1);
      const float4 q0 = rec[0][sl], q1 = rec[1][sl], q2 = rec[2][sl];
      const unsigned long long k = static_cast<unsigned long long>(__float_as_uint(q2.z)) |
                                   (static_cast<unsigned long long>(__float_as_uint(q2.w)) << 32);
      // isPointInTri (mesh_core.cpp:23-50) with the row terms hoisted; q0 = (ax, ay, e0x, e0y),
      // q1 = (e1x, e1y, d00, d01), q2 = (d11, inv, key)
      const float py = VP_SUB(static_cast<float>(y), q0.y);
      const float m0y = VP_MUL(q0.w, py), m1y = VP_MUL(q1.y, py);
      unsigned long long* rowp = keys + (unsigned)y * (unsigned)w;
#pragma unroll 1
      for (; x <= xe; ++x) {
        const float px = VP_SUB(static_cast<float>(x), q0.x);
        const float d02 = VP_ADD(VP_MUL(q0.z, px), m0y);
        const float d12 = VP_ADD(VP_MUL(q1.x, px), m1y);
        const float u = VP_MUL(VP_SUB(VP_MUL(q2.x, d02), VP_MUL(q1.w, d12)), q2.y);
        const float v = VP_MUL(VP_SUB(VP_MUL(q1.z, d12), VP_MUL(q1.w, d02)), q2.y);
        if (uv_inside(u, v)) atomicMax(rowp + x, k);
      }
    }
  }
}

}  // namespace vp
