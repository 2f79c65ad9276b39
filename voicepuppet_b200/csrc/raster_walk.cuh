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

// The flattened walk of the boxes the lanes do not walk themselves.  q0 = (ax, ay, e0x, e0y), q1 = (e1x, e1y, d00, d01),
// q2 = (d11, inv, key low, key high), box = (x_lo, y_lo, x_hi, y_hi) of the calling lane's triangle; big: this lane has
// such a box.  All 32 lanes must call.
__device__ __forceinline__ void walk_flattened(const float4 q0_own, const float4 q1_own, const float4 q2_own,
                                               const int4 box, bool big, float4 (*rec)[32],
                                               unsigned long long* __restrict__ keys, int w, unsigned lane) {
  int nitems = 0, bw = 0, nseg = 1;
  if (big) {
    bw = box.z - box.x + 1;
    nseg = (bw + kSegPixels - 1) / kSegPixels;
    nitems = (box.w - box.y + 1) * nseg;
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
  __syncwarp();  // the previous readers are done with rec
  if (nitems > 0) {
    const int slot = __popc(live & lt_mask);  // live triangles are compacted: slot order == lane order
    rec[0][slot] = q0_own;
    rec[1][slot] = q1_own;
    rec[2][slot] = q2_own;
    rec[3][slot] = make_float4(__int_as_float(box.x), __int_as_float(box.y), __int_as_float(bw | (nseg << 16)),
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
      // isPointInTri (mesh_core.cpp:23-50) with the row terms hoisted
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

// s: bounding box + edge set-up (valid when n > 0); key: the triangle's z-buffer key; n: box pixels (0 = nothing to
// do); group: 0 or the lanes (4 / 8) that share the walk of each other's boxes, group_min: see below; rec: this warp's staging,
// float4[4][32]; keys: the frame's z-buffer.  All 32 lanes must call.
// (Testing two adjacent pixels per iteration of the lane's own loop -- 52 instead of 2 x 38 instructions -- was
// measured in round 2 and changed nothing at 512x512 / 1024x1024, profiles/r02f_pairs_chunks.txt: not kept.)
__device__ __forceinline__ void walk_boxes(const TriSetup& s, unsigned long long key, int n, int inline_max, int group,
                                           int group_min, float4 (*rec)[32], unsigned long long* __restrict__ keys, int w,
                                           unsigned lane) {
  const float4 q0_own = make_float4(s.ax, s.ay, s.e0x, s.e0y), q1_own = make_float4(s.e1x, s.e1y, s.d00, s.d01),
               q2_own = make_float4(s.d11, s.inv, __uint_as_float(static_cast<uint32_t>(key)),
                                    __uint_as_float(static_cast<uint32_t>(key >> 32)));
  // (a warp takes the group walk when its boxes hold at least group_min pixels together: below that -- 512x512 and
  // smaller, where most boxes have 1..5 pixels -- the per-box cost of the group walk exceeds what it saves)
  if (group > 0 && __reduce_add_sync(kFullWarp, (n > 0 && n <= inline_max) ? n : 0) >= group_min) {
    // GROUP WALK.  With one box per lane the 32 lanes of a REDG hit 32 different sectors, and the LSU / L2 take such a
    // spread reduction at 1.64 cycles per lane and SM (tools/diag_redg.cu) -- 2.3 of the 3.9 us per 1024x1024 frame.
    // Lanes that hit the same sector are 2-3 x cheaper (0.76 cycles per lane in runs of 4, 0.50 in runs of 32).  So
    // the `group` lanes of an aligned group walk their boxes TOGETHER, one box after the other, in row-major order:
    // lane j of the group takes pixels j, j + group, ... of the box, i.e. neighbouring lanes take neighbouring pixels
    // of a row.  The records travel through the warp's staging; pixel k of a box is (k % bw, k / bw), the division by a
    // 16-bit reciprocal (exact while k * bw < 65536).  The sums of `group` box sizes also differ less between groups
    // than the sizes between lanes do, so fewer lane slots idle.  Same expressions per pixel as everywhere.
    const unsigned bw = n > 0 ? (unsigned)(s.x_hi - s.x_lo + 1) : 1u;
    const bool inl = n > 0 && n <= inline_max && bw * (unsigned)n < 65536u;  // (also n < 65536: it travels in 16 bits)
    const unsigned magic = (65536u + bw - 1u) / bw;
    __syncwarp();  // the previous call's readers are done with rec
    rec[0][lane] = q0_own;
    rec[1][lane] = q1_own;
    rec[2][lane] = q2_own;
    rec[3][lane] = make_float4(__int_as_float(s.x_lo | (s.y_lo << 16)), __int_as_float((int)(bw | (inl ? (unsigned)n << 16 : 0u))),
                               __int_as_float((int)magic), __int_as_float(s.y_lo * w + s.x_lo));
    __syncwarp();
    const unsigned g0 = lane & ~(unsigned)(group - 1), sub = lane & (unsigned)(group - 1);
    for (int i = 0; i < group; ++i) {
      const float4 q3 = rec[3][g0 + i];
      const unsigned ni = (unsigned)__float_as_int(q3.y) >> 16;
      if (ni == 0u) continue;  // uniform over the group
      const float4 q0 = rec[0][g0 + i], q1 = rec[1][g0 + i], q2 = rec[2][g0 + i];
      const unsigned long long k64 = static_cast<unsigned long long>(__float_as_uint(q2.z)) |
                                     (static_cast<unsigned long long>(__float_as_uint(q2.w)) << 32);
      const int xy = __float_as_int(q3.x);
      const int x_lo = xy & 0xFFFF, y_lo = xy >> 16;
      const unsigned tbw = (unsigned)__float_as_int(q3.y) & 0xFFFFu, tmagic = (unsigned)__float_as_int(q3.z);
      const unsigned org = (unsigned)__float_as_int(q3.w);  // element index of the box origin in the frame
#pragma unroll 1
      for (unsigned k = sub; k < ni; k += (unsigned)group) {
        const unsigned dy = (k * tmagic) >> 16, dx = k - dy * tbw;
        const float px = VP_SUB(static_cast<float>(x_lo + (int)dx), q0.x);
        const float py = VP_SUB(static_cast<float>(y_lo + (int)dy), q0.y);
        const float d02 = VP_ADD(VP_MUL(q0.z, px), VP_MUL(q0.w, py));
        const float d12 = VP_ADD(VP_MUL(q1.x, px), VP_MUL(q1.y, py));
        const float u = VP_MUL(VP_SUB(VP_MUL(q2.x, d02), VP_MUL(q1.w, d12)), q2.y);
        const float v = VP_MUL(VP_SUB(VP_MUL(q1.z, d12), VP_MUL(q1.w, d02)), q2.y);
        if (uv_inside(u, v)) atomicMax(keys + (org + dy * (unsigned)w + dx), k64);
      }
    }
    const bool big = n > 0 && !inl;
    if (__any_sync(kFullWarp, big))
      walk_flattened(q0_own, q1_own, q2_own, make_int4(s.x_lo, s.y_lo, s.x_hi, s.y_hi), big, rec, keys, w, lane);
    return;
  }
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
  const bool big = n > inline_max;
  if (__any_sync(kFullWarp, big))
    walk_flattened(q0_own, q1_own, q2_own, make_int4(s.x_lo, s.y_lo, s.x_hi, s.y_hi), big, rec, keys, w, lane);
}

}  // namespace vp
