// K23: vertex stage and z-buffer scatter in ONE kernel, plus the resolve pass that goes with it.
//
// The separate kernels (reconstruct.cu K2 -> raster.cuh K3) hand the projected vertices over through a
// float4 record per vertex and frame in global memory, and K3 then pays three dependent random gathers of
// those records per triangle before it can start (ncu r01z: the dominant kernel, instruction-issue bound on
// that set-up).  A vertex tile already holds its own and halo vertices in shared memory, and triangles are
// sorted by their smallest vertex, so every tile OWNS a contiguous run of triangles whose corners are all
// local to it.  One CTA per (tile, run of frames) therefore does, per frame:
//   C  own vertices: fan normal -> lighting -> packed colour, stored to vcol[frame][vertex] (4 B);
//      every local vertex (own + halo): the folded float64 projective map -> (x, S - y, -z) float32 in shared
//      memory (a halo vertex is projected again by each tile that needs it: same inputs, explicit fma, so the
//      coordinates are bit-identical everywhere -- no cracks between tiles);
//   E  owned triangles: corners from shared memory, bounding box, edge set-up (mesh_core.cpp:194-204, 23-50),
//      then the warp's (triangle, bounding-box row) pairs are flattened and dealt out one per lane, so a lane
//      walks at most 16 pixels of one row whatever the triangle sizes are -- 1-pixel boxes at 256x256 and 17-pixel
//      boxes at 1024x1024 keep the lanes equally busy; winners go to the global 64-bit z-buffer with RED.MAX.
// The flat colour needs all three corner colours and a halo corner's colour is only known to its own tile, so it
// moves to the resolve pass: winner -> corner ids (one 16-byte gather by ORIGINAL triangle index) -> three 4-byte
// colour gathers -> (c0 + c1 + c2) / 3 per byte lane.  Neither the vertex records nor the per-triangle colours
// exist any more.
// Reference: utils/reconstruct_mesh.py:35-52,100-168, utils/cython/mesh_core.cpp:169-231.
#include <algorithm>
#include <cstdlib>

#include "launch.h"
#include "raster_keys.cuh"
#include "raster_walk.cuh"
#include "vertex.cuh"

namespace vp {

struct FusedArgs {
  VertexArgs v;                 // tiles, tile_list, fan, halo, base, tex, disp, fshared, nframes, frames_per_block
  const int* own_tri_off;       // [ntiles + 1] run of owned triangles per tile
  const uint32_t* own_ltri;     // [ntri] corners as 3 x 10-bit tile-local vertex indices
  const int4* tri;              // [ntri] .w = ORIGINAL triangle index (the tie-break)
  uint32_t* vcol;               // [frames][vcol_stride] packed RGB per vertex
  unsigned vcol_stride;
  unsigned long long* keys;     // [frames][h * w]
  EpochKey km;
  int h, w;
  int inline_max, group, group_min;  // see raster_walk.cuh
};

// E phase for one frame: the tile's owned triangles against the frame's z-buffer.
// scr: projected local vertices (x, y, z, -); s_tri: per owned triangle (packed local corners, inverted
// original index = the key's low bits); rec: this warp's 4 x 32 float4 staging of triangle set-ups.
__device__ __forceinline__ void raster_owned(const float4* __restrict__ scr, const uint2* __restrict__ s_tri, int nt,
                                             float4 (*rec)[32], unsigned long long* __restrict__ keys,
                                             const EpochKey& km, int h, int w, int tid, int inline_max, int group, int group_min) {
  const unsigned lane = (unsigned)tid & 31u;
  const int wm1 = w - 1, hm1 = h - 1;
  for (int p0 = 0; p0 < nt; p0 += kTileV) {  // nt is CTA-uniform: warp-uniform trip count
    const int j = p0 + tid;
    int n = 0;
    TriSetup s;
    s.x_lo = s.y_lo = s.x_hi = s.y_hi = 0;
    unsigned long long key = 0ull;
    if (j < nt) {
      const uint2 tl = s_tri[j];
      const float4 v0 = scr[tl.x & 1023u], v1 = scr[(tl.x >> 10) & 1023u], v2 = scr[(tl.x >> 20) & 1023u];
      const float kBig = 1073741824.0f;  // 2^30: below it ceil/floor and the int casts are exact and in range
      const bool tame = fabsf(v0.x) < kBig && fabsf(v1.x) < kBig && fabsf(v2.x) < kBig && fabsf(v0.y) < kBig &&
                        fabsf(v1.y) < kBig && fabsf(v2.y) < kBig;  // false for NaN / inf
      bool nonempty;
      if (tame) {
        // mesh_core.cpp:194-203 for finite coordinates: min / max are order independent, the casts exact
        s.x_lo = max(__float2int_ru(fminf(v0.x, fminf(v1.x, v2.x))), 0);
        s.x_hi = min(__float2int_rd(fmaxf(v0.x, fmaxf(v1.x, v2.x))), wm1);
        s.y_lo = max(__float2int_ru(fminf(v0.y, fminf(v1.y, v2.y))), 0);
        s.y_hi = min(__float2int_rd(fmaxf(v0.y, fmaxf(v1.y, v2.y))), hm1);
        nonempty = s.x_hi >= s.x_lo && s.y_hi >= s.y_lo;
      } else {
        nonempty = tri_bbox(s, v0.x, v0.y, v1.x, v1.y, v2.x, v2.y, h, w);
      }
      if (nonempty) {
        const float d = flat_depth(v0.z, v1.z, v2.z);
        if (d > kInitDepth) {  // mesh_core.cpp:211 against the constant initial depth; false for NaN
          const uint32_t b = __float_as_uint(__fadd_rn(d, 0.0f));  // -0 -> +0
          const uint32_t code = b ^ (static_cast<uint32_t>(static_cast<int>(b) >> 31) | 0x80000000u);
          key = km.epoch_field | (static_cast<unsigned long long>(code) << km.tri_bits) |
                static_cast<unsigned long long>(tl.y);
          tri_edges(s, v0.x, v0.y, v1.x, v1.y, v2.x, v2.y);
          n = (s.x_hi - s.x_lo + 1) * (s.y_hi - s.y_lo + 1);
        }
      }
    }
    walk_boxes(s, key, n, inline_max, group, group_min, rec, keys, w, lane);
  }
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(kTileV, MIN_BLOCKS) fused_tile_kernel(const FusedArgs a) {
  __shared__ float4 s_pos[2][kTileLV];   // staged positions (float32, tile-relative), frame f / f + 1
  __shared__ float4 s_scr[2][kTileLV];   // projected local vertices of frame f / f + 1
  __shared__ __align__(16) FrameFast s_frame[2];
  __shared__ uint2 s_tri[kTileLT];
  __shared__ float4 s_rec[kTileV / 32][4][32];

  const VertexArgs& va = a.v;
  const int tile_id = __ldg(va.tile_list + blockIdx.x);
  const TileDesc td = va.tiles[tile_id];
  const int tid = threadIdx.x;
  const int nq_v = (td.nlv + kTileV - 1) / kTileV;  // CTA-uniform
  const int f_begin = blockIdx.y * va.frames_per_block;
  const int f_end = min(va.nframes, f_begin + va.frames_per_block);
  if (f_begin >= f_end) return;

  const int tri_begin = __ldg(a.own_tri_off + tile_id);
  const int nt = __ldg(a.own_tri_off + tile_id + 1) - tri_begin;
  for (int j = tid; j < nt; j += kTileV)
    s_tri[j] = make_uint2(__ldg(a.own_ltri + tri_begin + j), a.km.tri_mask - (uint32_t)__ldg(&a.tri[tri_begin + j].w));

  LocalVerts lv;
  lv.load(va, td, tid);
  const bool own = tid < td.nv;
  const bool has0 = tid < td.nlv, has1 = tid + kTileV < td.nlv;
  double hx = 0.0, hy = 0.0, hz = 0.0;  // float64 base of the second local vertex (always a halo vertex)
  if (has1) {
    hx = __ldg(va.base + 3 * (size_t)lv.gv[1]);
    hy = __ldg(va.base + 3 * (size_t)lv.gv[1] + 1);
    hz = __ldg(va.base + 3 * (size_t)lv.gv[1] + 2);
  }
  uint32_t fan[kFanWords] = {0, 0, 0, 0, 0};
  float tr = 0.f, tg = 0.f, tb = 0.f;
  if (own) {
#pragma unroll
    for (int k = 0; k < kFanWords; ++k) fan[k] = __ldg(va.fan + (size_t)lv.gv[0] * kFanWords + k);
    if (va.tex) {
      tr = __ldg(va.tex + 3 * (size_t)lv.gv[0]);
      tg = __ldg(va.tex + 3 * (size_t)lv.gv[0] + 1);
      tb = __ldg(va.tex + 3 * (size_t)lv.gv[0] + 2);
    }
  }

  // prologue: frame f_begin staged, displacement of frame f_begin + 1 in flight
  lv.fetch(va, f_begin, nq_v);
  stage_frame_constants(va, &s_frame[0], f_begin, tid);
  lv.stage(s_pos[0], tid, nq_v);
  float3 pv = lv.own_staged();
  float c0x = lv.dx[0], c0y = lv.dy[0], c0z = lv.dz[0];  // displacements of the frame being finished
  float c1x = lv.dx[1], c1y = lv.dy[1], c1z = lv.dz[1];
  if (f_begin + 1 < f_end) lv.fetch(va, f_begin + 1, nq_v);
  __syncthreads();

  const size_t npix = (size_t)a.h * a.w;
  for (int f = f_begin; f < f_end; ++f) {
    const int buf = (f - f_begin) & 1;
    // ---- C: colours of the own vertices, projection of every local vertex ---------------------------
    if (own) {
      float nx, ny, nz;
      fan_normal_sum(reinterpret_cast<const char*>(s_pos[buf]), fan, pv, nx, ny, nz);
      a.vcol[(size_t)f * a.vcol_stride + lv.gv[0]] = light_fast(s_frame[buf], nx, ny, nz, tr, tg, tb);
    }
    if (has0) {
      const float3 p = project_fast(s_frame[buf], lv.bx + (double)c0x, lv.by + (double)c0y, lv.bz + (double)c0z);
      s_scr[buf][tid] = make_float4(p.x, p.y, p.z, 0.f);
    }
    if (has1) {
      const float3 p = project_fast(s_frame[buf], hx + (double)c1x, hy + (double)c1y, hz + (double)c1z);
      s_scr[buf][tid + kTileV] = make_float4(p.x, p.y, p.z, 0.f);
    }
    // ---- stage frame f + 1 into the other buffers (their readers passed the previous barrier) ----
    if (f + 1 < f_end) {
      stage_frame_constants(va, &s_frame[buf ^ 1], f + 1, tid);
      lv.stage(s_pos[buf ^ 1], tid, nq_v);
      pv = lv.own_staged();
      c0x = lv.dx[0];
      c0y = lv.dy[0];
      c0z = lv.dz[0];
      c1x = lv.dx[1];
      c1y = lv.dy[1];
      c1z = lv.dz[1];
      if (f + 2 < f_end) lv.fetch(va, f + 2, nq_v);
    }
    __syncthreads();
    // ---- E: the tile's own triangles of frame f (s_scr[buf] is rewritten two barriers from now) -----
    raster_owned(s_scr[buf], s_tri, nt, s_rec[tid >> 5], a.keys + (size_t)f * npix, a.km, a.h, a.w, tid, a.inline_max, a.group, a.group_min);
  }
}

// Resolve pass of the fused path: 4 pixels per thread; winner -> its corners (by ORIGINAL triangle index) ->
// three packed vertex colours -> flat colour (mesh_core.cpp:219); every pixel is written (uncovered or stale-epoch
// key -> 0), so neither the image nor the z-buffer needs a clear.  Requires (h*w) % 4 == 0.
__global__ void __launch_bounds__(256)
resolve_vcol_kernel(const unsigned long long* __restrict__ keys, EpochKey km, const uint32_t* __restrict__ vcol,
                    unsigned vcol_stride, const int4* __restrict__ tri_by_orig, unsigned char* __restrict__ image,
                    unsigned char* __restrict__ mask, size_t npix) {
  const int frame = blockIdx.y;
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 pixels
  if (q * 4 >= npix) return;
  const size_t base = (size_t)frame * npix + q * 4;
  const ulonglong2 k01 = __ldcs(reinterpret_cast<const ulonglong2*>(keys + base));
  const ulonglong2 k23 = __ldcs(reinterpret_cast<const ulonglong2*>(keys + base + 2));
  const unsigned long long k[4] = {k01.x, k01.y, k23.x, k23.y};
  const uint32_t* vc = vcol + (size_t)frame * vcol_stride;
  int4 t4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t t = km.tri_mask - (static_cast<uint32_t>(k[i]) & km.tri_mask);  // winner's ORIGINAL index
    const bool same = i > 0 && k[i] == k[i - 1];   // same triangle as the pixel to the left: reuse its colour
    t4[i] = same ? make_int4(-2, 0, 0, 0) : ((k[i] >= km.epoch_field) ? __ldg(tri_by_orig + t) : make_int4(-1, 0, 0, 0));
  }
  uint32_t col[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (t4[i].x == -2)
      col[i] = col[i - 1];
    else
      col[i] = (t4[i].x >= 0) ? flat_color_packed(__ldg(vc + t4[i].x), __ldg(vc + t4[i].y), __ldg(vc + t4[i].z)) : 0u;
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

// Measured on B200 (profiles/r02d_paths.txt): the separate kernels win at every BASELINE resolution (the fused
// kernel runs its two phases back to back in the same warps at 20 warps/SM, the separate kernels overlap across
// the two chunk streams), so the fused kernel is taken only on request (vp_set_raster_path(m, 2)).
bool fused_available(const vp_model* m) {
  return m->fused_ok && m->fused_mode == 2 && m->ntiles > 0 && m->vertex_mode == 0;
}

int launch_fused(vp_model* m, const float* disp_dev, int nframes, const void* frame_constants, uint32_t* vcol,
                 unsigned long long* keys, uint32_t epoch, int res, cudaStream_t st) {
  if (nframes == 0 || m->ntiles == 0) return VP_OK;
  FusedArgs a;
  VertexArgs& v = a.v;
  v = VertexArgs();
  v.tiles = m->tiles;
  v.tile_list = m->tile_list;
  v.fan = m->fan;
  v.halo = m->halo;
  v.base = m->base;
  v.tex = m->have_tex ? m->tex : nullptr;
  v.disp = disp_dev;
  v.disp_stride = (size_t)m->rows_pad;
  v.fshared = static_cast<const FrameConst*>(frame_constants);
  v.nframes = nframes;
  v.nver = m->nver;
  a.own_tri_off = m->own_tri_off;
  a.own_ltri = m->own_ltri;
  a.tri = m->tri;
  a.vcol = vcol;
  a.vcol_stride = (unsigned)m->vrec_stride;
  a.keys = keys;
  a.km = make_epoch_key(m->ntri, epoch);
  a.h = res;
  a.w = res;
  a.inline_max = inline_box_pixels();
  a.group = walk_group_lanes();
  a.group_min = walk_group_min_pixels();
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
  const int minb = 5, waves = 2;   // measured best of 4 / 5 / 6 CTAs per SM and 1..4 waves (profiles/r02c_fused_ablation.txt)
  // a CTA takes a run of frames (tile constants and the pipeline prologue are paid once per CTA); runs are sized
  // for about `waves` resident waves of CTAs
  const int groups = std::max(1, std::min(nframes, (sms * minb * waves) / std::max(m->ntiles, 1)));
  v.frames_per_block = (nframes + groups - 1) / groups;
  dim3 grid(m->ntiles, (nframes + v.frames_per_block - 1) / v.frames_per_block);
  fused_tile_kernel<minb><<<grid, kTileV, 0, st>>>(a);
  VP_LAUNCH_CHECK();
  return VP_OK;
}

int launch_resolve_vcol(const vp_model* m, const unsigned long long* keys, const uint32_t* vcol, uint32_t epoch,
                        unsigned char* image, unsigned char* mask, int nframes, int h, int w, cudaStream_t st) {
  const size_t npix = (size_t)h * w;
  if (nframes == 0 || npix == 0) return VP_OK;
  dim3 grid((unsigned)((npix / 4 + 255) / 256), nframes);
  resolve_vcol_kernel<<<grid, 256, 0, st>>>(keys, make_epoch_key(m->ntri, epoch), vcol, (unsigned)m->vrec_stride,
                                            m->tri_by_orig, image, mask, npix);
  VP_LAUNCH_CHECK();
  return VP_OK;
}

}  // namespace vp
