// Reconstruction kernels: the per-clip identity contraction (K0), the per-frame expression
// basis contraction (K1, FP32 SIMT flavour; the tcgen05 flavour lives in basis_tc.cu) and the
// fused vertex kernel (K2: normals + rotation + SH illumination + projection).
// Reference: utils/reconstruct_mesh.py:20-29 (Shape_formation), :58-62 (Texture_formation),
// :35-52 (Compute_norm), :100-120 (Projection_layer), :129-168 (Illumination_layer),
// :172-223 (Reconstruction / Reconstruction_rotation).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "launch.h"
#include "ptx.cuh"
#include "vertex.cuh"

namespace vp {

// =========================================================================================
// K0: out[r] = mean[r] + sum_k basis[r][k] * coeff[k]  (- center[r % 3]), float64 accumulate.
// Runs once per clip ("identity mean precomputed once").
// =========================================================================================
// Four lanes share a basis row and read it as 16-byte vectors (lane l takes vectors l, l + 4, ...), so a
// warp streams 8 consecutive rows as one contiguous, fully coalesced span; float64 accumulation, two
// shuffles to combine the four partial sums.
template <typename B, typename O, int K>
__global__ void __launch_bounds__(256)
identity_kernel(const B* __restrict__ basis, const double* __restrict__ mean, const float* __restrict__ coeff,
                double c0, double c1, double c2, O* __restrict__ out, int rows) {
  constexpr int kVec = 16 / sizeof(B);       // elements per 16-byte vector
  constexpr int kVecsPerRow = K / kVec;      // 20 (float) or 40 (double)
  static_assert(K % (4 * kVec) == 0, "row must split into 16-byte vectors over four lanes");
  __shared__ double cs[K];
  if (threadIdx.x < K) cs[threadIdx.x] = (double)coeff[threadIdx.x];
  __syncthreads();
  const int sub = threadIdx.x & 3;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  double acc = 0.0;
  if (r < rows) {
    const uint4* row = reinterpret_cast<const uint4*>(basis + (size_t)r * K);
#pragma unroll
    for (int j = 0; j < kVecsPerRow / 4; ++j) {
      const int v = sub + 4 * j;
      const uint4 raw = __ldcs(row + v);   // read once per clip: streaming
      B e[kVec];
      memcpy(e, &raw, 16);
#pragma unroll
      for (int i = 0; i < kVec; ++i) acc += (double)e[i] * cs[v * kVec + i];
    }
  }
  acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
  acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 2);
  if (r < rows && sub == 0) {
    acc += mean[r];
    const int axis = r % 3;
    acc -= (axis == 0) ? c0 : (axis == 1 ? c1 : c2);
    out[r] = static_cast<O>(acc);
  }
}

int launch_identity(vp_model* m, const float* id_dev, const float* tex_dev, cudaStream_t st) {
  const int grid = (4 * m->rows + 255) / 256;
  if (id_dev) {
    if (m->idb64)
      identity_kernel<double, double, VP_N_ID><<<grid, 256, 0, st>>>(
          static_cast<const double*>(m->idb), m->meanshape, id_dev, m->center[0], m->center[1], m->center[2], m->base,
          m->rows);
    else
      identity_kernel<float, double, VP_N_ID><<<grid, 256, 0, st>>>(
          static_cast<const float*>(m->idb), m->meanshape, id_dev, m->center[0], m->center[1], m->center[2], m->base,
          m->rows);
    VP_LAUNCH_CHECK();
  }
  if (tex_dev) {
    if (m->texb64)
      identity_kernel<double, float, VP_N_TEX><<<grid, 256, 0, st>>>(static_cast<const double*>(m->texb), m->meantex,
                                                                     tex_dev, 0.0, 0.0, 0.0, m->tex, m->rows);
    else
      identity_kernel<float, float, VP_N_TEX><<<grid, 256, 0, st>>>(static_cast<const float*>(m->texb), m->meantex,
                                                                    tex_dev, 0.0, 0.0, 0.0, m->tex, m->rows);
    VP_LAUNCH_CHECK();
  }
  return VP_OK;
}

// =========================================================================================
// K1 (SIMT): disp[t][r] = sum_k exb[r][k] * ex[t][k], FP32.
// One CTA owns 128 basis rows: every thread TMA-bulk-copies its own 256-byte row into a padded
// shared-memory tile (row pitch 272 B, so the 128-bit reads below are bank-conflict free),
// keeps the 64 coefficients of its row in registers and streams over the frames; the frame
// coefficients are broadcast reads from shared memory.  The basis is read from HBM exactly once.
// =========================================================================================
constexpr int kBasisRows = 128;
constexpr int kBasisFrames = 32;  // frames staged per pass
constexpr int kBasisPitch = 68;   // floats

__global__ void __launch_bounds__(kBasisRows)
basis_simt_kernel(const float* __restrict__ exb, const float* __restrict__ ex, float* __restrict__ disp, int nframes,
                  int rows_pad) {
  __shared__ __align__(16) float a_s[kBasisRows * kBasisPitch];
  __shared__ __align__(16) float ex_s[kBasisFrames * VP_N_EX];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  const int row = blockIdx.x * kBasisRows + tid;
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) ptx::mbar_arrive_expect_tx(&bar, kBasisRows * VP_N_EX * sizeof(float));
  ptx::bulk_g2s(a_s + tid * kBasisPitch, exb + (size_t)row * VP_N_EX, VP_N_EX * sizeof(float), &bar);
  ptx::mbar_wait(&bar, 0);

  float a[VP_N_EX];
  {
    const float4* p = reinterpret_cast<const float4*>(a_s + tid * kBasisPitch);
#pragma unroll
    for (int j = 0; j < VP_N_EX / 4; ++j) {
      const float4 v = p[j];
      a[4 * j + 0] = v.x;
      a[4 * j + 1] = v.y;
      a[4 * j + 2] = v.z;
      a[4 * j + 3] = v.w;
    }
  }

  for (int t0 = 0; t0 < nframes; t0 += kBasisFrames) {
    const int nt = min(kBasisFrames, nframes - t0);
    __syncthreads();
    {
      const float4* src = reinterpret_cast<const float4*>(ex + (size_t)t0 * VP_N_EX);
      float4* dst = reinterpret_cast<float4*>(ex_s);
      for (int i = tid; i < kBasisFrames * VP_N_EX / 4; i += kBasisRows)
        dst[i] = (i < nt * (VP_N_EX / 4)) ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (int t = 0; t < nt; t += 4) {  // 4 frames in flight per thread (ex_s is zero padded)
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < VP_N_EX / 4; ++j) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 e = *reinterpret_cast<const float4*>(ex_s + (t + u) * VP_N_EX + 4 * j);
          acc[u] = fmaf(a[4 * j + 0], e.x, acc[u]);
          acc[u] = fmaf(a[4 * j + 1], e.y, acc[u]);
          acc[u] = fmaf(a[4 * j + 2], e.z, acc[u]);
          acc[u] = fmaf(a[4 * j + 3], e.w, acc[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (t + u < nt) disp[(size_t)(t0 + t + u) * rows_pad + row] = acc[u];
    }
  }
}

// Same contraction, the basis tile landed by TWO TMA tensor loads per CTA (the 2-D descriptor of the tcgen05 kernel:
// box = 32 floats x 128 rows, 128-byte swizzle) instead of 128 bulk copies of 256 bytes: the per-SM TMA unit was the
// limiter of the GEMV-sized launches (ncu round 2: 17.6 us for ONE frame = 1.5 TB/s, every CTA queueing 128 tiny
// copies).  The swizzle also makes the row-per-thread reads conflict free without padding: 16-byte chunk c of row r
// of a K half sits at (r / 8) * 1024 + (r % 8) * 128 + ((c ^ (r % 8)) * 16), so the 8 lanes of a quarter-warp hit 8
// different bank groups.
__global__ void __launch_bounds__(kBasisRows)
basis_simt_tma_kernel(const __grid_constant__ CUtensorMap tmap_a, const float* __restrict__ ex, float* __restrict__ disp,
                      int nframes, int rows_pad) {
  __shared__ __align__(1024) uint8_t a_s[kBasisRows * VP_N_EX * sizeof(float)];  // two K halves of 16 KB
  __shared__ __align__(16) float ex_s[kBasisFrames * VP_N_EX];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  const int row = blockIdx.x * kBasisRows + tid;
  constexpr int kHalf = kBasisRows * 32 * sizeof(float);
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
    ptx::mbar_arrive_expect_tx(&bar, 2 * kHalf);
    ptx::tma_load_2d(a_s, &tmap_a, 0, blockIdx.x * kBasisRows, &bar);
    ptx::tma_load_2d(a_s + kHalf, &tmap_a, 32, blockIdx.x * kBasisRows, &bar);
  }
  __syncthreads();
  ptx::mbar_wait(&bar, 0);

  float a[VP_N_EX];
  {
    const uint8_t* base = a_s + (tid >> 3) * 1024 + (tid & 7) * 128;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(base + h * kHalf + ((c ^ (tid & 7)) << 4));
        a[32 * h + 4 * c + 0] = v.x;
        a[32 * h + 4 * c + 1] = v.y;
        a[32 * h + 4 * c + 2] = v.z;
        a[32 * h + 4 * c + 3] = v.w;
      }
  }

  for (int t0 = 0; t0 < nframes; t0 += kBasisFrames) {
    const int nt = min(kBasisFrames, nframes - t0);
    __syncthreads();
    {
      const float4* src = reinterpret_cast<const float4*>(ex + (size_t)t0 * VP_N_EX);
      float4* dst = reinterpret_cast<float4*>(ex_s);
      for (int i = tid; i < kBasisFrames * VP_N_EX / 4; i += kBasisRows)
        dst[i] = (i < nt * (VP_N_EX / 4)) ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (int t = 0; t < nt; t += 4) {  // 4 frames in flight per thread (ex_s is zero padded)
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < VP_N_EX / 4; ++j) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 e = *reinterpret_cast<const float4*>(ex_s + (t + u) * VP_N_EX + 4 * j);
          acc[u] = fmaf(a[4 * j + 0], e.x, acc[u]);
          acc[u] = fmaf(a[4 * j + 1], e.y, acc[u]);
          acc[u] = fmaf(a[4 * j + 2], e.z, acc[u]);
          acc[u] = fmaf(a[4 * j + 3], e.w, acc[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (t + u < nt) disp[(size_t)(t0 + t + u) * rows_pad + row] = acc[u];
    }
  }
}

int launch_basis_simt(vp_model* m, const float* ex_dev, float* disp_dev, int nframes, cudaStream_t st) {
  if (nframes == 0) return VP_OK;
  if (m->have_tmap) {
    CUtensorMap map;
    static_assert(sizeof(map) == sizeof(m->tmap_exb), "tensor map storage size");
    std::memcpy(&map, m->tmap_exb, sizeof(map));
    basis_simt_tma_kernel<<<m->rows_pad / kBasisRows, kBasisRows, 0, st>>>(map, ex_dev, disp_dev, nframes, m->rows_pad);
  } else {  // no TMA descriptor (driver without cuTensorMapEncodeTiled): per-row bulk copies
    basis_simt_kernel<<<m->rows_pad / kBasisRows, kBasisRows, 0, st>>>(m->exb, ex_dev, disp_dev, nframes, m->rows_pad);
  }
  VP_LAUNCH_CHECK();
  return VP_OK;
}

// FP32 streamed kernel for GEMV-like batches, tcgen05 3xTF32 GEMM once the frame batch is a real
// dense contraction (BASELINE.json north_star); vp_set_basis_mode overrides the choice.
int launch_basis(vp_model* m, const float* ex_dev, float* disp_dev, int nframes, cudaStream_t st) {
  bool tensor = m->have_tmap && nframes >= kBasisTensorMinFrames;
  if (m->basis_mode == kBasisSimt) tensor = false;
  if (m->basis_mode == kBasisTensor) {
    VP_REQUIRE(m->have_tmap, "tensor-core basis kernel unavailable (no TMA descriptor)");
    tensor = true;
  }
  return tensor ? launch_basis_tc(m, ex_dev, disp_dev, nframes, st) : launch_basis_simt(m, ex_dev, disp_dev, nframes, st);
}

// =========================================================================================
// K2: fused vertex stage, one CTA per (vertex tile, run of frames).  Local vertex positions (own +
// halo) are staged in shared memory per frame; the summed face normal of every own vertex comes either
// from its FAN record (vertex_fan_kernel: 9 position gathers, any manifold mesh) or from a pass over
// the tile's triangles plus a ring gather in point_buf slot order (vertex_tile_kernel: any point_buf);
// finish_vertex then normalises, rotates, lights and projects.  ncu (profiles/r01d): the stage is bound by
// the LSU data pipe (shared-memory wavefronts of the 16-byte gathers, 81 % of peak), not by HBM, which
// is why the fan path halves the gathers instead of the bytes.
// =========================================================================================
// (per-frame constants, LocalVerts and the fan normal sum: vertex.cuh)

// Per-frame constants, once per frame instead of once per (tile, frame): the rotation as float32 and
// the SH coefficients with the band constants folded in (Illumination_layer, reconstruct_mesh.py:133-153:
// Y_k = K_k * b_k(n), lit_c = sum_k Y_k gamma'_ck with gamma' = gamma + 0.8 on band 0; K_k are products
// of a0..a2 and c0..c2 evaluated in float64).
__global__ void frame_prep_kernel(const FrameParams* __restrict__ params, FrameConst* __restrict__ out_all, int nframes,
                                  int rotate_first, double focal, double center, double image_size, double scale) {
  const int f = blockIdx.x * (blockDim.x / 64) + threadIdx.x / 64;
  const int t = threadIdx.x % 64;
  if (f >= nframes) return;
  const FrameParams& p = params[f];
  FrameShared& o = out_all[f].slow;
  if (t < 48) reinterpret_cast<uint32_t*>(&o.par)[t] = reinterpret_cast<const uint32_t*>(&p)[t];
  if (t < 12) o.rot[t] = t < 9 ? (float)p.rot[t] : 0.f;
  auto band_constant = [](int k) {
    return (k == 0) ? 0.8862269254527579
                    : (k <= 3 ? 1.772453850905516
                              : (k == 6 ? 0.7006239020497412 : (k == 8 ? 1.2135161953473121 : 2.4270323906946243)));
  };
  auto folded = [&](int c, int k) {  // sign * K_k * (gamma_ck + 0.8 [k == 0])
    const double sign = (k == 1 || k == 3 || k == 5 || k == 7) ? -1.0 : 1.0;
    return sign * band_constant(k) * (double)(p.gamma[9 * c + k] + (k == 0 ? 0.8f : 0.f));  // float32 add, like numpy's in-place += 0.8
  };
  if (t >= 32 && t < 60) {
    const int i = t - 32;
    if (i >= 27) {
      o.sh[i] = 0.f;
    } else {
      const int k = i % 9;
      const float kk = (float)band_constant(k);
      const float sign = (k == 1 || k == 3 || k == 5 || k == 7) ? -1.f : 1.f;
      o.sh[i] = sign * kk * (p.gamma[i] + (k == 0 ? 0.8f : 0.f));
    }
  }
  FrameFast& q = out_all[f].fast;
  if (t == 60) {  // the projective map, float64
    const double* R = p.rot;
    double M[9];
    if (rotate_first) {
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[3 * i + j] = R[3 * i] * R[j] + R[3 * i + 1] * R[3 + j] + R[3 * i + 2] * R[6 + j];
    } else {
      for (int i = 0; i < 9; ++i) M[i] = R[i];
    }
    const double t0 = (double)p.trans[0], t1 = (double)p.trans[1], t2 = (double)p.trans[2];
    const double bz = 10.0 - t2;
    const double bx = focal * t0 + center * bz, by = focal * t1 + center * bz;
    for (int i = 0; i < 3; ++i) {
      const double az = -M[3 * i + 2];
      const double ax = focal * M[3 * i] + center * az, ay = focal * M[3 * i + 1] + center * az;
      q.lin[i] = scale * ax;
      q.lin[4 + i] = scale * (image_size * az - ay);
      q.lin[8 + i] = az;
    }
    q.lin[3] = scale * bx;
    q.lin[7] = scale * (image_size * bz - by);
    q.lin[11] = bz;
  }
  if (t >= 61 && t < 64) {  // one colour channel each: the quadratic form of the unrotated normal
    const int c = t - 61;
    const double* R = p.rot;  // n_r = n R  (row vector), so b = R b_r and Q = R Q_r R'
    double g[9];
    for (int k = 0; k < 9; ++k) g[k] = folded(c, k);
    const double br[3] = {g[3], g[1], g[2]};
    const double Qr[9] = {g[8], 0.5 * g[4], 0.5 * g[7], 0.5 * g[4], -g[8], 0.5 * g[5], 0.5 * g[7], 0.5 * g[5], 3.0 * g[6]};
    double b[3], RQ[9], Q[9];
    for (int i = 0; i < 3; ++i) b[i] = R[3 * i] * br[0] + R[3 * i + 1] * br[1] + R[3 * i + 2] * br[2];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) RQ[3 * i + j] = R[3 * i] * Qr[j] + R[3 * i + 1] * Qr[3 + j] + R[3 * i + 2] * Qr[6 + j];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Q[3 * i + j] = RQ[3 * i] * R[3 * j] + RQ[3 * i + 1] * R[3 * j + 1] + RQ[3 * i + 2] * R[3 * j + 2];
    float* o10 = q.shq + 10 * c;
    o10[0] = (float)(g[0] - g[6]);
    o10[1] = (float)b[0];
    o10[2] = (float)b[1];
    o10[3] = (float)b[2];
    o10[4] = (float)Q[0];
    o10[5] = (float)Q[4];
    o10[6] = (float)Q[8];
    o10[7] = (float)(2.0 * Q[1]);
    o10[8] = (float)(2.0 * Q[2]);
    o10[9] = (float)(2.0 * Q[5]);
    if (c == 0) q.shq[30] = q.shq[31] = 0.f;
  }
}

// The raster-record-only finish (see FrameFast; light_fast / project_fast in vertex.cuh).
__device__ __forceinline__ void finish_vertex_fast(const VertexArgs& a, const FrameFast& ff, int f, int gv0, float nx,
                                                   float ny, float nz, float tr, float tg, float tb, double vx,
                                                   double vy, double vz) {
  const uint32_t rgba = light_fast(ff, nx, ny, nz, tr, tg, tb);
  const float3 p = project_fast(ff, vx, vy, vz);
  a.vrec[(size_t)f * a.vrec_stride + gv0] = make_float4(p.x, p.y, p.z, __uint_as_float(rgba));
}

// What both vertex kernels do once the summed face normal (nx, ny, nz) of the own vertex is known:
// normalise, rotate, 9-band SH lighting, colour (float32); position, rotation(s), perspective projection
// (float64); one float4 raster record and/or the reference's per-vertex outputs.
__device__ __forceinline__ void finish_vertex(const VertexArgs& a, const FrameShared& fs, int f, int gv0, int orig,
                                              float nx, float ny, float nz, float tr, float tg, float tb, double sx,
                                              double sy, double sz) {
  {
    const float inv = rsqrtf(nx * nx + ny * ny + nz * nz);  // 0 * inf -> NaN for a vertex without faces
    nx *= inv;
    ny *= inv;
    nz *= inv;
  }
  // rotated normal (reconstruct_mesh.py:184 / :208), lighting and colour in float32
  const float* rf = fs.rot;
  const float nrx = nx * rf[0] + ny * rf[3] + nz * rf[6];
  const float nry = nx * rf[1] + ny * rf[4] + nz * rf[7];
  const float nrz = nx * rf[2] + ny * rf[5] + nz * rf[8];
  float lit[3];
  {
    const float b4 = nrx * nry, b5 = nry * nrz, b6 = 3.f * nrz * nrz - 1.f, b7 = nrx * nrz,
                b8 = nrx * nrx - nry * nry;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* g = fs.sh + 9 * c;
      lit[c] = g[0] + g[1] * nry + g[2] * nrz + g[3] * nrx + g[4] * b4 + g[5] * b5 + g[6] * b6 + g[7] * b7 + g[8] * b8;
    }
  }
  const float cr = lit[0] * tr, cg = lit[1] * tg, cb = lit[2] * tb;

  // geometry in float64
  const double* R = fs.par.rot;
  if (a.rotate_first) {  // Reconstruction_rotation rotates the shape before projecting it (:211)
    double tx, ty, tz;
    rotate_row(R, sx, sy, sz, tx, ty, tz);
    sx = tx;
    sy = ty;
    sz = tz;
  }
  double px, py, zb;
  project(R, fs.par.trans, a.focal, a.center, sx, sy, sz, px, py, zb);
  const double pyf = a.image_size - py;  // reconstruct_mesh.py:187 / :215

  if (a.vrec) {
    // infer_bfmvid.py:93-105: (x, S - y, z_buffer) -> float32; colours clipped and truncated
    const uint32_t rgba = clip_trunc_byte(cr) | (clip_trunc_byte(cg) << 8) | (clip_trunc_byte(cb) << 16);
    a.vrec[(size_t)f * a.vrec_stride + gv0] =
        make_float4((float)(px * a.raster_scale), (float)(pyf * a.raster_scale), (float)zb, __uint_as_float(rgba));
  }
  if (a.has_out) {
    const size_t o = (size_t)f * a.nver + orig;
    if (a.out.shape) {
      a.out.shape[3 * o] = sx;
      a.out.shape[3 * o + 1] = sy;
      a.out.shape[3 * o + 2] = sz;
    }
    if (a.out.norm) {
      a.out.norm[3 * o] = nx;
      a.out.norm[3 * o + 1] = ny;
      a.out.norm[3 * o + 2] = nz;
    }
    if (a.out.color) {
      a.out.color[3 * o] = cr;
      a.out.color[3 * o + 1] = cg;
      a.out.color[3 * o + 2] = cb;
    }
    if (a.out.proj) {
      a.out.proj[2 * o] = px;
      a.out.proj[2 * o + 1] = a.out.flip_y ? pyf : py;
    }
    if (a.out.zbuf) a.out.zbuf[o] = zb;
  }
}

// K2, fan flavour (tiles whose vertices all have fan records: any manifold mesh).  One CTA per (tile,
// run of frames).  Per frame: the own vertex sums (u_i - v) x (u_i+1 - v) over its ring from 9 gathers of
// staged positions, finishes (finish_vertex), stages the next frame's positions into the other buffer,
// and the block synchronises once.  Shared memory holds positions only (no per-triangle pass).
// SLOTS: local vertex i is staged at shared-memory slot slot_tab[i] instead of i and the fan records come in slot
// space (fan_slot), which roughly halves the bank conflicts of the gathers in the quarter-warp model
// (tools/bank_conflict_sim.py); measured on B200: 56.3 -> 51.8 us per 75 frames at 256x256, so it is the default
// (vp_set_vertex_mode(m, 2) selects the identity placement, for the tests).
template <int MIN_BLOCKS, bool FAST, bool SLOTS = false>
__global__ void __launch_bounds__(kTileV, MIN_BLOCKS) vertex_fan_kernel(const VertexArgs a) {
  using Frame = typename std::conditional<FAST, FrameFast, FrameShared>::type;
  __shared__ float4 s_pos[2][kTileLV + (SLOTS ? 8 : 0)];
  __shared__ __align__(16) Frame s_frame[2];

  const int tile_id = __ldg(a.tile_list + blockIdx.x);
  const TileDesc td = a.tiles[tile_id];
  const int tid = threadIdx.x;
  const int nq_v = (td.nlv + kTileV - 1) / kTileV;  // CTA-uniform
  const int f_begin = blockIdx.y * a.frames_per_block;
  const int f_end = min(a.nframes, f_begin + a.frames_per_block);
  if (f_begin >= f_end) return;

  LocalVerts lv;
  lv.load(a, td, tid);
  const bool own = tid < td.nv;
  uint32_t fan[kFanWords] = {0, 0, 0, 0, 0};
  float tr = 0.f, tg = 0.f, tb = 0.f;
  uint32_t slot_offs = 0;  // SLOTS: byte offsets of this thread's two local vertices, 16 bits each
  if constexpr (SLOTS) {
    const int so = __ldg(a.slot_off + tile_id);
#pragma unroll
    for (int q = 0; q < kSlotsV; ++q) {
      const int i = tid + q * kTileV;
      // entries past nlv are still staged (CTA-uniform trip count): park them in a spare slot, not in slot 0
      const uint32_t slot = (i < td.nlv) ? (uint32_t)__ldg(a.slot_tab + so + i) : (uint32_t)kTileLV;
      slot_offs |= (slot << 4) << (16 * q);
    }
  }
  if (own) {
    const uint32_t* fan_tab = SLOTS ? a.fan_slot : a.fan;
#pragma unroll
    for (int k = 0; k < kFanWords; ++k) fan[k] = __ldg(fan_tab + (size_t)lv.gv[0] * kFanWords + k);
    if (a.tex) {
      tr = __ldg(a.tex + 3 * (size_t)lv.gv[0]);
      tg = __ldg(a.tex + 3 * (size_t)lv.gv[0] + 1);
      tb = __ldg(a.tex + 3 * (size_t)lv.gv[0] + 2);
    }
  }

  // prologue: frame f_begin staged, displacement of frame f_begin + 1 in flight
  lv.fetch(a, f_begin, nq_v);
  stage_frame_constants(a, &s_frame[0], f_begin, tid);
  if constexpr (SLOTS)
    lv.stage_at(reinterpret_cast<char*>(s_pos[0]), slot_offs, nq_v);
  else
    lv.stage(s_pos[0], tid, nq_v);
  float3 pv = lv.own_staged();                           // == the own vertex's staged position, without the shared-memory read
  float d0x = lv.dx[0], d0y = lv.dy[0], d0z = lv.dz[0];  // own displacement of the frame being finished
  if (f_begin + 1 < f_end) lv.fetch(a, f_begin + 1, nq_v);
  __syncthreads();

  for (int f = f_begin; f < f_end; ++f) {
    const int buf = (f - f_begin) & 1;
    if (own) {
      float nx, ny, nz;
      fan_normal_sum(reinterpret_cast<const char*>(s_pos[buf]), fan, pv, nx, ny, nz);
      if constexpr (FAST) {
        finish_vertex_fast(a, s_frame[buf], f, lv.gv[0], nx, ny, nz, tr, tg, tb, lv.bx + (double)d0x,
                           lv.by + (double)d0y, lv.bz + (double)d0z);
      } else {
        int orig = 0;
        if (a.has_out) orig = __ldg(a.v_int2orig + lv.gv[0]);
        finish_vertex(a, s_frame[buf], f, lv.gv[0], orig, nx, ny, nz, tr, tg, tb, lv.bx + (double)d0x,
                      lv.by + (double)d0y, lv.bz + (double)d0z);
      }
    }
    // ---- stage frame f + 1 into the other buffers (their readers passed the previous barrier) ----
    if (f + 1 < f_end) {
      stage_frame_constants(a, &s_frame[buf ^ 1], f + 1, tid);
      if constexpr (SLOTS)
        lv.stage_at(reinterpret_cast<char*>(s_pos[buf ^ 1]), slot_offs, nq_v);
      else
        lv.stage(s_pos[buf ^ 1], tid, nq_v);
      pv = lv.own_staged();
      d0x = lv.dx[0];
      d0y = lv.dy[0];
      d0z = lv.dz[0];
      if (f + 2 < f_end) lv.fetch(a, f + 2, nq_v);
    }
    __syncthreads();
  }
}

// K2, generic flavour (tiles where point_buf does not chain into fans: non-manifold meshes, faces listed
// for vertices they do not contain).  Per frame: triangle normals of the tile -> shared memory, block
// barrier, then per own vertex the ring sum in point_buf slot order (pad slots read a zero entry, like the
// zero row the reference appends), finish_vertex, stage the next frame, block barrier.
__global__ void __launch_bounds__(kTileV, 5) vertex_tile_kernel(const VertexArgs a) {
  __shared__ float4 s_pos[2][kTileLV];
  __shared__ float4 s_fn[kTileLT + 1];
  __shared__ __align__(16) FrameShared s_frame[2];
  static_assert(sizeof(FrameParams) == 192, "FrameParams layout");

  const TileDesc td = a.tiles[__ldg(a.tile_list + blockIdx.x)];
  const int tid = threadIdx.x;
  const int nq_v = (td.nlv + kTileV - 1) / kTileV;  // CTA-uniform trip counts
  const int nq_t = (td.nlt + kTileV - 1) / kTileV;
  const int f_begin = blockIdx.y * a.frames_per_block;
  const int f_end = min(a.nframes, f_begin + a.frames_per_block);
  if (f_begin >= f_end) return;

  LocalVerts lv;
  lv.load(a, td, tid);
  uint32_t lt[kSlotsT];
#pragma unroll
  for (int q = 0; q < kSlotsT; ++q) {
    const int j = tid + q * kTileV;
    lt[q] = (j < td.nlt) ? __ldg(a.ltri + td.ltri_off + j) : 0u;
  }
  const bool own = tid < td.nv;
  uint4 rg = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
  float tr = 0.f, tg = 0.f, tb = 0.f;
  if (own) {
    rg = __ldg(reinterpret_cast<const uint4*>(a.ring) + lv.gv[0]);
    if (a.tex) {
      tr = __ldg(a.tex + 3 * (size_t)lv.gv[0]);
      tg = __ldg(a.tex + 3 * (size_t)lv.gv[0] + 1);
      tb = __ldg(a.tex + 3 * (size_t)lv.gv[0] + 2);
    }
  }
  if (tid == 0) s_fn[kTileLT] = make_float4(0.f, 0.f, 0.f, 0.f);  // what pad slots of the ring read

  lv.fetch(a, f_begin, nq_v);
  stage_frame_constants(a, &s_frame[0], f_begin, tid);
  lv.stage(s_pos[0], tid, nq_v);
  float d0x = lv.dx[0], d0y = lv.dy[0], d0z = lv.dz[0];
  if (f_begin + 1 < f_end) lv.fetch(a, f_begin + 1, nq_v);
  __syncthreads();

  for (int f = f_begin; f < f_end; ++f) {
    const int buf = (f - f_begin) & 1;
    // ---- triangle normals (reconstruct_mesh.py:41-46) --------------------------------------
#pragma unroll
    for (int q = 0; q < kSlotsT; ++q)
      if (q < nq_t) {
        const float4 p1 = s_pos[buf][lt[q] & 1023u], p2 = s_pos[buf][(lt[q] >> 10) & 1023u],
                     p3 = s_pos[buf][(lt[q] >> 20) & 1023u];
        const float e1x = p1.x - p2.x, e1y = p1.y - p2.y, e1z = p1.z - p2.z;
        const float e2x = p2.x - p3.x, e2y = p2.y - p3.y, e2z = p2.z - p3.z;
        s_fn[tid + q * kTileV] = make_float4(e1y * e2z - e1z * e2y, e1z * e2x - e1x * e2z, e1x * e2y - e1y * e2x, 0.f);
      }
    __syncthreads();
    if (own) {
      float nx = 0.f, ny = 0.f, nz = 0.f;
      const uint32_t w[4] = {rg.x, rg.y, rg.z, rg.w};
#pragma unroll
      for (int s = 0; s < VP_RING; ++s) {
        const uint32_t j = min((w[s >> 1] >> ((s & 1) * 16)) & 0xFFFFu, (uint32_t)kTileLT);
        const float4 fn = s_fn[j];
        nx += fn.x;
        ny += fn.y;
        nz += fn.z;
      }
      int orig = 0;
      if (a.has_out) orig = __ldg(a.v_int2orig + lv.gv[0]);
      finish_vertex(a, s_frame[buf], f, lv.gv[0], orig, nx, ny, nz, tr, tg, tb, lv.bx + (double)d0x,
                    lv.by + (double)d0y, lv.bz + (double)d0z);
    }
    if (f + 1 < f_end) {
      stage_frame_constants(a, &s_frame[buf ^ 1], f + 1, tid);
      lv.stage(s_pos[buf ^ 1], tid, nq_v);
      d0x = lv.dx[0];
      d0y = lv.dy[0];
      d0z = lv.dz[0];
      if (f + 2 < f_end) lv.fetch(a, f + 2, nq_v);
    }
    __syncthreads();
  }
}

// Per-frame constants of `nframes` frames into m->ws_fshared (one launch per sequence); the returned pointer,
// advanced by frame_constants_stride() per frame, is what launch_vertex takes as `frame_constants`.
int prepare_frame_constants(vp_model* m, const FrameParams* params_dev, int nframes, int rotate_first, double focal,
                            double center, double image_size, double raster_scale, cudaStream_t st, const void** out) {
  *out = nullptr;
  if (nframes == 0) return VP_OK;
  VP_CUDA(m->ws_fshared.reserve((size_t)nframes * sizeof(FrameConst), m->device));
  FrameConst* fc = m->ws_fshared.as<FrameConst>();
  frame_prep_kernel<<<(nframes + 3) / 4, 256, 0, st>>>(params_dev, fc, nframes, rotate_first, focal, center, image_size,
                                                      raster_scale);
  VP_LAUNCH_CHECK();
  *out = fc;
  return VP_OK;
}
size_t frame_constants_stride() { return sizeof(FrameConst); }

int launch_vertex(vp_model* m, const float* disp_dev, const FrameParams* params_dev, int nframes, int rotate_first,
                  double focal, double center, double image_size, double raster_scale, float4* vrec_dev,
                  const ReconOut& out, cudaStream_t st, const void* frame_constants) {
  if (nframes == 0 || m->ntiles == 0) return VP_OK;
  VertexArgs a;
  a.tiles = m->tiles;
  a.tile_list = m->tile_list;
  a.fan = m->fan;
  a.slot_off = m->slot_off;
  a.slot_tab = m->slot_tab;
  a.fan_slot = m->fan_slot;
  a.ltri = m->ltri;
  a.halo = m->halo;
  a.ring = m->ring;
  a.v_int2orig = m->v_int2orig_dev;
  a.base = m->base;
  a.tex = m->have_tex ? m->tex : nullptr;
  a.disp = disp_dev;
  a.disp_stride = (size_t)m->rows_pad;
  const void* prepared = frame_constants;
  if (prepared == nullptr)
    VP_TRY(prepare_frame_constants(m, params_dev, nframes, rotate_first, focal, center, image_size, raster_scale, st,
                                   &prepared));
  a.fshared = static_cast<const FrameConst*>(prepared);
  a.nframes = nframes;
  a.rotate_first = rotate_first;
  a.has_out = (out.shape || out.norm || out.color || out.proj || out.zbuf) ? 1 : 0;
  a.focal = focal;
  a.center = center;
  a.image_size = image_size;
  a.raster_scale = raster_scale;
  a.vrec = vrec_dev;
  a.vrec_stride = (size_t)m->vrec_stride;
  a.out = out;
  a.nver = m->nver;
  static const int fpb_env = [] { const char* e = std::getenv("VPB200_VERTEX_FPB"); return e ? std::atoi(e) : 0; }();
  static const int minb_env = [] { const char* e = std::getenv("VPB200_VERTEX_MINB"); return e ? std::atoi(e) : 0; }();
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
  // Frames per CTA: the tile constants (dependent global loads) and the pipeline prologue are paid once per
  // CTA, so a CTA takes a run of frames; runs are sized for about `waves` resident waves of CTAs.
  auto frames_per_block = [&](int ntiles, int blocks_per_sm, int waves) {
    if (fpb_env > 0) return fpb_env;
    const int groups = std::max(1, std::min(nframes, (sms * blocks_per_sm * waves) / std::max(ntiles, 1)));
    return (nframes + groups - 1) / groups;
  };
  const int n_fan = (m->vertex_mode == 1) ? 0 : m->n_fan_tiles;
  if (n_fan > 0) {
    const int minb = minb_env > 0 ? minb_env : 8;
    a.frames_per_block = frames_per_block(n_fan, minb, 2);
    dim3 grid(n_fan, (nframes + a.frames_per_block - 1) / a.frames_per_block);
    const bool fast = !a.has_out && a.vrec != nullptr;  // raster records only: the folded constants
    if (fast && m->have_slots && m->vertex_mode != 2) {
      vertex_fan_kernel<8, true, true><<<grid, kTileV, 0, st>>>(a);
    } else if (fast) {
      if (minb <= 6)
        vertex_fan_kernel<6, true><<<grid, kTileV, 0, st>>>(a);
      else if (minb == 7)
        vertex_fan_kernel<7, true><<<grid, kTileV, 0, st>>>(a);
      else
        vertex_fan_kernel<8, true><<<grid, kTileV, 0, st>>>(a);
    } else {
      vertex_fan_kernel<6, false><<<grid, kTileV, 0, st>>>(a);
    }
    VP_LAUNCH_CHECK();
  }
  if (m->ntiles - n_fan > 0) {
    a.tile_list = m->tile_list + n_fan;
    a.frames_per_block = frames_per_block(m->ntiles - n_fan, 5, 2);
    dim3 grid(m->ntiles - n_fan, (nframes + a.frames_per_block - 1) / a.frames_per_block);
    vertex_tile_kernel<<<grid, kTileV, 0, st>>>(a);
    VP_LAUNCH_CHECK();
  }
  return VP_OK;
}

// =========================================================================================
// Illumination_layer on caller-supplied arrays (reconstruct_mesh.py:129-168), float64.
// =========================================================================================
__global__ void illumination_kernel(const double* __restrict__ texture, const double* __restrict__ norm,
                                    const float* __restrict__ gamma, double* __restrict__ color,
                                    double* __restrict__ lighting, int n) {
  __shared__ float g[VP_N_GAMMA];
  if (threadIdx.x < VP_N_GAMMA) g[threadIdx.x] = gamma[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double lit[3];
  sh_lighting<double>(g, norm[3 * (size_t)i], norm[3 * (size_t)i + 1], norm[3 * (size_t)i + 2], lit);
  for (int c = 0; c < 3; ++c) {
    if (color) color[3 * (size_t)i + c] = lit[c] * texture[3 * (size_t)i + c];
    if (lighting) lighting[3 * (size_t)i + c] = lit[c] * 128.0;
  }
}

// Projection_layer on a caller-supplied shape (reconstruct_mesh.py:100-120), float64.
__global__ void projection_kernel(const double* __restrict__ shape, FrameParams par, double focal, double center,
                                  double* __restrict__ proj, double* __restrict__ zbuf, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double px, py, zb;
  project(par.rot, par.trans, focal, center, shape[3 * (size_t)i], shape[3 * (size_t)i + 1], shape[3 * (size_t)i + 2],
          px, py, zb);
  proj[2 * (size_t)i] = px;
  proj[2 * (size_t)i + 1] = py;
  zbuf[i] = zb;
}

}  // namespace vp

using namespace vp;

static int illumination_impl(DevBuf& buf, int device, int n, const double* texture, const double* norm,
                             const float* gamma, double* color, double* lighting) {
  const size_t vb = (size_t)n * 3 * sizeof(double);
  VP_CUDA(buf.reserve(4 * vb + 256, device));
  char* b = buf.as<char>();
  double *d_tex = reinterpret_cast<double*>(b), *d_norm = reinterpret_cast<double*>(b + vb),
         *d_col = reinterpret_cast<double*>(b + 2 * vb), *d_lit = reinterpret_cast<double*>(b + 3 * vb);
  float* d_gamma = reinterpret_cast<float*>(b + 4 * vb);
  if (texture) VP_CUDA(cudaMemcpy(d_tex, texture, vb, cudaMemcpyHostToDevice));
  VP_CUDA(cudaMemcpy(d_norm, norm, vb, cudaMemcpyHostToDevice));
  VP_CUDA(cudaMemcpy(d_gamma, gamma, VP_N_GAMMA * sizeof(float), cudaMemcpyHostToDevice));
  illumination_kernel<<<(n + 255) / 256, 256>>>(d_tex, d_norm, d_gamma, color ? d_col : nullptr,
                                                lighting ? d_lit : nullptr, n);
  VP_LAUNCH_CHECK();
  if (color) VP_CUDA(cudaMemcpy(color, d_col, vb, cudaMemcpyDeviceToHost));
  if (lighting) VP_CUDA(cudaMemcpy(lighting, d_lit, vb, cudaMemcpyDeviceToHost));
  return VP_OK;
}

extern "C" int vp_illumination(int device, int n, const double* texture, const double* norm, const float* gamma,
                               double* color, double* lighting) {
  VP_REQUIRE(n >= 0 && norm && gamma && (color == nullptr || texture != nullptr), "null argument");
  if (n == 0) return VP_OK;
  VP_CUDA(cudaSetDevice(device));
  DevBuf buf;
  const int rc = illumination_impl(buf, device, n, texture, norm, gamma, color, lighting);
  buf.release();
  return rc;
}

static int projection_impl(DevBuf& buf, int device, int n, const double* shape, const double* rotation,
                           const float* translation, double focal, double center, double* projection,
                           double* z_buffer) {
  const size_t vb = (size_t)n * sizeof(double);
  VP_CUDA(buf.reserve(6 * vb, device));
  double* d_shape = buf.as<double>();
  double* d_proj = d_shape + 3 * (size_t)n;
  double* d_z = d_proj + 2 * (size_t)n;
  FrameParams par;
  for (int k = 0; k < 9; ++k) par.rot[k] = rotation[k];
  for (int k = 0; k < 3; ++k) par.trans[k] = translation[k];
  for (int k = 0; k < VP_N_GAMMA; ++k) par.gamma[k] = 0.f;
  VP_CUDA(cudaMemcpy(d_shape, shape, 3 * vb, cudaMemcpyHostToDevice));
  projection_kernel<<<(n + 255) / 256, 256>>>(d_shape, par, focal, center, d_proj, d_z, n);
  VP_LAUNCH_CHECK();
  VP_CUDA(cudaMemcpy(projection, d_proj, 2 * vb, cudaMemcpyDeviceToHost));
  VP_CUDA(cudaMemcpy(z_buffer, d_z, vb, cudaMemcpyDeviceToHost));
  return VP_OK;
}

extern "C" int vp_projection(int device, int n, const double* shape, const double* rotation9,
                             const float* translation3, double focal, double center, double* projection,
                             double* z_buffer) {
  VP_REQUIRE(n >= 0 && shape && rotation9 && translation3 && projection && z_buffer, "null argument");
  if (n == 0) return VP_OK;
  VP_CUDA(cudaSetDevice(device));
  DevBuf buf;
  const int rc = projection_impl(buf, device, n, shape, rotation9, translation3, focal, center, projection, z_buffer);
  buf.release();
  return rc;
}

// Reconstruction / Reconstruction_rotation for `frames->nframes` coefficient rows at once.
extern "C" int vp_reconstruct(vp_model* m, const vp_frames* fr, const vp_recon_out* out) {
  VP_REQUIRE(m != nullptr && fr != nullptr && out != nullptr, "null argument");
  VP_REQUIRE(fr->nframes >= 0, "nframes >= 0");
  VP_REQUIRE(fr->nframes == 0 || (fr->rotation && fr->translation && fr->gamma), "null per-frame array");
  std::lock_guard<std::mutex> lock(m->mu);
  VP_REQUIRE(m->have_base, "no base shape (call vp_set_identity or vp_set_base_shape first)");
  VP_REQUIRE(!out->face_color || m->have_tex, "face_color requested but no texture set");
  VP_CUDA(cudaSetDevice(m->device));
  VP_TRY(wait_for_renders(m));   // shares ws_params / ws_disp / ws_fshared with renders that may still be in flight
  cudaStream_t st = nullptr;
  const int T = fr->nframes;
  const int chunk = 64;
  const size_t nv = (size_t)m->nver;
  for (int t0 = 0; t0 < T; t0 += chunk) {
    const int n = std::min(chunk, T - t0);
    // per-frame parameters
    std::vector<FrameParams> hp((size_t)n);
    for (int i = 0; i < n; ++i) {
      const size_t t = (size_t)t0 + i;
      for (int k = 0; k < 9; ++k) hp[i].rot[k] = fr->rotation[9 * t + k];
      for (int k = 0; k < 3; ++k) hp[i].trans[k] = fr->translation[3 * t + k];
      for (int k = 0; k < VP_N_GAMMA; ++k) hp[i].gamma[k] = fr->gamma[VP_N_GAMMA * t + k];
    }
    VP_CUDA(m->ws_params.reserve((size_t)chunk * sizeof(FrameParams), m->device));
    VP_CUDA(cudaMemcpyAsync(m->ws_params.ptr, hp.data(), (size_t)n * sizeof(FrameParams), cudaMemcpyHostToDevice, st));
    float* disp = nullptr;
    if (fr->ex) {
      VP_CUDA(m->ws_ex.reserve((size_t)chunk * VP_N_EX * sizeof(float), m->device));
      VP_CUDA(m->ws_disp.reserve((size_t)chunk * m->rows_pad * sizeof(float), m->device));
      VP_CUDA(cudaMemcpyAsync(m->ws_ex.ptr, fr->ex + (size_t)t0 * VP_N_EX, (size_t)n * VP_N_EX * sizeof(float),
                              cudaMemcpyHostToDevice, st));
      disp = m->ws_disp.as<float>();
      VP_TRY(launch_basis(m, m->ws_ex.as<float>(), disp, n, st));
    }
    // outputs: carve one scratch allocation
    const size_t b_shape = out->face_shape ? nv * 3 * sizeof(double) * n : 0;
    const size_t b_norm = out->face_norm ? nv * 3 * sizeof(float) * n : 0;
    const size_t b_color = out->face_color ? nv * 3 * sizeof(float) * n : 0;
    const size_t b_proj = out->projection ? nv * 2 * sizeof(double) * n : 0;
    const size_t b_z = out->z_buffer ? nv * sizeof(double) * n : 0;
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    VP_CUDA(m->ws_out.reserve(al(b_shape) + al(b_norm) + al(b_color) + al(b_proj) + al(b_z) + 256, m->device));
    char* p = m->ws_out.as<char>();
    ReconOut ro;
    ro.flip_y = out->flip_y;
    if (b_shape) { ro.shape = reinterpret_cast<double*>(p); p += al(b_shape); }
    if (b_proj) { ro.proj = reinterpret_cast<double*>(p); p += al(b_proj); }
    if (b_z) { ro.zbuf = reinterpret_cast<double*>(p); p += al(b_z); }
    if (b_norm) { ro.norm = reinterpret_cast<float*>(p); p += al(b_norm); }
    if (b_color) { ro.color = reinterpret_cast<float*>(p); p += al(b_color); }
    VP_TRY(launch_vertex(m, disp, m->ws_params.as<FrameParams>(), n, fr->rotate_shape_first, fr->focal, fr->center,
                         out->image_size, 1.0, nullptr, ro, st));
    const size_t off = (size_t)t0 * nv;
    if (b_shape) VP_CUDA(cudaMemcpyAsync(out->face_shape + 3 * off, ro.shape, b_shape, cudaMemcpyDeviceToHost, st));
    if (b_norm) VP_CUDA(cudaMemcpyAsync(out->face_norm + 3 * off, ro.norm, b_norm, cudaMemcpyDeviceToHost, st));
    if (b_color) VP_CUDA(cudaMemcpyAsync(out->face_color + 3 * off, ro.color, b_color, cudaMemcpyDeviceToHost, st));
    if (b_proj) VP_CUDA(cudaMemcpyAsync(out->projection + 2 * off, ro.proj, b_proj, cudaMemcpyDeviceToHost, st));
    if (b_z) VP_CUDA(cudaMemcpyAsync(out->z_buffer + off, ro.zbuf, b_z, cudaMemcpyDeviceToHost, st));
    VP_CUDA(cudaStreamSynchronize(st));
  }
  return VP_OK;
}
