// K1, tensor-core flavour: disp[t][r] = sum_k exb[r][k] * ex[t][k] as a tcgen05 3xTF32 GEMM.
// (reference: the expression einsum of Shape_formation, utils/reconstruct_mesh.py:21-22)
//
// GEMM view: D[M = 128 basis rows][N = frames <= 128] = A[M][K = 64] * B[N][K]^T, operands K-major.
// One launch contracts any number of frames: the work items are (frame block of 128, basis row tile) pairs,
// numbered block-major (w = block * ntiles + tile), so that every CTA is on the same frame block at about the same
// time and the 27 MB basis is read from HBM by the first block and from L2 by the others; the 148 CTAs take the
// items round-robin, which also spreads the 837 % 148 remainder over the blocks (round 2: one 128-frame launch per
// block cost 23.5 us by CUDA events for a 13.8 us body).
// Persistent, warp-specialised kernel, one CTA per SM, each CTA walks the items w, w+grid, ...:
//   warp 0 (one lane)   TMA producer: two tensor loads per tile (one per 32-float K half) land the
//                       128 x 64 fp32 tile (32 KB) in the canonical 128-byte-swizzled K-major layout
//                       the UMMA shared-memory descriptor expects; 3-stage ring, plus L2 prefetches
//                       (cp.async.bulk.prefetch.tensor) four tiles ahead
//   warps 10-17         3xTF32 split of the landed tile: the tensor core reads the top 19 bits of each fp32
//                       container, so the landed tile IS the "hi" operand; lo = x - trunc_tf32(x) goes
//                       into a second tile (elementwise, so the swizzle is kept); 2-stage ring
//   warp 1 (one lane)   issues 24 tcgen05.mma per tile: (A_lo*B_hi + A_hi*B_lo + A_hi*B_hi) x 8 K-steps
//                       of 8, FP32 accumulation in TMEM (2 accumulators of 128 columns), and commits
//                       them to the mbarriers that free the A stage and release the epilogue
//   warps 2-9           epilogue, two warps per TMEM lane quarter: tcgen05.ld (32 rows x 16 frames per
//                       load), frame-major stores, 128 contiguous bytes per warp store
// The frame coefficients (B) of a block are split by the split warps when a CTA enters the block: they wait for the
// MMAs of the previous item (the last readers of the old B tiles), rewrite the tiles and only then release the
// item to the MMA issuer; the epilogue still has two accumulators to drain meanwhile, so the stores never pause.
// The dropped lo*lo term is 2^-22 relative, so the result has fp32-grade accuracy; the basis is read from HBM
// once per launch, as fp32.
// The contraction is HBM-bound (K = 64: at most 32 flop/B); tensor cores are used to get the FP32
// SIMT pipe out of the way, not because the math is heavy.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "launch.h"
#include "ptx.cuh"

namespace vp {

namespace {

constexpr int kTcM = 128;                 // basis rows per tile (UMMA M)
constexpr int kTcN = 128;                 // max frames per launch (UMMA N, TMEM columns per accumulator)
constexpr int kStagesL = 2;               // "lo" tiles and TMEM accumulators
constexpr int kHalfA = kTcM * 128;        // bytes of one K-half (32 floats) of an A tile
constexpr int kTileA = 2 * kHalfA;        // 32 KB
constexpr int kOffAhi = 0;
constexpr int kStagesA = 3;               // TMA ring of fp32 A tiles (the "hi" operand in place)
constexpr int kOffAlo = kStagesA * kTileA;
constexpr int kOffB = kOffAlo + kStagesL * kTileA;  // B hi/lo tiles (1024-byte aligned; size depends on the frame count), then the barriers
// warp roles: 0 = TMA producer, 1 = MMA issuer, 2-9 = epilogue (two warps per TMEM lane quarter),
// 10-17 = operand split
constexpr int kWarpTma = 0, kWarpMma = 1, kWarpEpi0 = 2, kWarpSplit0 = 10;
constexpr int kEpiThreads = 256, kSplitThreads = 256;
constexpr int kTcThreads = (kWarpSplit0 * 32) + kSplitThreads;  // 576
constexpr int kPrefetchTiles = 4;         // L2 prefetch distance of the TMA producer, in tiles
constexpr uint32_t kTf32Mask = 0xFFFFE000u;

inline int tc_smem_bytes(int n_mma) { return kOffB + 2 * n_mma * 256 + 128; }

// 3xTF32 split by truncation: hi keeps the 19 bits the tensor core reads, lo = x - hi is exact in fp32
// (|lo| < 2^-10 |x|) and is truncated to its own top 19 bits by the tensor core: 2^-21 relative overall.
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & kTf32Mask);
  hi.y = __uint_as_float(__float_as_uint(v.y) & kTf32Mask);
  hi.z = __uint_as_float(__float_as_uint(v.z) & kTf32Mask);
  hi.w = __uint_as_float(__float_as_uint(v.w) & kTf32Mask);
  lo.x = v.x - hi.x;
  lo.y = v.y - hi.y;
  lo.z = v.z - hi.z;
  lo.w = v.w - hi.w;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 in [0,14), LBO (unused for swizzled K-major) in [16,30), SBO = 1024 B (8 rows x
// 128 B) >> 4 in [32,46), descriptor version 1 in [46,48), layout type 2 = SWIZZLE_128B in [61,64).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// kind::tf32 instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (2 at bits 7-9 / 10-12),
// both K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__device__ __forceinline__ uint32_t instr_desc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tma_prefetch_2d(const void* tensor_map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tensor_map), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(kTcThreads, 1)
basis_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const float* __restrict__ ex, float* __restrict__ disp,
                int nframes, int rows_pad, int ntiles, long long* __restrict__ trace, int store_policy) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // optional per-role timeline of CTA 0 (diagnostics): trace[role * 64 + it * 4 + k] = clock64()
  const bool tracing = trace != nullptr && blockIdx.x == 0;
#define VP_TRACE(role, it, k) do { if (tracing && (it) < 16) trace[(role) * 64 + (it) * 4 + (k)] = clock64(); } while (0)
  const int nblocks = (nframes + kTcN - 1) / kTcN;
  const int total = nblocks * ntiles;      // work items (frame block, row tile), block-major
  const int n_cap = (min(nframes, kTcN) + 15) & ~15;   // widest block of this launch: sizes the B tiles
  uint8_t* smem_b = smem + kOffB;          // B hi tile, then B lo tile; each 2 K-halves of n_mma * 128 bytes
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_b + 4 * n_cap * 128);  // TMA landed
  uint64_t* bar_afree = bar_full + kStagesA;                         // [3] MMAs that read the A stage completed
  uint64_t* bar_split = bar_afree + kStagesA;                        // [2] hi/lo tiles ready for the MMA
  uint64_t* bar_mma = bar_split + kStagesL;                          // [2] accumulator complete (and lo tile free)
  uint64_t* bar_accfree = bar_mma + kStagesL;                        // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_accfree + kStagesL);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction
  const int lane = tid & 31;

  if (tid == 0) VP_TRACE(0, 0, 0);                       // kernel entry
  if (trace != nullptr && tid == 0) {                    // per-CTA entry time (ns): trace[256 + 2 * cta]
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    trace[256 + 2 * blockIdx.x] = (long long)ns;
  }
  if (tid == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
    ptx::prefetch_tensormap(&tmap_a);
    for (int s = 0; s < kStagesA; ++s) {
      ptx::mbar_init(bar_full + s, 1);
      ptx::mbar_init(bar_afree + s, 1);
    }
    for (int s = 0; s < kStagesL; ++s) {
      ptx::mbar_init(bar_split + s, kSplitThreads);
      ptx::mbar_init(bar_mma + s, 1);
      ptx::mbar_init(bar_accfree + s, kEpiThreads);
    }
    ptx::fence_mbar_init();
    // The first A tiles do not depend on anything the other warps set up: start the HBM stream now,
    // under the TMEM allocation and the split of the frame coefficients.
    for (int p = 0; p < kStagesA; ++p) {
      const int wp = (int)blockIdx.x + p * (int)gridDim.x;
      if (wp < total) {
        const int mp = wp % ntiles;
        uint8_t* dst = smem + kOffAhi + p * kTileA;
        ptx::mbar_arrive_expect_tx(bar_full + p, kTileA);
        ptx::tma_load_2d(dst, &tmap_a, 0, mp * kTcM, bar_full + p);
        ptx::tma_load_2d(dst + kHalfA, &tmap_a, 32, mp * kTcM, bar_full + p);
      }
    }
    for (int p = kStagesA; p < kStagesA + kPrefetchTiles; ++p) {
      const int wp = (int)blockIdx.x + p * (int)gridDim.x;
      if (wp < ntiles) {  // first frame block only: the later ones find the basis in L2
        tma_prefetch_2d(&tmap_a, 0, wp * kTcM);
        tma_prefetch_2d(&tmap_a, 32, wp * kTcM);
      }
    }
  }
  if (warp == kWarpMma) {
    ptx::tmem_alloc(tmem_slot, kStagesL * kTcN);
    ptx::tmem_relinquish();
  }
  // frame coefficients of frame block fb -> hi / lo tiles in the swizzled K-major layout: row n (frame), 16-byte
  // chunk c of K-half h lives at h * half_b + (n / 8) * 1024 + (n % 8) * 128 + ((c ^ (n % 8)) * 16)
  auto split_b = [&](int fb) {
    const int nb = min(kTcN, nframes - fb * kTcN), nb_mma = (nb + 15) & ~15, half_b = nb_mma * 128;
    const float* exb = ex + (size_t)fb * kTcN * VP_N_EX;
    uint8_t* bhi = smem_b;
    uint8_t* blo = smem_b + 2 * half_b;
    for (int q = tid - kWarpSplit0 * 32; q < nb_mma * 16; q += kSplitThreads) {
      const int n = q >> 4, c16 = q & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < nb) v = __ldg(reinterpret_cast<const float4*>(exb + (size_t)n * VP_N_EX) + c16);
      float4 h, l;
      split4(v, h, l);
      const int off = (c16 >> 3) * half_b + (n >> 3) * 1024 + (n & 7) * 128 + (((c16 & 7) ^ (n & 7)) << 4);
      *reinterpret_cast<float4*>(bhi + off) = h;
      *reinterpret_cast<float4*>(blo + off) = l;
    }
  };
  if (warp >= kWarpSplit0 && (int)blockIdx.x < total) {
    split_b((int)blockIdx.x / ntiles);
    ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  }
  if (tid == 0) VP_TRACE(0, 0, 1);                       // prologue of thread 0 done (barriers, first TMA loads, prefetches)
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (tid == 0) VP_TRACE(0, 0, 2);                       // every warp past the set-up barrier (TMEM allocated, B split)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int first = blockIdx.x, step = gridDim.x;

  if (warp == kWarpTma) {
    // ===== TMA producer (whole warp loops, one elected lane issues); stages 0..kStagesA-1 of the first
    // round were issued in the prologue =====
    int it = kStagesA;
    for (int w = first + kStagesA * step; w < total; w += step, ++it) {
      const int s = it % kStagesA;
      const int m = w % ntiles;
      ptx::mbar_wait(bar_afree + s, ((it / kStagesA) - 1) & 1);
      if (elect_one()) {
        const int wp = w + kPrefetchTiles * step;
        if (wp < ntiles) {
          tma_prefetch_2d(&tmap_a, 0, wp * kTcM);
          tma_prefetch_2d(&tmap_a, 32, wp * kTcM);
        }
        uint8_t* dst = smem + kOffAhi + s * kTileA;
        ptx::mbar_arrive_expect_tx(bar_full + s, kTileA);
        ptx::tma_load_2d(dst, &tmap_a, 0, m * kTcM, bar_full + s);
        ptx::tma_load_2d(dst + kHalfA, &tmap_a, 32, m * kTcM, bar_full + s);
        VP_TRACE(0, it, 2);
      }
      __syncwarp();
    }
  } else if (warp == kWarpMma) {
    // ===== MMA issuer (whole warp loops, one elected lane issues) =====
    const uint64_t desc_hi = (64ull << 32) | (1ull << 46) | (2ull << 61);  // SBO, version, SWIZZLE_128B
    const uint32_t smem_base = ptx::smem_u32(smem);
    int it = 0;
    for (int w = first; w < total; w += step, ++it) {
      const int sa = it % kStagesA, sl = it & 1;
      const int fb = w / ntiles;
      const int n_mma = (min(kTcN, nframes - fb * kTcN) + 15) & ~15, half_b = n_mma * 128;
      const uint32_t idesc = instr_desc_tf32(kTcM, n_mma);
      const uint32_t b_hi = (smem_base + kOffB) >> 4, b_lo = (smem_base + kOffB + 2 * half_b) >> 4;
      const uint32_t half_b16 = half_b >> 4;
      VP_TRACE(1, it, 0);
      ptx::mbar_wait(bar_split + sl, (it >> 1) & 1);
      VP_TRACE(1, it, 1);
      if (it >= kStagesL) ptx::mbar_wait(bar_accfree + sl, ((it >> 1) - 1) & 1);
      VP_TRACE(1, it, 2);
      ptx::tc_fence_after();
      const uint32_t a_hi = (smem_base + kOffAhi + sa * kTileA) >> 4, a_lo = (smem_base + kOffAlo + sl * kTileA) >> 4;
      const uint32_t d_tmem = tmem_base + (uint32_t)(sl * kTcN);
      if (elect_one()) {
#pragma unroll
        for (int part = 0; part < 3; ++part) {  // small terms first
          const uint32_t a_base = (part == 0) ? a_lo : a_hi;
          const uint32_t b_base = (part == 1) ? b_lo : b_hi;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // K = 8 per instruction; 4 steps per 128-byte swizzle atom
            const uint32_t ka = a_base + (ks >> 2) * (kHalfA >> 4) + (ks & 3) * 2;
            const uint32_t kb = b_base + (ks >> 2) * half_b16 + (ks & 3) * 2;
            ptx::mma_tf32_ss(d_tmem, desc_hi | (1ull << 16) | (uint64_t)(ka & 0x3FFFu),
                             desc_hi | (1ull << 16) | (uint64_t)(kb & 0x3FFFu), idesc, (part | ks) != 0);
          }
        }
        ptx::tc_commit(bar_mma + sl);
        ptx::tc_commit(bar_afree + sa);
      }
      __syncwarp();
      VP_TRACE(1, it, 3);
    }
  } else if (warp < kWarpSplit0) {
    // ===== epilogue warps: drain accumulator `it` (TMEM -> registers -> global, frame-major) =====
    const int ew = warp - kWarpEpi0;     // 0..7
    const int quarter = warp & 3;        // TMEM lanes this warp may read: 32 * (warp % 4) ..
    const int group = ew >> 2;           // two warps per lane quarter alternate over 16-column chunks
    int it = 0;
    for (int w = first; w < total; w += step, ++it) {
      const int sl = it & 1;
      const int fb = w / ntiles, m = w - fb * ntiles;
      const int nb = min(kTcN, nframes - fb * kTcN), n_mma = (nb + 15) & ~15;  // frames of this block
      if (ew == 0 && lane == 0) VP_TRACE(3, it, 0);
      ptx::mbar_wait(bar_mma + sl, (it >> 1) & 1);
      if (ew == 0 && lane == 0) VP_TRACE(3, it, 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(sl * kTcN);
      float* out = disp + (size_t)fb * kTcN * rows_pad + (size_t)m * kTcM + quarter * 32 + lane;
      for (int c0 = group * 16; c0 < n_mma; c0 += 32) {
        uint32_t r[16];
        ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, r);
        ptx::tmem_ld_wait();
        float* o = out + (size_t)c0 * rows_pad;
        if (c0 + 16 <= nb) {
          if (store_policy == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) __stcs(o + (size_t)j * rows_pad, __uint_as_float(r[j]));
          } else {  // default policy: the displacements stay in L2 for the vertex kernel that reads them next
#pragma unroll
            for (int j = 0; j < 16; ++j) o[(size_t)j * rows_pad] = __uint_as_float(r[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < nb) __stcs(o + (size_t)j * rows_pad, __uint_as_float(r[j]));
        }
      }
      // (loading all of a warp's chunks back to back, waiting once and releasing the accumulator before the stores
      // was measured in round 2: 19.5 vs 18.4 us at 75 frames, equal at 96 / 128 -- no gain, not kept)
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_accfree + sl);
      if (ew == 0 && lane == 0) VP_TRACE(3, it, 2);
    }
  } else {
    // ===== split warps: hi / lo tiles of A tile `it` =====
    const int wt = tid - kWarpSplit0 * 32;  // 0..255
    int it = 0, fb_cur = first / ntiles;
    for (int w = first; w < total; w += step, ++it) {
      const int sa = it % kStagesA, sl = it & 1;
      if (wt == 0) VP_TRACE(2, it, 0);
      if (it >= kStagesL) ptx::mbar_wait(bar_mma + sl, ((it >> 1) - 1) & 1);  // MMAs of tile it-2 no longer read lo[sl]
      const int fb = w / ntiles;
      if (fb != fb_cur) {
        // entering a new frame block: the MMAs of item it-1 are the last readers of the old B tiles
        ptx::mbar_wait(bar_mma + (sl ^ 1), ((it - 1) >> 1) & 1);
        split_b(fb);
        fb_cur = fb;
      }
      ptx::mbar_wait(bar_full + sa, (it / kStagesA) & 1);
      if (wt == 0) VP_TRACE(2, it, 1);
      float4* hi4 = reinterpret_cast<float4*>(smem + kOffAhi + sa * kTileA);
      float4* lo4 = reinterpret_cast<float4*>(smem + kOffAlo + sl * kTileA);
      float4 v[kTileA / 16 / kSplitThreads];
#pragma unroll
      for (int u = 0; u < kTileA / 16 / kSplitThreads; ++u) v[u] = hi4[wt + u * kSplitThreads];
#pragma unroll
      for (int u = 0; u < kTileA / 16 / kSplitThreads; ++u) {
        float4 h, l;
        split4(v[u], h, l);
#ifdef VP_TC_HI_WRITEBACK  // not needed: kind::tf32 reads the top 19 bits of the fp32 container (verified bit-identical)
        hi4[wt + u * kSplitThreads] = h;
#endif
        lo4[wt + u * kSplitThreads] = l;
      }
      if (wt == 0) VP_TRACE(2, it, 2);
      ptx::fence_proxy_async();
      ptx::mbar_arrive(bar_split + sl);
      if (wt == 0) VP_TRACE(2, it, 3);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (trace != nullptr && warp == kWarpEpi0 && lane == 0) {   // per-CTA exit time (ns), after the last epilogue
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    trace[257 + 2 * blockIdx.x] = (long long)ns;
  }
  if (warp == kWarpMma) ptx::tmem_dealloc(tmem_base, kStagesL * kTcN);
#undef VP_TRACE
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

}  // namespace

static_assert(sizeof(CUtensorMap) == sizeof(((vp_model*)nullptr)->tmap_exb), "tensor map storage size");

// Builds the TMA descriptor of the expression basis (once per model).
int basis_tc_prepare(vp_model* m) {
  m->have_tmap = false;
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return VP_ERR_CUDA;
  }
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)VP_N_EX, (cuuint64_t)m->rows_pad};
  const cuuint64_t strides[1] = {(cuuint64_t)VP_N_EX * sizeof(float)};
  const cuuint32_t box[2] = {32, (cuuint32_t)kTcM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, m->exb, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return VP_ERR_CUDA;
  }
  std::memcpy(m->tmap_exb, &map, sizeof(map));
  VP_CUDA(cudaFuncSetAttribute(basis_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(kTcN)));
  m->have_tmap = true;
  return VP_OK;
}

// One launch whatever the frame count: the kernel walks (frame block of 128, row tile) items, the basis comes from
// HBM for the first block and from L2 for the others.
int launch_basis_tc(vp_model* m, const float* ex_dev, float* disp_dev, int nframes, cudaStream_t st,
                    long long* trace_dev) {
  if (nframes == 0) return VP_OK;
  VP_REQUIRE(m->have_tmap, "tensor map not prepared");
  CUtensorMap map;
  std::memcpy(&map, m->tmap_exb, sizeof(map));
  const int ntiles = m->rows_pad / kTcM;
  VP_REQUIRE((long long)((nframes + kTcN - 1) / kTcN) * ntiles < (1ll << 30), "frame count too large for one launch");
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
  const int grid = std::min(ntiles, sms);
  const int store_policy = 0;  // streaming stores (default-policy stores measured equal in round 1 and again in round 2, call 27)
  const int n_cap = (std::min(nframes, kTcN) + 15) & ~15;
  basis_tc_kernel<<<grid, kTcThreads, tc_smem_bytes(n_cap), st>>>(map, ex_dev, disp_dev, nframes, m->rows_pad, ntiles,
                                                                  trace_dev, store_policy);
  VP_LAUNCH_CHECK();
  return VP_OK;
}

}  // namespace vp
