// K1, tensor-core flavour: disp[t][r] = sum_k exb[r][k] * ex[t][k] as a tcgen05 3xTF32 GEMM.
// (reference: the expression einsum of Shape_formation, utils/reconstruct_mesh.py:21-22)
//
// GEMM view: D[M = 128 basis rows][N = frames] += A[M][K = 64] * B[N][K]^T, both operands K-major.
//   * A tile (128 x 64 fp32 = 32 KB) arrives by two TMA tensor loads (one per 32-float K half) in
//     the canonical 128-byte-swizzled K-major layout the UMMA shared-memory descriptor expects;
//   * 3xTF32: every fp32 operand x is split in shared memory into hi = tf32(x) and lo = tf32(x - hi);
//     D = A_lo*B_hi + A_hi*B_lo + A_hi*B_hi with FP32 accumulation in TMEM recovers ~fp32 accuracy
//     (the dropped lo*lo term is 2^-22 relative).  The basis is read from HBM once, as fp32;
//   * one elected thread issues the 24 tcgen05.mma (3 products x 8 K-steps of 8) per frame tile and
//     commits them to an mbarrier; the four warps then drain their TMEM lane quarter with
//     tcgen05.ld (32 lanes x 16 columns) and store frame-major, 128 contiguous bytes per warp store.
// The contraction is HBM-bound (K = 64: at most 32 flop/B); tensor cores are used to get the FP32
// SIMT pipe out of the way, not because the math is heavy.
#include <cuda.h>

#include <cstring>

#include "launch.h"
#include "ptx.cuh"

namespace vp {

namespace {

constexpr int kTcM = 128;                 // basis rows per CTA (UMMA M)
constexpr int kTcN = 64;                  // frames per accumulator tile (TMEM columns)
constexpr int kHalfA = kTcM * 128;        // bytes of one K-half (32 floats) of the A tile
constexpr int kHalfB = kTcN * 128;
constexpr int kOffAhi = 0;
constexpr int kOffAlo = 2 * kHalfA;
constexpr int kOffBhi = 4 * kHalfA;
constexpr int kOffBlo = 4 * kHalfA + 2 * kHalfB;
constexpr int kOffBar = 4 * kHalfA + 4 * kHalfB;
constexpr int kTcSmem = kOffBar + 64 + 1024;  // + alignment slack

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi.x = tf32_round(v.x);
  hi.y = tf32_round(v.y);
  hi.z = tf32_round(v.z);
  hi.w = tf32_round(v.w);
  lo.x = tf32_round(v.x - hi.x);
  lo.y = tf32_round(v.y - hi.y);
  lo.z = tf32_round(v.z - hi.z);
  lo.w = tf32_round(v.w - hi.w);
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

__global__ void __launch_bounds__(128)
basis_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const float* __restrict__ ex, float* __restrict__ disp,
                int nframes, int rows_pad) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_mma = bar_tma + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 2);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTcM;

  if (tid == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::mbar_init(bar_tma, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, kTcN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (tid == 0) {
    ptx::mbar_arrive_expect_tx(bar_tma, 2 * kHalfA);
    ptx::tma_load_2d(smem + kOffAhi, &tmap_a, 0, row0, bar_tma);
    ptx::tma_load_2d(smem + kOffAhi + kHalfA, &tmap_a, 32, row0, bar_tma);
  }
  ptx::mbar_wait(bar_tma, 0);

  // split the A tile in place (elementwise, so the swizzled layout is preserved)
  {
    float4* hi4 = reinterpret_cast<float4*>(smem + kOffAhi);
    float4* lo4 = reinterpret_cast<float4*>(smem + kOffAlo);
#pragma unroll 4
    for (int i = tid; i < 2 * kHalfA / 16; i += 128) {
      float4 h, l;
      split4(hi4[i], h, l);
      hi4[i] = h;
      lo4[i] = l;
    }
  }

  const uint32_t a_hi = ptx::smem_u32(smem + kOffAhi), a_lo = ptx::smem_u32(smem + kOffAlo);
  const uint32_t b_hi = ptx::smem_u32(smem + kOffBhi), b_lo = ptx::smem_u32(smem + kOffBlo);
  uint32_t phase = 0;
  for (int t0 = 0; t0 < nframes; t0 += kTcN) {
    const int nt = min(kTcN, nframes - t0);
    const int n_mma = (nt + 15) & ~15;
    // frame coefficients -> hi / lo tiles in the same swizzled K-major layout: row n (frame), 16-byte
    // chunk c of K-half h lives at h * kHalfB + (n / 8) * 1024 + (n % 8) * 128 + ((c ^ (n % 8)) * 16)
    for (int q = tid; q < n_mma * 16; q += 128) {
      const int n = q >> 4, c16 = q & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < nt) v = __ldg(reinterpret_cast<const float4*>(ex + (size_t)(t0 + n) * VP_N_EX) + c16);
      float4 h, l;
      split4(v, h, l);
      const int off = (c16 >> 3) * kHalfB + (n >> 3) * 1024 + (n & 7) * 128 + (((c16 & 7) ^ (n & 7)) << 4);
      *reinterpret_cast<float4*>(smem + kOffBhi + off) = h;
      *reinterpret_cast<float4*>(smem + kOffBlo + off) = l;
    }
    ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    ptx::tc_fence_before();    // and the previous tile's tcgen05.ld are ordered before the next MMA
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t idesc = instr_desc_tf32(kTcM, n_mma);
      uint32_t acc = 0;
#pragma unroll
      for (int part = 0; part < 3; ++part) {  // small terms first
        const uint32_t a_base = (part == 0) ? a_lo : a_hi;
        const uint32_t b_base = (part == 1) ? b_lo : b_hi;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // K = 8 per instruction; 4 steps per 128-byte swizzle atom
          const uint32_t koff_a = (ks >> 2) * kHalfA + (ks & 3) * 32;
          const uint32_t koff_b = (ks >> 2) * kHalfB + (ks & 3) * 32;
          ptx::mma_tf32_ss(tmem_base, smem_desc_sw128(a_base + koff_a), smem_desc_sw128(b_base + koff_b), idesc, acc);
          acc = 1;
        }
      }
      ptx::tc_commit(bar_mma);
    }
    ptx::mbar_wait(bar_mma, phase);
    phase ^= 1;
    ptx::tc_fence_after();
    // epilogue: warp w owns TMEM lanes 32w..32w+31 (= basis rows), columns = frames
    float* out = disp + (size_t)t0 * rows_pad + row0 + warp * 32 + lane;
    for (int c0 = 0; c0 < n_mma; c0 += 16) {
      uint32_t r[16];
      ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c0 + j < nt) out[(size_t)(c0 + j) * rows_pad] = __uint_as_float(r[j]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTcN);
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
  VP_CUDA(cudaFuncSetAttribute(basis_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
  m->have_tmap = true;
  return VP_OK;
}

int launch_basis_tc(vp_model* m, const float* ex_dev, float* disp_dev, int nframes, cudaStream_t st) {
  if (nframes == 0) return VP_OK;
  VP_REQUIRE(m->have_tmap, "tensor map not prepared");
  CUtensorMap map;
  std::memcpy(&map, m->tmap_exb, sizeof(map));
  basis_tc_kernel<<<m->rows_pad / kTcM, 128, kTcSmem, st>>>(map, ex_dev, disp_dev, nframes, m->rows_pad);
  VP_LAUNCH_CHECK();
  return VP_OK;
}

}  // namespace vp
