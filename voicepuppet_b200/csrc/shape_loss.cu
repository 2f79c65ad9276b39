// Training-time twin of the path: the vertex loss of BFMNet,
// voicepuppet/bfmnet/bfmnet.py:215-268 (Shape_formation :215-227, add_cost_function :229-268), forward and
// backward, on the kernels of the hot path.
//
// The reference forms two [B*T, 3N] float32 shapes -- label coefficients and label identity + predicted
// expression -- and takes masked L1 norms of their difference and of its temporal difference.  Identity and
// mean cancel in that difference, so with Delta = ex_label - ex_pred and D = exBase . Delta (K1, the same
// contraction the render path runs per frame):
//   loss  = 1/B sum_{b,t} [t < len_b]     sum_r M_r |D[b,t,r]|
//         + 1/B sum_{b,t} [t < len_b - 1] sum_r M_r |D[b,t,r] - D[b,t+1,r]|
//   dloss/dDelta[b,t,:] = G[b,t,:] . exBase,
//   G[b,t,r] = M_r/B ( [t<len_b] sgn D[t] + [t<len_b-1] sgn(D[t]-D[t+1]) - [t-1<len_b-1] sgn(D[t-1]-D[t]) )
// and dloss/dex_pred = -dloss/dDelta.  Three launches: K1 on Delta, one pass that reduces the loss and
// overwrites D with G in place, and the transposed contraction G . exBase (row slabs, deterministic two-stage
// reduction).  M is the reference's mouth mask (10 on mouth vertices, 1 elsewhere, :134-137), permuted into the
// library's row order by vp_loss_mask_create.
#include <algorithm>
#include <vector>

#include "launch.h"

namespace vp {

namespace {

constexpr int kLossBlock = 256;

// One thread per basis row r of one sequence b, walking t: D[t-1], D[t], D[t+1] live in registers, the loss
// terms are accumulated in double, G[t] overwrites D[t] once D[t+1] has been read.
__global__ void __launch_bounds__(kLossBlock)
loss_and_grad_seed_kernel(float* __restrict__ d, const float* __restrict__ mask, const int* __restrict__ seq_len,
                          int frames, int rows_pad, float inv_batch, double* __restrict__ block_partials) {
  const int b = blockIdx.y;
  const int r = blockIdx.x * kLossBlock + threadIdx.x;
  const int len = min(max(__ldg(seq_len + b), 0), frames);
  double acc = 0.0;
  if (r < rows_pad) {
    const float m = __ldg(mask + r);
    float* col = d + (size_t)b * frames * rows_pad + r;
    float prev = 0.f, cur = col[0];
    float s_prev = 0.f;  // sgn(D[t-1] - D[t]) of the previous step, already masked by [t-1 < len-1]
    for (int t = 0; t < frames; ++t) {
      const float next = (t + 1 < frames) ? col[(size_t)(t + 1) * rows_pad] : 0.f;
      float g = 0.f;
      if (t < len) {
        acc += (double)(m * fabsf(cur));
        g += (cur > 0.f) - (cur < 0.f);
      }
      float s_cur = 0.f;
      if (t < len - 1 && t + 1 < frames) {
        const float dd = cur - next;
        acc += (double)(m * fabsf(dd));
        s_cur = (dd > 0.f) - (dd < 0.f);
      }
      g += s_cur - s_prev;
      col[(size_t)t * rows_pad] = g * m * inv_batch;
      s_prev = s_cur;
      prev = cur;
      cur = next;
    }
    (void)prev;
  }
  // block reduction in a fixed order -> one partial per block
  __shared__ double s[kLossBlock / 32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int w = 0; w < kLossBlock / 32; ++w) total += s[w];
    block_partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = total * (double)inv_batch;
  }
}

__global__ void sum_partials_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += partials[i];  // fixed assignment -> deterministic
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = s[0];
}

// Transposed contraction, stage 1: partial[slab][a][k] = sum_{r in slab} G[a][r] * exb[r][k].
// A CTA owns a slab of kSlabRows rows and a tile of kTileA frames; the basis rows of the slab are staged in
// shared memory once per 64-row step, every thread accumulates kTileA * 64 / 256 outputs in registers.
constexpr int kSlabRows = 1024, kStepRows = 64, kTileA = 32;

__global__ void __launch_bounds__(256)
grad_contract_kernel(const float* __restrict__ g, const float* __restrict__ exb, float* __restrict__ partial,
                     int nframes, int rows_pad) {
  __shared__ float s_e[kStepRows][VP_N_EX + 1];
  __shared__ float s_g[kTileA][kStepRows + 1];
  const int slab = blockIdx.x, a0 = blockIdx.y * kTileA;
  const int k = threadIdx.x & 63, aq = threadIdx.x >> 6;  // thread -> coefficient k, frames aq, aq + 4, ...
  float acc[kTileA / 4];
#pragma unroll
  for (int i = 0; i < kTileA / 4; ++i) acc[i] = 0.f;
  const int r_begin = slab * kSlabRows, r_end = min(rows_pad, r_begin + kSlabRows);
  for (int r0 = r_begin; r0 < r_end; r0 += kStepRows) {
    __syncthreads();
    for (int i = threadIdx.x; i < kStepRows * VP_N_EX; i += 256) {
      const int rr = i / VP_N_EX, kk = i % VP_N_EX;
      s_e[rr][kk] = (r0 + rr < r_end) ? __ldg(exb + (size_t)(r0 + rr) * VP_N_EX + kk) : 0.f;
    }
    for (int i = threadIdx.x; i < kTileA * kStepRows; i += 256) {
      const int aa = i / kStepRows, rr = i % kStepRows;
      s_g[aa][rr] = (a0 + aa < nframes && r0 + rr < r_end) ? g[(size_t)(a0 + aa) * rows_pad + r0 + rr] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < kStepRows; ++rr) {
      const float e = s_e[rr][k];
#pragma unroll
      for (int i = 0; i < kTileA / 4; ++i) acc[i] = fmaf(s_g[aq + 4 * i][rr], e, acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < kTileA / 4; ++i) {
    const int a = a0 + aq + 4 * i;
    if (a < nframes) partial[((size_t)slab * nframes + a) * VP_N_EX + k] = acc[i];
  }
}

// stage 2: out[a][k] = sum_slab partial[slab][a][k], slabs in ascending order
__global__ void grad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int nslabs, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int s = 0; s < nslabs; ++s) acc += partial[(size_t)s * n + i];
  out[i] = acc;
}

}  // namespace
}  // namespace vp

using namespace vp;

// mask[nver*3] float32 in the model's (original) vertex order -> device array in the library's row order
// (rows_pad floats, zero padded), owned by the caller (vp_loss_mask_destroy).
extern "C" int vp_loss_mask_create(vp_model* m, const float* mask_host, float** mask_dev) {
  VP_REQUIRE(m != nullptr && mask_host != nullptr && mask_dev != nullptr, "null argument");
  *mask_dev = nullptr;
  std::lock_guard<std::mutex> lock(m->mu);
  VP_CUDA(cudaSetDevice(m->device));
  std::vector<float> tmp((size_t)m->rows_pad, 0.f);
  const std::vector<int>& i2o = m->topo.v_int2orig;
  for (size_t i = 0; i < i2o.size(); ++i)
    for (int a = 0; a < 3; ++a) tmp[3 * i + a] = mask_host[3 * (size_t)i2o[i] + a];
  float* d = nullptr;
  VP_CUDA(cudaMalloc(reinterpret_cast<void**>(&d), tmp.size() * sizeof(float)));
  VP_CUDA(cudaMemcpy(d, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice));
  *mask_dev = d;
  return VP_OK;
}

extern "C" void vp_loss_mask_destroy(float* mask_dev) {
  if (mask_dev) cudaFree(mask_dev);
}

// delta_ex_dev [batch*frames][64] = ex_label - ex_pred, seq_len_dev [batch] int32, mask_dev from
// vp_loss_mask_create.  loss_dev: one double.  grad_delta_dev (may be NULL): [batch*frames][64] = dloss/dDelta.
// Asynchronous on `stream`; workspaces belong to the model (calls on one model are serialised).
extern "C" int vp_expression_loss_dev(vp_model* m, const float* delta_ex_dev, const int* seq_len_dev,
                                      const float* mask_dev, int batch, int frames, double* loss_dev,
                                      float* grad_delta_dev, void* stream) {
  VP_REQUIRE(m != nullptr && delta_ex_dev && seq_len_dev && mask_dev && loss_dev, "null argument");
  VP_REQUIRE(batch > 0 && frames > 0 && (long long)batch * frames < (1 << 24), "bad batch / frames");
  std::lock_guard<std::mutex> lock(m->mu);
  VP_CUDA(cudaSetDevice(m->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int A = batch * frames;
  const int nblk = (m->rows_pad + kLossBlock - 1) / kLossBlock;
  const int nslabs = (m->rows_pad + kSlabRows - 1) / kSlabRows;
  VP_CUDA(m->ws_disp.reserve((size_t)A * m->rows_pad * sizeof(float), m->device));
  const size_t part_bytes = (size_t)nslabs * A * VP_N_EX * sizeof(float);
  VP_CUDA(m->ws_out.reserve(part_bytes + (size_t)nblk * batch * sizeof(double) + 256, m->device));
  float* d = m->ws_disp.as<float>();
  double* block_partials = reinterpret_cast<double*>(m->ws_out.as<char>() + ((part_bytes + 255) & ~size_t(255)));
  VP_TRY(launch_basis(m, delta_ex_dev, d, A, st));                       // D = exBase . Delta (K1)
  dim3 grid(nblk, batch);
  loss_and_grad_seed_kernel<<<grid, kLossBlock, 0, st>>>(d, mask_dev, seq_len_dev, frames, m->rows_pad,
                                                         1.0f / (float)batch, block_partials);
  VP_LAUNCH_CHECK();
  sum_partials_kernel<<<1, 256, 0, st>>>(block_partials, nblk * batch, loss_dev);
  VP_LAUNCH_CHECK();
  if (grad_delta_dev) {
    float* partial = m->ws_out.as<float>();
    dim3 g2(nslabs, (A + kTileA - 1) / kTileA);
    grad_contract_kernel<<<g2, 256, 0, st>>>(d, m->exb, partial, A, m->rows_pad);
    VP_LAUNCH_CHECK();
    const int n = A * VP_N_EX;
    grad_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(partial, grad_delta_dev, nslabs, n);
    VP_LAUNCH_CHECK();
  }
  return VP_OK;
}
