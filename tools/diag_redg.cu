// Lane-atomic rate of the no-return global reductions the scatter kernel uses (REDG.MAX), 64-bit against 32-bit,
// spread addresses (every lane its own sector) against a pixel-run pattern (runs of 4 consecutive elements).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/diag_redg tools/diag_redg.cu   (GPU box: run it)
#include <cstdio>
#include <cuda_runtime.h>

template <typename T>
__global__ void red_kernel(T* buf, size_t mask, int iters, int run, unsigned long long salt) {
  const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long h = (gtid / run) * 0x9E3779B97F4A7C15ull + salt;
  for (int i = 0; i < iters; ++i) {
    h = h * 6364136223846793005ull + 1442695040888963407ull;
    const size_t idx = (((h >> 20) * run) + (gtid % run)) & mask;
    atomicMax(buf + idx, (T)(h >> 8));
  }
}

template <typename T>
void run_case(const char* name, int run) {
  const size_t n = size_t(1) << 23;  // 8 Mi elements: 64 MB (u64) / 32 MB (u32): L2 resident, like a chunk's hot z-buffer rows
  T* buf;
  cudaMalloc(&buf, n * sizeof(T));
  cudaMemset(buf, 0, n * sizeof(T));
  const int blocks = 148 * 5, threads = 256, iters = 64;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(a);
    red_kernel<T><<<blocks, threads>>>(buf, n - 1, iters, run, 12345ull + rep);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (rep >= 2 && ms < best) best = ms;
  }
  const double lanes = double(blocks) * threads * iters;
  const double cyc_per_lane_sm = best * 1e-3 * 1.965e9 / (lanes / 148.0);
  printf("%-28s %8.1f us  %6.1f G lane-atomics/s  %.2f cycles per lane-atomic per SM (at 1965 MHz)\n", name, best * 1e3,
         lanes / best / 1e6, cyc_per_lane_sm);
  cudaFree(buf);
}

int main() {
  run_case<unsigned long long>("REDG.MAX.64 spread", 1);
  run_case<unsigned int>("REDG.MAX.32 spread", 1);
  run_case<unsigned long long>("REDG.MAX.64 runs of 4", 4);
  run_case<unsigned int>("REDG.MAX.32 runs of 4", 4);
  run_case<unsigned long long>("REDG.MAX.64 runs of 32", 32);
  run_case<unsigned int>("REDG.MAX.32 runs of 32", 32);
  return 0;
}
