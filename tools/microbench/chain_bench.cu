// Microbenchmark of the sequential chain (chain_solve / factor_rows of mpc_core.h) in isolation:
// cycles per call for `lanes` problems swept in lock step by one warp (lanes = problems).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o chain_bench chain_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../rrt_mpc_b200/csrc/mpc_core.h"
using namespace mpc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void bench(int N, int F, int lanes, int reps, long long* cyc, double* sink) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  // fill: diagonally dominant band so that values stay finite
  for (int i = threadIdx.x; i < lanes * F; i += blockDim.x) smem[i] = 0.01 * ((i * 37) % 11);
  __syncthreads();
  if (threadIdx.x < 32 && lane < lanes) {
    View w{smem + (size_t)lane * F, N};
    for (int k = 0; k <= N + 1; ++k) { double* b = w.blk(k); for (int d = 0; d < 44; ++d) b[d] = 0.01 * ((k + d) % 7); for (int j = 0; j < 6; ++j) b[15 + j] = 4.0 + (k % 3); }
    factor_band(w);
  }
  __syncthreads();
  long long t0 = 0, t1 = 0, t2 = 0;
  if (threadIdx.x < 32) {
    View w{smem + (size_t)(lane < lanes ? lane : 0) * F, N};
    __syncwarp();
    t0 = clock64();
    if (lane < lanes) for (int r = 0; r < reps; ++r) chain_solve(w);
    __syncwarp();
    t1 = clock64();
    if (lane < lanes) { for (int k = 0; k <= N + 1; ++k) { double* b = w.blk(k); for (int d = 0; d < 44; ++d) b[d] = 0.01 * ((k + d) % 7); for (int j = 0; j < 6; ++j) b[15 + j] = 4.0 + (k % 3); } }
    __syncwarp();
    long long t1b = clock64();
    if (lane < lanes) factor_band(w);
    __syncwarp();
    t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0) / reps; cyc[1] = t2 - t1b; }
    if (lane < lanes) sink[blockIdx.x * 32 + lane] = w.bx(1)[0];
  }
}

int main(int argc, char** argv) {
  int N = argc > 1 ? atoi(argv[1]) : 50;
  int F = footprint(N);
  long long* cyc; double* sink;
  CK(cudaMalloc(&cyc, 16)); CK(cudaMalloc(&sink, 148 * 32 * 8));
  CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int maxl = (227 * 1024) / (F * 8);
  if (maxl > 32) maxl = 32;
  int ls[] = {1, 2, 5, 8, 13, 16, 32};
  for (int li = 0; li < 7; ++li) {
    int lanes = ls[li]; if (lanes > maxl) continue;
    bench<<<148, 64, lanes * F * 8>>>(N, F, lanes, 20, cyc, sink);
    CK(cudaDeviceSynchronize());
    long long h[2]; CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
    printf("N=%d lanes=%2d: chain_solve %lld cycles (%.1f per stage-pass), factor_band %lld cycles\n", N, lanes, h[0], h[0] / (2.0 * (N + 1)), h[1]);
  }
  return 0;
}
