// How fast can ONE warp issue independent DFMAs on sm_100a, and does the active-lane count matter?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int K>
__global__ void k(int iters, int lanes, double a, double b, long long* cyc, double* sink) {
  double acc[K];
#pragma unroll
  for (int i = 0; i < K; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  const bool act = (threadIdx.x & 31) < lanes;
  __syncthreads();
  long long t0 = clock64();
  if (act) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < K; ++i) acc[i] = fma(acc[i], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) s += acc[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  long long* cyc; double* sink;
  cudaMalloc(&cyc, 8); cudaMalloc(&sink, 148 * 1024 * 8);
  const int iters = 4000;
  for (int warps : {1, 2, 4, 8, 16}) {
    for (int lanes : {2, 16, 32}) {
      long long h;
      k<16><<<148, 32 * warps>>>(iters, lanes, 0.999999, 1e-6, cyc, sink); cudaDeviceSynchronize();
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      double c16 = (double)h / (iters * 16.0);
      k<4><<<148, 32 * warps>>>(iters, lanes, 0.999999, 1e-6, cyc, sink); cudaDeviceSynchronize();
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      double c4 = (double)h / (iters * 4.0);
      printf("warps/SM %2d (per sub-partition %.2f) active lanes %2d: %.2f cycles per DFMA instr per warp with 16 independent chains, %.2f with 4\n",
             warps, warps / 4.0, lanes, c16, c4);
    }
  }
  return 0;
}
