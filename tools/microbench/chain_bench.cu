// Microbenchmark of the twisted chain (chain_twisted_lanes / factor of mpc_exec.cuh) in isolation: cycles per
// call for P problems swept in lock step by one warp (lanes 0..P-1 top halves, P..2P-1 bottom halves).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../rrt_mpc_b200/csrc/mpc_exec.cuh"
using namespace mpc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ void fill(const View& w) {
  HalfView hs[2] = {w.top(), w.bottom()};
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k <= hs[s].H; ++k) { double* b = hs[s].blk(k); for (int d = 0; d < 44; ++d) b[d] = 0.01 * ((k + d) % 7); for (int j = 0; j < 6; ++j) b[15 + j] = 4.0 + (k % 3); }
}
template <int PIPE>
__global__ void bench(int N, int F, int P, int reps, long long* cyc, double* sink) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < P * F; i += blockDim.x) smem[i] = 0.01 * ((i * 37) % 11);
  __syncthreads();
  if (threadIdx.x < 32) {
    const int prob = lane < P ? lane : (lane < 2 * P ? lane - P : 0);
    int fpad, xpad; layout_pads(N, P, fpad, xpad);
    View w{smem + (size_t)prob * F, N, fpad, xpad};
    if (lane < P) { fill(w); factor_band(w); }
    __syncwarp();
    const bool act = lane < 2 * P;
    const int partner = lane < P ? lane + P : (lane < 2 * P ? lane - P : lane);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) chain_twisted_lanes<PIPE>(act, lane >= P, partner, w);
    long long t1 = clock64();
    if (lane < P) fill(w);
    __syncwarp();
    long long t2 = clock64();
    if (lane < P) factor_band(w);
    __syncwarp();
    long long t3 = clock64();
    if (lane == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0) / reps; cyc[1] = t3 - t2; }
    if (lane < P) sink[blockIdx.x * 32 + lane] = w.top().bx(1)[0];
  }
}
int main(int argc, char** argv) {
  int N = argc > 1 ? atoi(argv[1]) : 50;
  int F = footprint(N);
  long long* cyc; double* sink;
  CK(cudaMalloc(&cyc, 16)); CK(cudaMalloc(&sink, 148 * 32 * 8));
  CK(cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CK(cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CK(cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int maxl = (227 * 1024) / (F * 8);
  if (maxl > 16) maxl = 16;
  int ls[] = {1, 2, 5};
  for (int pipe = 0; pipe < 3; ++pipe)
    for (int li = 0; li < 3; ++li) {
      int P = ls[li]; if (P > maxl) continue;
      if (pipe == 2) bench<2><<<148, 64, P * F * 8>>>(N, F, P, 20, cyc, sink); else if (pipe) bench<1><<<148, 64, P * F * 8>>>(N, F, P, 20, cyc, sink); else bench<0><<<148, 64, P * F * 8>>>(N, F, P, 20, cyc, sink);
      CK(cudaDeviceSynchronize());
      long long h[2]; CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
      printf("N=%d P=%2d pipe=%d: twisted solve %lld cycles (%.1f per stage-pass of a half), sequential factor %lld cycles\n", N, P, pipe, h[0], h[0] / (2.0 * (half_bot(N) - 1)), h[1]);
    }
  return 0;
}
