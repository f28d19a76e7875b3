// Which part of a chain stage costs what: (a) arithmetic only, (b) loads only, (c) both, single warp.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../rrt_mpc_b200/csrc/mpc_core.h"
using namespace mpc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__global__ void parts(int N, int reps, long long* cyc, double* sink) {
  extern __shared__ double smem[];
  for (int i = threadIdx.x; i < footprint(N); i += blockDim.x) smem[i] = 0.001 * ((i * 37) % 11);
  __syncthreads();
  View w{smem, N, 0, 0};
  double a[6] = {1, 2, 3, 4, 5, 6}, out[6];
  ChainRegs r;
  chain_load_fwd(w, 0, r);
  long long t0 = clock64();
  for (int i = 0; i < reps; ++i) { chain_math_fwd(r, a, out); r.nb[0] += out[0] * 1e-30; }
  long long t1 = clock64();
  double acc = 0;
  for (int i = 0; i < reps; ++i) { chain_load_fwd(w, i % N, r); acc += r.la[3] + r.lc[5] + r.nb[2]; }
  long long t2 = clock64();
  for (int i = 0; i < reps; ++i) { chain_load_fwd(w, i % N, r); chain_math_fwd(r, a, out); double* bk = w.bx(i % N); for (int j = 0; j < 6; ++j) bk[j] = out[j]; }
  long long t3 = clock64();
  // pure dependent-pivot skeleton: 6 pivots, each 7 independent fp64 ops depending on previous pivot
  double x = 1.0, y0 = 0, y1 = 0, y2 = 0, y3 = 0, y4 = 0, y5 = 0;
  for (int i = 0; i < reps * 6; ++i) {
    double nxv = fma(x, 0.999, 1e-3);
    y0 = fma(x, 1.1, y0); y1 = fma(x, 1.2, y1); y2 = fma(x, 1.3, y2); y3 = fma(x, 1.4, y3); y4 = fma(x, 1.5, y4); y5 = x * 1.7 + y5 * 0;
    x = nxv;
  }
  long long t4 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0) / reps; cyc[1] = (t2 - t1) / reps; cyc[2] = (t3 - t2) / reps; cyc[3] = (t4 - t3) / reps; }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a[0] + a[5] + out[3] + acc + x + y0 + y1 + y2 + y3 + y4 + y5;
}
int main() {
  int N = 50; long long* cyc; double* sink;
  CK(cudaMalloc(&cyc, 64)); CK(cudaMalloc(&sink, 148 * 128 * 8));
  CK(cudaFuncSetAttribute(parts, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int threads = 32; threads <= 128; threads *= 2) {
    parts<<<148, threads, footprint(N) * 8>>>(N, 2000, cyc, sink);
    CK(cudaDeviceSynchronize());
    long long h[4]; CK(cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost));
    printf("threads=%3d: math-only %lld cyc/stage, loads-only %lld, load+math+store %lld, skeleton(42 fp64, 6 dependent groups) %lld\n", threads, h[0], h[1], h[2], h[3]);
  }
  return 0;
}
