// fp64 pipe microbenchmark for B200 (sm_100a): DFMA peak throughput, dependent-chain
// latencies (DFMA, DMUL, SHFL of a double, LDS.64) used as the roofline denominator and
// as design constants for the ADMM kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void dfma_tput(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void lat_kernel(double* out, long long* cyc, int iters, double a, double b) {
  __shared__ double sm[64];
  int lane = threadIdx.x;
  sm[lane] = (double)((lane + 1) & 31);   // pointer-chase table stored as doubles
  sm[lane + 32] = 0.0;
  __syncwarp();
  double x = lane * 1e-3;
  long long t0, t1;
  // DFMA dependent chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
  t1 = clock64();
  if (lane == 0) cyc[0] = t1 - t0;
  // DMUL chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = x * a; x = x * a; x = x * a; x = x * a; }
  t1 = clock64();
  if (lane == 0) cyc[1] = t1 - t0;
  // DADD chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = x + b; x = x + b; x = x + b; x = x + b; }
  t1 = clock64();
  if (lane == 0) cyc[2] = t1 - t0;
  // shuffle of a double (2x SHFL.32) dependent chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
  }
  t1 = clock64();
  if (lane == 0) cyc[3] = t1 - t0;
  // LDS.64 pointer chase
  int idx = lane;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    idx = (int)sm[idx]; idx = (int)sm[idx]; idx = (int)sm[idx]; idx = (int)sm[idx];
  }
  t1 = clock64();
  if (lane == 0) cyc[4] = t1 - t0;
  // shfl + fma (broadcast then fma) chain: the triangular-solve handoff pattern
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    x = fma(__shfl_sync(0xffffffffu, x, 0), a, b); x = fma(__shfl_sync(0xffffffffu, x, 1), a, b);
    x = fma(__shfl_sync(0xffffffffu, x, 2), a, b); x = fma(__shfl_sync(0xffffffffu, x, 3), a, b);
  }
  t1 = clock64();
  if (lane == 0) cyc[5] = t1 - t0;
  // DIV chain (double division) and rcp
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = a / x; x = a / x; x = a / x; x = a / x; }
  t1 = clock64();
  if (lane == 0) cyc[6] = t1 - t0;
  out[lane] = x + idx;
}

template <int ILP>
double run_tput(int blocks, int threads, int iters) {
  double* out; CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) dfma_tput<ILP><<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    dfma_tput<ILP><<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaFree(out));
  double flops = 2.0 * ILP * (double)iters * blocks * threads;
  return flops / (best * 1e-3) / 1e12;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device: %s, SMs %d, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int sms = p.multiProcessorCount;
  printf("{\"fp64_dfma_tflops\": {");
  double best = 0;
  {
    double t;
    t = run_tput<4>(sms * 8, 256, 20000);  printf("\"ilp4_occ2048\": %.3f, ", t); if (t > best) best = t;
    t = run_tput<8>(sms * 8, 256, 10000);  printf("\"ilp8_occ2048\": %.3f, ", t); if (t > best) best = t;
    t = run_tput<8>(sms * 4, 256, 10000);  printf("\"ilp8_occ1024\": %.3f, ", t); if (t > best) best = t;
    t = run_tput<16>(sms * 2, 256, 10000); printf("\"ilp16_occ512\": %.3f, ", t); if (t > best) best = t;
    t = run_tput<8>(sms * 1, 128, 20000);  printf("\"ilp8_occ128\": %.3f, ", t); if (t > best) best = t;
    t = run_tput<1>(sms * 1, 32, 40000);   printf("\"ilp1_1warp\": %.4f", t);
  }
  printf("}, \"fp64_peak_tflops\": %.3f}\n", best);
  double* out; long long* cyc; CK(cudaMalloc(&out, 64 * sizeof(double))); CK(cudaMalloc(&cyc, 8 * sizeof(long long)));
  int iters = 4096;
  lat_kernel<<<1, 32>>>(out, cyc, iters, 0.999999, 1e-6);
  lat_kernel<<<1, 32>>>(out, cyc, iters, 0.999999, 1e-6);
  CK(cudaDeviceSynchronize());
  long long h[8]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
  const char* names[7] = {"dfma", "dmul", "dadd", "shfl_f64", "lds64_chase", "shfl_bcast+dfma", "ddiv"};
  printf("{\"latency_cycles\": {");
  for (int i = 0; i < 7; ++i) printf("\"%s\": %.2f%s", names[i], (double)h[i] / (4.0 * iters), i < 6 ? ", " : "");
  printf("}}\n");
  return 0;
}
