// cp.async.bulk (TMA 1-D) global(L2) -> shared: latency of one copy and steady-state time per copy at a given
// number of copies in flight, for the tile sizes the factor ring uses.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const double* src, int bytes, int depth, int n, long long* out) {
  extern __shared__ __align__(16) double sm[];
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sm);   // 16 barriers
  double* tiles = sm + 16;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(mbar + i)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  uint32_t phase = 0;
  auto issue = [&](int i) {
    int slot = i % depth; uint32_t mb = s32(mbar + slot);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s32(tiles) + (uint32_t)(slot * bytes)), "l"(src + (size_t)(i % 64) * (bytes / 8)), "r"(bytes), "r"(mb) : "memory");
  };
  auto wait = [&](int i) {
    int slot = i % depth; uint32_t mb = s32(mbar + slot), par = (phase >> slot) & 1u, ok;
    do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(mb), "r"(par) : "memory"); } while (!ok);
    phase ^= 1u << slot;
  };
  // warm L2
  for (int i = 0; i < 64; ++i) { issue(i); wait(i); }
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { issue(i); wait(i); }
  long long t1 = clock64();
  for (int i = 0; i < depth; ++i) issue(i);
  for (int i = 0; i < n; ++i) { wait(i); if (i + depth < n + depth) issue(i + depth); }
  long long t2 = clock64();
  for (int i = 0; i < depth; ++i) wait(n + i);
  out[0] = (t1 - t0) / n; out[1] = (t2 - t1) / n;
}
int main() {
  double* src; long long* out;
  CK(cudaMalloc(&src, 64 * 8192)); CK(cudaMemset(src, 0, 64 * 8192)); CK(cudaMalloc(&out, 16));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  int sizes[] = {368, 1840, 2944, 5888};
  for (int s : sizes) for (int depth : {1, 2, 4, 8}) {
    k<<<148, 32, 128 + depth * s + 64>>>(src, s, depth, 400, out);
    CK(cudaDeviceSynchronize());
    long long h[2]; CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
    printf("bytes %5d depth %d: latency %lld cycles, steady-state %lld cycles/copy\n", s, depth, h[0], h[1]);
  }
  return 0;
}
