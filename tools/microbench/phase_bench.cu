#include <string.h>
// Per-phase cycle breakdown of one warp solving one real problem (dev tool): tags set by solve_problem.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../rrt_mpc_b200/csrc/mpc_solve.h"
using namespace mpc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
struct TimedExec {
  int lane; long long acc[8]; int cnt[8]; int cur; long long t;
  __device__ void tag(int g) { long long n = clock64(); acc[cur] += n - t; cnt[g]++; cur = g; t = n; }
  template <class F> __device__ __forceinline__ void stages(int n, F f) { for (int k = lane; k < n; k += 32) f(k); __syncwarp(); }
  template <class F> __device__ __forceinline__ void single(F f) { if (lane == 0) f(); __syncwarp(); }
  template <class F> __device__ __forceinline__ void reduce_max(int n, double* r, int nr, F f) {
    for (int i = 0; i < nr; ++i) r[i] = 0.0;
    for (int k = lane; k < n; k += 32) f(k, r);
    for (int i = 0; i < nr; ++i) { double v = r[i]; for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); r[i] = v; }
    __syncwarp();
  }
  template <class F> __device__ __forceinline__ int any(int n, F f) { int a = 0; for (int k = lane; k < n; k += 32) a |= f(k); a = __any_sync(0xffffffffu, a); __syncwarp(); return a; }
  __device__ __forceinline__ void factor(const View& w) { if (lane == 0) factor_band(w); __syncwarp(); }
  __device__ __forceinline__ void solve(const View& w) { if (lane == 0) chain_solve(w); __syncwarp(); }
  __device__ __forceinline__ void solve_iter(const View& w) { solve(w); }
};
__global__ void k(Params p, Settings s, const double* x0, const double* ref, const double* up, double* warm, double* out, long long* cyc, int* cnt, int* iters) {
  extern __shared__ double smem[];
  View w{smem, p.N, 0, 0};
  TimedExec ex; ex.lane = threadIdx.x; for (int i = 0; i < 8; ++i) { ex.acc[i] = 0; ex.cnt[i] = 0; } ex.cur = 0; ex.t = clock64();
  ProblemIO io; io.x0 = x0; io.ref = RefWin{ref, 0, p.N + 1, 1.0}; io.u_prev = up; io.warm = warm; io.scratch = warm + warm_size(p.N);
  io.u0 = out; io.Xp = out + 2; io.Up = out + 2 + 4 * (p.N + 1); int st, info[4]; double pr, du;
  io.status = &st; io.iters = iters; io.pri_res = &pr; io.dua_res = &du; io.info = info;
  solve_problem(ex, w, p, s, io);
  ex.tag(7);
  if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) { cyc[i] = ex.acc[i]; cnt[i] = ex.cnt[i]; }
}
int main(int argc, char** argv) {
  int N = argc > 1 ? atoi(argv[1]) : 50;
  Params p; p.L = 3.5; p.dt = 0.1; p.N = N; double q[4] = {4, 4, 0.6, 0.1}, qn[4] = {8, 8, 1, 0.2};
  memset(p.pq, 0, sizeof p.pq); memset(p.pqn, 0, sizeof p.pqn); memset(p.pr, 0, sizeof p.pr);
  for (int i = 0; i < 4; ++i) { p.pq[i][i] = 2.0 * q[i]; p.pqn[i][i] = 2.0 * qn[i]; } p.pr[0][0] = 0.06; p.pr[1][1] = 0.5;
  p.u_lo[0] = -35; p.u_hi[0] = 35; p.u_lo[1] = -0.6; p.u_hi[1] = 0.6; p.v_lo = 0; p.v_hi = 90;
  p.du_lo[0] = -12; p.du_hi[0] = 12; p.du_lo[1] = -0.02; p.du_hi[1] = 0.02; p.w_v = 1e3; p.w_u = 5e2; p.w_du = 5e2;
  Settings s; s.eps_abs = s.eps_rel = 1e-6; s.rho0 = 0.1; s.alpha = 1.6; s.sigma = 1e-6; s.adaptive_rho_tolerance = 5; s.rho_eq_factor = 1e3;
  s.rho_min = 1e-6; s.rho_max = 1e6; s.delta = 1e-6; s.max_iter = 60000; s.check_termination = 25; s.adaptive_rho = 1; s.adaptive_rho_interval = 50;
  s.polish_passes = 3; s.polish_refine_iter = 3; s.warm_start = 0; s.polish_retry = 0; s.early_polish = 0; s.early_polish_start = 50;
  std::vector<double> ref(4 * (N + 1)), x0(4), up(2, 0.0);
  for (int k2 = 0; k2 <= N; ++k2) { double th = 0.03 * k2; ref[4 * k2] = 100 + 2 * k2 * cos(0.3 + th / 2); ref[4 * k2 + 1] = 100 + 2 * k2 * sin(0.3 + th / 2); ref[4 * k2 + 2] = 0.3 + th; ref[4 * k2 + 3] = 14.0; }
  x0[0] = 101; x0[1] = 99; x0[2] = 0.35; x0[3] = 12; up[0] = 1.0; up[1] = 0.05;
  double *dref, *dx0, *dup, *dwarm, *dout; long long* dc; int *dn, *dit;
  CK(cudaMalloc(&dref, ref.size() * 8)); CK(cudaMalloc(&dx0, 32)); CK(cudaMalloc(&dup, 16)); CK(cudaMalloc(&dwarm, 2 * warm_size(N) * 8)); CK(cudaMalloc(&dout, (6 * N + 16) * 8));
  CK(cudaMalloc(&dc, 64)); CK(cudaMalloc(&dn, 32)); CK(cudaMalloc(&dit, 4));
  CK(cudaMemcpy(dref, ref.data(), ref.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dx0, x0.data(), 32, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dup, up.data(), 16, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, footprint(N) * 8));
  for (int rep = 0; rep < 2; ++rep) { k<<<1, 32, footprint(N) * 8>>>(p, s, dx0, dref, dup, dwarm, dout, dc, dn, dit); CK(cudaDeviceSynchronize()); }
  long long c[8]; int n[8], it; CK(cudaMemcpy(c, dc, 64, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(n, dn, 32, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&it, dit, 4, cudaMemcpyDeviceToHost));
  const char* nm[8] = {"setup+initial", "chain_solve", "update(A1)", "residuals", "assemble", "factor", "rhs(A2)", "polish+out"};
  printf("N=%d iters=%d\n", N, it);
  for (int i = 0; i < 8; ++i) printf("  %-14s calls %5d  total %10lld cyc  per call %8.0f\n", nm[i], n[i], c[i], n[i] ? (double)c[i] / n[i] : 0.0);
  return 0;
}
