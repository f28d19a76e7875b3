// Where does a rendezvous round of the CTA kernel go?  One CTA, P warps, P real problems, clock64 around every
// phase and both barriers of every round (dev tool).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../rrt_mpc_b200/csrc/mpc_exec.cuh"
using namespace mpc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct TimedCta : CtaExec<1> {
  long long acc[12]; int cur; long long t;   // 0..7 tags, 8 wait barrier A, 9 chain/idle between barriers, 10 wait B
  __device__ void tag(int g) { long long n = clock64(); acc[cur] += n - t; cur = g; t = n; }
  __device__ int round(int kind, const View& w, int i0, int i1) {
    int saved = cur;
    tag(8);
    if (lane == 0) sh->req[warp] = (kind == 1);
    cta_bar(32 * P);
    tag(9);
    const int snap = sh->active;
    if (warp == 0) {
      const int prob = lane < P ? lane : lane - P;
      const bool act = lane < 2 * P && sh->req[prob < P ? prob : 0];
      View v{smem0 + (size_t)(prob < P ? prob : 0) * F, N, fpad, xpad};
      chain_twisted_lanes<false>(act, lane >= P, lane < P ? lane + P : (lane < 2 * P ? lane - P : lane), v);
    }
    tag(11);
    if (kind == 2) factor_twisted_lanes(lane, w, i0, i1, i1 >= hmax());
    tag(10);
    cta_bar(32 * P);
    tag(saved);
    return snap;
  }
  __device__ void solve(const View& w) { round(1, w, 0, 0); }
  __device__ void factor(const View& w) { const int n = hmax(); for (int i0 = 0; i0 < n; i0 += chunk) round(2, w, i0, min(i0 + chunk, n)); }
  __device__ void drain() { if (lane == 0) atomicSub(&sh->active, 1); while (round(0, View{smem0, N, fpad, xpad}, 0, 0) > 0) {} }
};

__global__ void __launch_bounds__(256) k(Params p, Settings s, int P, int F, const double* x0, const double* ref, const double* up, double* warm, double* out, long long* cyc, int* iters) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, N = p.N;
  CtaShared* sh = reinterpret_cast<CtaShared*>(smem + (size_t)P * F);
  if (threadIdx.x == 0) sh->active = P;
  if (threadIdx.x < 32) sh->req[threadIdx.x] = 0;
  __syncthreads();
  int fpad_, xpad_; layout_pads(N, P, fpad_, xpad_);
  View w{smem + (size_t)warp * F, N, fpad_, xpad_};
  TimedCta ex; ex.lane = lane; ex.warp = warp; ex.P = P; ex.N = N; ex.F = F; ex.smem0 = smem; ex.sh = sh; ex.chunk = (half_bot(N) + 1) / 2; ex.fpad = fpad_; ex.xpad = xpad_;
  for (int i = 0; i < 12; ++i) ex.acc[i] = 0; ex.cur = 0; ex.t = clock64();
  ProblemIO io; io.x0 = x0 + 4 * warp; io.ref = RefWin{ref + 4 * (N + 1) * warp, 0, N + 1, 1.0}; io.u_prev = up + 2 * warp;
  io.warm = warm + 2 * warm_size(N) * warp; io.scratch = io.warm + warm_size(N);
  double* o = out + (6 * N + 16) * warp; io.u0 = o; io.Xp = o + 2; io.Up = o + 2 + 4 * (N + 1);
  int st, info[4]; double pr, du; io.status = &st; io.iters = iters + warp; io.pri_res = &pr; io.dua_res = &du; io.info = info;
  solve_problem(ex, w, p, s, io);
  ex.tag(7);
  ex.drain();
  ex.tag(7);
  if (lane == 0) for (int i = 0; i < 12; ++i) cyc[12 * warp + i] = ex.acc[i];
}

int main(int argc, char** argv) {
  int N = argc > 1 ? atoi(argv[1]) : 50, P = argc > 2 ? atoi(argv[2]) : 5;
  Params p; p.L = 3.5; p.dt = 0.1; p.N = N; double q[4] = {4, 4, 0.6, 0.1}, qn[4] = {8, 8, 1, 0.2};
  for (int i = 0; i < 4; ++i) { p.q[i] = q[i]; p.qn[i] = qn[i]; } p.r[0] = 0.03; p.r[1] = 0.25;
  p.u_lo[0] = -35; p.u_hi[0] = 35; p.u_lo[1] = -0.6; p.u_hi[1] = 0.6; p.v_lo = 0; p.v_hi = 90;
  p.du_lo[0] = -12; p.du_hi[0] = 12; p.du_lo[1] = -0.02; p.du_hi[1] = 0.02; p.w_v = 1e3; p.w_u = 5e2; p.w_du = 5e2;
  Settings s; s.eps_abs = s.eps_rel = 1e-6; s.rho0 = 0.1; s.alpha = 1.6; s.sigma = 1e-6; s.adaptive_rho_tolerance = 5; s.rho_eq_factor = 1e3;
  s.rho_min = 1e-6; s.rho_max = 1e6; s.delta = 1e-6; s.max_iter = 60000; s.check_termination = 25; s.adaptive_rho = 1; s.adaptive_rho_interval = 50;
  s.polish_passes = 3; s.polish_refine_iter = 3; s.warm_start = 0; s.polish_retry = 0; s.early_polish = 0; s.early_polish_start = 50;
  std::vector<double> ref(4 * (N + 1) * P), x0(4 * P), up(2 * P);
  for (int b = 0; b < P; ++b) {
    for (int k2 = 0; k2 <= N; ++k2) { double th = (0.02 + 0.004 * b) * k2; double* r = &ref[4 * ((N + 1) * b + k2)]; r[0] = 100 + 2 * k2 * cos(0.3 * b + th / 2); r[1] = 100 + 2 * k2 * sin(0.3 * b + th / 2); r[2] = 0.3 * b + th; r[3] = 14.0; }
    x0[4 * b] = 101; x0[4 * b + 1] = 99; x0[4 * b + 2] = 0.3 * b + 0.05; x0[4 * b + 3] = 12; up[2 * b] = 1.0; up[2 * b + 1] = 0.05;
  }
  int F = footprint(N);
  double *dref, *dx0, *dup, *dwarm, *dout; long long* dc; int* dit;
  CK(cudaMalloc(&dref, ref.size() * 8)); CK(cudaMalloc(&dx0, x0.size() * 8)); CK(cudaMalloc(&dup, up.size() * 8)); CK(cudaMalloc(&dwarm, 2 * warm_size(N) * 8 * P)); CK(cudaMalloc(&dout, (6 * N + 16) * 8 * P));
  CK(cudaMalloc(&dc, 12 * 8 * P)); CK(cudaMalloc(&dit, 4 * P));
  CK(cudaMemcpy(dref, ref.data(), ref.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dx0, x0.data(), x0.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dup, up.data(), up.size() * 8, cudaMemcpyHostToDevice));
  size_t sm = (size_t)P * F * 8 + sizeof(CtaShared) + 16;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  for (int rep = 0; rep < 2; ++rep) { k<<<1, 32 * P, sm>>>(p, s, P, F, dx0, dref, dup, dwarm, dout, dc, dit); CK(cudaDeviceSynchronize()); }
  std::vector<long long> c(12 * P); std::vector<int> it(P);
  CK(cudaMemcpy(c.data(), dc, 12 * 8 * P, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(it.data(), dit, 4 * P, cudaMemcpyDeviceToHost));
  const char* nm[12] = {"setup", "(solve tag)", "A1", "residuals", "assemble", "(factor tag)", "A2", "polish+out", "wait barA", "chain", "wait barB", "factor chunk"};
  printf("N=%d P=%d F=%d\n", N, P, F);
  for (int b = 0; b < P; ++b) {
    long long tot = 0; for (int i = 0; i < 12; ++i) tot += c[12 * b + i];
    printf(" warp %d iters %4d total %9lld:", b, it[b], tot);
    for (int i = 0; i < 12; ++i) if (c[12 * b + i] > tot / 200) printf("  %s %.1f%%", nm[i], 100.0 * c[12 * b + i] / tot);
    printf("\n");
  }
  return 0;
}
