// mpc_exec.cuh — CUDA execution policies for the per-problem driver (mpc_solve.h): how stage-parallel
// phases, reductions and the sequential chain operations map onto warps / a CTA.
#pragma once
#include <cuda_runtime.h>
#include "mpc_solve.h"

namespace mpc {

// One lane sweeps one half of one problem: `bottom` selects the half, `partner` is the lane that holds the other
// half of the same problem.  Called by ALL 32 lanes of the warp (inactive lanes only take part in the shuffles).
template <bool PIPE>
__device__ __forceinline__ void chain_twisted_lanes(bool active, bool bottom, int partner, const View& w) {
  double a[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, o[6], xm[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const HalfView h = bottom ? w.bottom() : w.top();
  if (active) half_forward<PIPE>(h, a);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 6; ++j) o[j] = __shfl_sync(0xffffffffu, a[5 - j], partner);      // partner's border accumulators, index-reversed
  if (active && !bottom) {
#pragma unroll
    for (int j = 0; j < 6; ++j) a[j] += o[j];
    middle_solve(w, a, xm);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 6; ++j) o[j] = __shfl_sync(0xffffffffu, xm[5 - j], partner);     // bottom lanes receive the reversed middle solution
  if (active) half_backward<PIPE>(h, bottom ? o : xm);
  __syncwarp();
}
// twisted factorisation of local stages [i0, i1) of both halves by two lanes; `last` also forms the middle block
__device__ __forceinline__ void factor_twisted_lanes(int lane, const View& w, int i0, int i1, bool last) {
  if (lane == 0) factor_half(w.top(), i0, i1);
  else if (lane == 1) factor_half(w.bottom(), i0, i1);
  __syncwarp();
  if (last && lane == 0) factor_middle(w);
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Warp execution policy
// ------------------------------------------------------------------------------------------------
struct WarpExec {
  int lane;
  __device__ __forceinline__ void tag(int) {}
  template <class F> __device__ __forceinline__ void stages(int n, F f) {
    for (int k = lane; k < n; k += 32) f(k);
    __syncwarp();
  }
  template <class F> __device__ __forceinline__ void single(F f) {
    if (lane == 0) f();
    __syncwarp();
  }
  template <class F> __device__ __forceinline__ void reduce_max(int n, double* r, int nr, F f) {
    for (int i = 0; i < nr; ++i) r[i] = 0.0;
    for (int k = lane; k < n; k += 32) f(k, r);
    for (int i = 0; i < nr; ++i) {
      double v = r[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
      r[i] = v;
    }
    __syncwarp();
  }
  template <class F> __device__ __forceinline__ int any(int n, F f) {
    int a = 0;
    for (int k = lane; k < n; k += 32) a |= f(k);
    a = __any_sync(0xffffffffu, a);
    __syncwarp();
    return a;
  }
  __device__ __forceinline__ void factor(const View& w) {
    const int hmax = max(half_top(w.N), half_bot(w.N));
    factor_twisted_lanes(lane, w, 0, hmax, true);
  }
  __device__ __forceinline__ void solve(const View& w) {
    chain_twisted_lanes<false>(lane < 2, lane == 1, lane ^ 1, w);
  }
};

// ------------------------------------------------------------------------------------------------
// CTA execution policy ("transposed chain"): the CTA holds P problems, warp p runs the stage-parallel
// phases of problem p, and the sequential triangular sweeps of ALL P problems run in lock step in warp 0
// with lanes = problems (one DFMA warp-instruction then serves P problems instead of one lane).  A chain
// operation is a rendezvous round: [bar] warp 0 sweeps every problem that posted a request [bar].
// Factorisations run in the owner warp, chunked over rounds, overlapped with the other problems' sweeps.
// ------------------------------------------------------------------------------------------------
struct CtaShared {
  int req[32];
  int active;      // warps that still have work; read only between the two barriers of a round
};

__device__ __forceinline__ void cta_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

template <bool PIPE>
struct CtaExec {
  int lane, warp, P, N, F;
  double* smem0;
  CtaShared* sh;
  int chunk;       // factor stages per round
  __device__ __forceinline__ void tag(int) {}
  template <class Fn> __device__ __forceinline__ void stages(int n, Fn f) {
    for (int k = lane; k < n; k += 32) f(k);
    __syncwarp();
  }
  template <class Fn> __device__ __forceinline__ void single(Fn f) {
    if (lane == 0) f();
    __syncwarp();
  }
  template <class Fn> __device__ __forceinline__ void reduce_max(int n, double* r, int nr, Fn f) {
    for (int i = 0; i < nr; ++i) r[i] = 0.0;
    for (int k = lane; k < n; k += 32) f(k, r);
    for (int i = 0; i < nr; ++i) {
      double v = r[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
      r[i] = v;
    }
    __syncwarp();
  }
  template <class Fn> __device__ __forceinline__ int any(int n, Fn f) {
    int a = 0;
    for (int k = lane; k < n; k += 32) a |= f(k);
    a = __any_sync(0xffffffffu, a);
    __syncwarp();
    return a;
  }
  // one rendezvous round; kind 0 none, 1 sweep request, 2 factor chunk [i0,i1) in the owner warp
  __device__ __forceinline__ int round(int kind, const View& w, int i0, int i1) {
    if (lane == 0) sh->req[warp] = (kind == 1);
    cta_bar(32 * P);
    const int snap = sh->active;
    if (warp == 0) {
      // lanes 0..P-1: top halves, lanes P..2P-1: bottom halves of the P problems of this CTA
      const int prob = lane < P ? lane : lane - P;
      const bool act = lane < 2 * P && sh->req[prob < P ? prob : 0];
      View v{smem0 + (size_t)(prob < P ? prob : 0) * F, N};
      chain_twisted_lanes<PIPE>(act, lane >= P, lane < P ? lane + P : (lane < 2 * P ? lane - P : lane), v);
    }
    if (kind == 2) factor_twisted_lanes(lane, w, i0, i1, i1 >= hmax());
    cta_bar(32 * P);
    return snap;
  }
  __device__ __forceinline__ void solve(const View& w) { round(1, w, 0, 0); }
  __device__ __forceinline__ int hmax() const { return max(half_top(N), half_bot(N)); }
  __device__ __forceinline__ void factor(const View& w) {
    const int n = hmax();
    for (int i0 = 0; i0 < n; i0 += chunk) round(2, w, i0, min(i0 + chunk, n));
  }
  __device__ __forceinline__ void drain() {
    if (lane == 0) atomicSub(&sh->active, 1);
    while (round(0, View{smem0, N}, 0, 0) > 0) {}
  }
};

}  // namespace mpc
