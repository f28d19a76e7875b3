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
  __device__ __forceinline__ void group_sync() const { __syncwarp(); }
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
__device__ __forceinline__ int next_problem_warp(int* counter, int lane) {
  int p = 0;
  if (lane == 0) p = atomicAdd(counter, 1);
  return __shfl_sync(0xffffffffu, p, 0);
}

struct CtaShared {
  int req[32];
  int active;      // problems (warp groups) that still have work; read only between the two barriers of a round
  int next[16];    // problem index fetched by a group leader
  int anyv[16][2];
  double red[16][2][8];
};

__device__ __forceinline__ void cta_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// CTA execution policy ("transposed chain"), WPP warps per problem:
//  * stage-parallel phases of problem p run in its own group of WPP warps, lanes striding over stages.  WPP = 2
//    puts 64 lanes on the <= 51 stages of a horizon-50 problem (one pass instead of two) and spreads the phase work
//    of 5 resident problems evenly over the 4 SM sub-partitions;
//  * the sequential sweeps of ALL P problems run in lock step in warp 0, lanes 0..P-1 = top halves, P..2P-1 = bottom
//    halves, as a rendezvous round [bar] sweep [bar];
//  * factorisations run in the owner group (two lanes), chunked over rounds, overlapped with the others' sweeps.
template <int WPP>
struct CtaExec {
  int lane, warp, P, N, F;
  double* smem0;
  CtaShared* sh;
  int chunk;       // factor stages per round
  int fpad, xpad;  // bank-conflict pads of every problem's layout in this CTA
  __device__ __forceinline__ int prob() const { return warp / WPP; }
  __device__ __forceinline__ int gl() const { return (warp % WPP) * 32 + lane; }      // lane within the group
  __device__ __forceinline__ void group_sync() const {
    if (WPP == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(2 + prob()), "r"(32 * WPP) : "memory");
  }
  __device__ __forceinline__ void tag(int) {}
  template <class Fn> __device__ __forceinline__ void stages(int n, Fn f) {
    for (int k = gl(); k < n; k += 32 * WPP) f(k);
    group_sync();
  }
  template <class Fn> __device__ __forceinline__ void single(Fn f) {
    if (gl() == 0) f();
    group_sync();
  }
  template <class Fn> __device__ __forceinline__ void reduce_max(int n, double* r, int nr, Fn f) {
    for (int i = 0; i < nr; ++i) r[i] = 0.0;
    for (int k = gl(); k < n; k += 32 * WPP) f(k, r);
    for (int i = 0; i < nr; ++i) {
      double v = r[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
      r[i] = v;
    }
    if (WPP > 1) {
      const int sub = warp % WPP;
      if (lane == 0) for (int i = 0; i < nr; ++i) sh->red[prob()][sub][i] = r[i];
      group_sync();
      for (int i = 0; i < nr; ++i) r[i] = fmax(sh->red[prob()][0][i], sh->red[prob()][1][i]);
    }
    group_sync();
  }
  template <class Fn> __device__ __forceinline__ int any(int n, Fn f) {
    int a = 0;
    for (int k = gl(); k < n; k += 32 * WPP) a |= f(k);
    a = __any_sync(0xffffffffu, a);
    if (WPP > 1) {
      if (lane == 0) sh->anyv[prob()][warp % WPP] = a;
      group_sync();
      a = sh->anyv[prob()][0] | sh->anyv[prob()][1];
    }
    group_sync();
    return a;
  }
  __device__ __forceinline__ int fetch(int* counter) {                 // next problem index for the whole group
    if (WPP == 1) return next_problem_warp(counter, lane);
    if (gl() == 0) sh->next[prob()] = atomicAdd(counter, 1);
    group_sync();
    const int b = sh->next[prob()];
    group_sync();
    return b;
  }
  // one rendezvous round; kind 0 none, 1 sweep request, 2 factor chunk [i0,i1) in the owner group
  __device__ __forceinline__ int round(int kind, const View& w, int i0, int i1) {
    if (gl() == 0) sh->req[prob()] = (kind == 1);
    cta_bar(32 * WPP * P);
    const int snap = sh->active;
    if (warp == 0) {
      // lanes 0..P-1: top halves, lanes P..2P-1: bottom halves of the P problems of this CTA
      const int pr = lane < P ? lane : (lane < 2 * P ? lane - P : 0);
      const bool act = lane < 2 * P && sh->req[pr];
      View v{smem0 + (size_t)pr * F, N, fpad, xpad};
      chain_twisted_lanes<false>(act, lane >= P, lane < P ? lane + P : (lane < 2 * P ? lane - P : lane), v);
    }
    if (kind == 2 && warp % WPP == 0) factor_twisted_lanes(lane, w, i0, i1, i1 >= hmax());
    cta_bar(32 * WPP * P);
    return snap;
  }
  __device__ __forceinline__ void solve(const View& w) { round(1, w, 0, 0); }
  __device__ __forceinline__ int hmax() const { return max(half_top(N), half_bot(N)); }
  __device__ __forceinline__ void factor(const View& w) {
    const int n = hmax();
    for (int i0 = 0; i0 < n; i0 += chunk) round(2, w, i0, min(i0 + chunk, n));
  }
  __device__ __forceinline__ void drain() {
    if (gl() == 0) atomicSub(&sh->active, 1);
    while (round(0, View{smem0, N, fpad, xpad}, 0, 0) > 0) {}
  }
};

}  // namespace mpc
