// mpc_exec.cuh — CUDA execution policy for the per-problem driver (mpc_solve.h): how stage-parallel
// phases, reductions and the sequential chain operations map onto the warps that own a problem.
#pragma once
#include <cuda_runtime.h>
#include "mpc_solve.h"

namespace mpc {

// One lane sweeps one half of one problem: `bottom` selects the half, `partner` is the lane that holds the other
// half of the same problem.  Called by ALL 32 lanes of the warp (inactive lanes only take part in the shuffles).
// PIPE selects the sweep variant of mpc_core.h: 0 plain, 1 two register sets, 2 rolling refill.
template <int PIPE>
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
  const HalfView h = lane == 1 ? w.bottom() : w.top();
  if (lane < 2) factor_half(h, i0, i1);            // ONE instruction stream for both halves (two branches would serialise them)
  __syncwarp();
  if (last && lane == 0) factor_middle(w);
  __syncwarp();
}

__device__ __forceinline__ int next_problem_warp(int* counter, int lane) {
  int p = 0;
  if (lane == 0) p = atomicAdd(counter, 1);
  return __shfl_sync(0xffffffffu, p, 0);
}

// ------------------------------------------------------------------------------------------------
// Group execution policy: a group of WPP warps owns one problem and runs it independently of every other group in
// the CTA (no CTA-wide rendezvous).  Stage-parallel phases stride over the 32*WPP lanes of the group; the twisted
// sweeps / factorisation run in two lanes of ONE warp of the group, chosen so that the chain warps of the resident
// groups spread over the four SM sub-partitions (warp w issues on sub-partition w % 4).
// ------------------------------------------------------------------------------------------------
#ifdef MPC_TIMING
__device__ unsigned long long g_tag_cycles[16];
#endif
struct GroupShared {
  int next;
  int anyv[2];
  double red[2][8];
};

template <int WPP>
struct GroupExec {
  int lane, warp;     // warp index within the CTA
  GroupShared* sh;    // this group's slot
  __device__ __forceinline__ int grp() const { return warp / WPP; }
  __device__ __forceinline__ int sub() const { return warp % WPP; }
  __device__ __forceinline__ int gl() const { return sub() * 32 + lane; }
  // groups 0,1,2,3,4,5 (warps 2g, 2g+1) -> chain warps 0,2,5,7,8,10 -> sub-partitions 0,2,1,3,0,2
  __device__ __forceinline__ bool chain_warp() const { return WPP == 1 || sub() == ((grp() >> 1) & 1); }
  __device__ __forceinline__ void group_sync() const {
    if (WPP == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + grp()), "r"(32 * WPP) : "memory");
  }
#ifdef MPC_TIMING     // dev builds only (tools/tag_times.py): cycles between consecutive tags, summed per tag over the launch
  long long t_last = 0; int t_cur = 0;
  __device__ __forceinline__ void tag(int n) {
    const long long now = clock64();
    if (gl() == 0 && t_last) atomicAdd(&g_tag_cycles[t_cur], (unsigned long long)(now - t_last));
    t_last = now; t_cur = n;
  }
#else
  __device__ __forceinline__ void tag(int) {}
#endif
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
      if (lane == 0) for (int i = 0; i < nr; ++i) sh->red[sub()][i] = r[i];
      group_sync();
      for (int i = 0; i < nr; ++i) r[i] = fmax(sh->red[0][i], sh->red[1][i]);
    }
    group_sync();
  }
  template <class Fn> __device__ __forceinline__ int any(int n, Fn f) {
    int a = 0;
    for (int k = gl(); k < n; k += 32 * WPP) a |= f(k);
    a = __any_sync(0xffffffffu, a);
    if (WPP > 1) {
      if (lane == 0) sh->anyv[sub()] = a;
      group_sync();
      a = sh->anyv[0] | sh->anyv[1];
    }
    group_sync();
    return a;
  }
  __device__ __forceinline__ int fetch(int* counter) {
    if (WPP == 1) return next_problem_warp(counter, lane);
    if (gl() == 0) sh->next = atomicAdd(counter, 1);
    group_sync();
    const int b = sh->next;
    group_sync();
    return b;
  }
  // callers reach solve()/factor() right after a phase's group_sync, so the right-hand side / band is visible
  __device__ __forceinline__ void solve(const View& w) {           // polish / refinement solves: compact code
    if (chain_warp()) chain_twisted_lanes<0>(lane < 2, lane == 1, lane ^ 1, w);
    group_sync();
  }
  __device__ __forceinline__ void solve_iter(const View& w) {      // the solve of every ADMM iteration: rolling prefetch
    if (chain_warp()) chain_twisted_lanes<2>(lane < 2, lane == 1, lane ^ 1, w);
    group_sync();
  }
  __device__ __forceinline__ void factor(const View& w) {
    if (chain_warp()) factor_twisted_lanes(lane, w, 0, max(half_top(w.N), half_bot(w.N)), true);
    group_sync();
  }
};

}  // namespace mpc
