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

// ------------------------------------------------------------------------------------------------
// Sweeps of the odd-even block solve (mpc_oe.h), one lane per ROW of a half: lanes 0..5 the top half, 6..11 the bottom
// half, one instruction stream.  Called by all 32 lanes of the chain warp (lanes >= 12 only keep the warp converged).
// Per step and lane: 3 x LDS.128 (block row, requested one step ahead) + 6 fmas + one 64-bit store of its element of the
// new vector, __syncwarp, 3 x LDS.128 to read the whole vector back.  The loop is branch-free: both halves run the trip
// count of the longer one, a half that is done keeps computing on in-bounds words of the workspace and stores nothing
// (predicated), so the warp never diverges inside the dependent chain.  Shared-memory accesses are explicit PTX on 32-bit
// shared addresses (through generic pointers every access pays an address-space conversion), two register sets, unrolled
// by two.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lds_row(unsigned a, double* r) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r[0]), "=d"(r[1]) : "r"(a));
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(r[2]), "=d"(r[3]) : "r"(a));
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+32];" : "=d"(r[4]), "=d"(r[5]) : "r"(a));
}
__device__ __forceinline__ void lds_col(unsigned a, double* r) {       // six doubles 48 bytes apart (a block column)
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r[0]) : "r"(a));
  asm volatile("ld.shared.f64 %0, [%1+48];" : "=d"(r[1]) : "r"(a));
  asm volatile("ld.shared.f64 %0, [%1+96];" : "=d"(r[2]) : "r"(a));
  asm volatile("ld.shared.f64 %0, [%1+144];" : "=d"(r[3]) : "r"(a));
  asm volatile("ld.shared.f64 %0, [%1+192];" : "=d"(r[4]) : "r"(a));
  asm volatile("ld.shared.f64 %0, [%1+240];" : "=d"(r[5]) : "r"(a));
}
__device__ __forceinline__ double lds_f64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f64_if(bool p, unsigned a, double v) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.shared.f64 [%1], %2; }" ::"r"((unsigned)p), "r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void oe_forward_lanes(int lane, const View& w, const OEView& oe) {
  const bool act = lane < 12;
  const bool bottom = (lane / 6) & 1;
  const int r = lane % 6;
  const OEHalf h = oe_half(w, oe, bottom);
  const int cnt = h.cnt, cmax = oe.jm > oe.nb ? oe.jm : oe.nb;     // uniform trip count
  const int xs = h.xstep * 8;                                      // bytes between consecutive local rows (+-48)
  const unsigned xlast = smem_u32(h.xlast) + 8 * r;
  unsigned xrow = smem_u32(h.x0) + xs;                             // row of local stage i (i = 1)
  unsigned gp = smem_u32(h.g0) + 48 * r;                           // row r of the block of step i
  double a[6], gA[6], gB[6], nbA, nbB, y;
  lds_row(smem_u32(h.x0), a);                                      // y_0 = b'_0
  double ymid = lds_f64(xlast);                                    // the half's term for the middle if it has no step
  lds_row(gp, gA); nbA = lds_f64(cnt == 1 ? xlast : xrow + 8 * r);
  for (int i = 1; i <= cmax; i += 2) {
    lds_row(gp + 288, gB); nbB = lds_f64(i + 1 == cnt ? xlast : xrow + xs + 8 * r);          // step i+1
    y = oe_row_dot(gA, nbA, a);
    ymid = (i == cnt) ? y : ymid;
    sts_f64_if(act && i < cnt, xrow + 8 * r, y);
    __syncwarp();
    lds_row(xrow, a);
    if (i + 1 > cmax) break;
    lds_row(gp + 576, gA); nbA = lds_f64(i + 2 == cnt ? xlast : xrow + 2 * xs + 8 * r);      // step i+2
    y = oe_row_dot(gB, nbB, a);
    ymid = (i + 1 == cnt) ? y : ymid;
    sts_f64_if(act && i + 1 < cnt, xrow + xs + 8 * r, y);
    __syncwarp();
    lds_row(xrow + xs, a);
    gp += 576; xrow += 2 * xs;
  }
  // middle stage: x_m = S'_m^-1 (top term + bottom term); lanes r and r + 6 hold element r of the two terms
  const double yo = __shfl_sync(0xffffffffu, ymid, bottom ? lane - 6 : (lane + 6) & 31);
  double* mrow = w.nx(2 * oe.jm);
  if (lane < 6) mrow[r] = ymid + yo;
  __syncwarp();
  double s[6];
  row_load(mrow, s);
  const double xm = oe_sym_row(oe.sinv + OE_SYM * oe.jm, r, s);
  __syncwarp();
  if (lane < 6) mrow[r] = xm;
  __syncwarp();
}
__device__ __forceinline__ void oe_backward_lanes(int lane, const View& w, const OEView& oe) {
  const bool act = lane < 12;
  const bool bottom = (lane / 6) & 1;
  const int c = lane % 6;
  const OEHalf h = oe_half(w, oe, bottom);
  const int cmax = oe.jm > oe.nb ? oe.jm : oe.nb;
  const int xs = h.xstep * 8;
  int i = h.cnt - 1;                                               // local stage of this half at the current step (< 0: done)
  unsigned xrow = smem_u32(h.x0) + xs * i;                         // row of local stage i
  unsigned gp = smem_u32(h.g0) + 288 * i + 8 * c;                  // column c of the block of step i+1
  double a[6], gA[6], gB[6], zA, zB;
  lds_row(smem_u32(w.nx(2 * oe.jm)), a);                           // x_m
  lds_col(gp, gA); zA = lds_f64(xrow + 8 * c);
  for (int s = 0; s < cmax; s += 2, i -= 2) {
    lds_col(gp - 288, gB); zB = lds_f64(xrow - xs + 8 * c);        // local stage i-1
    sts_f64_if(act && i >= 0, xrow + 8 * c, oe_row_dot(gA, zA, a));
    __syncwarp();
    lds_row(xrow, a);
    if (s + 1 >= cmax) break;
    lds_col(gp - 576, gA); zA = lds_f64(xrow - 2 * xs + 8 * c);    // local stage i-2
    sts_f64_if(act && i - 1 >= 0, xrow - xs + 8 * c, oe_row_dot(gB, zB, a));
    __syncwarp();
    lds_row(xrow - xs, a);
    gp -= 576; xrow -= 2 * xs;
  }
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
  long long t_last = 0; int t_cur = 0; unsigned long long* tags = nullptr;
  __device__ __forceinline__ void tag(int n) {
    const long long now = clock64();
    if (gl() == 0 && t_last && tags) atomicAdd(&tags[t_cur], (unsigned long long)(now - t_last));
    t_last = now; t_cur = n;
  }
#else
  __device__ __forceinline__ void tag(int) {}
#endif
  template <class Fn> __device__ __forceinline__ void stages(int n, Fn f) {
    for (int k = gl(); k < n; k += 32 * WPP) f(k);
    group_sync();
  }
  // stages of one parity (par = 1: odd, 0: even), lanes over them
  template <class Fn> __device__ __forceinline__ void stages_par(int n, int par, Fn f) {
    for (int k = 2 * gl() + par; k < n; k += 64 * WPP) f(k);
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
  // ---- odd-even block solve of the ADMM iterations (mpc_oe.h): stage-parallel parts with lanes over stages, the two
  // halves of the reduced system in lanes 0 (top) and 1 (bottom) of the group's chain warp, one instruction stream
  __device__ __forceinline__ void oe_factor(const View& w, const Params& p, const Mode& m, const OEView& oe) {
    stages_par(w.N + 1, 1, [&](int k) { oe_factor_odd(w, p, m, oe, k); });
    stages_par(w.N + 1, 0, [&](int k) { oe_factor_even(w, p, m, oe, k); });
    if (chain_warp()) {
      double U[21], Uo[21];
#pragma unroll
      for (int t = 0; t < 21; ++t) U[t] = 0.0;
      if (lane < 2) oe_factor_half(oe_half(w, oe, lane == 1), U);
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 21; ++t) Uo[t] = __shfl_sync(0xffffffffu, U[t], 1);
      if (lane == 0) oe_factor_middle(oe, U, Uo);
    }
    group_sync();
  }
  __device__ __forceinline__ void oe_forward(const View& w, const OEView& oe) {
    if (chain_warp()) oe_forward_lanes(lane, w, oe);
    group_sync();
  }
  __device__ __forceinline__ void oe_backward(const View& w, const OEView& oe) {
    if (chain_warp()) oe_backward_lanes(lane, w, oe);
    group_sync();
  }
  __device__ __forceinline__ void factor(const View& w) {
    if (chain_warp()) factor_twisted_lanes(lane, w, 0, max(half_top(w.N), half_bot(w.N)), true);
    group_sync();
  }
};

}  // namespace mpc
