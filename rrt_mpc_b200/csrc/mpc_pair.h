// mpc_pair.h — the "pair" form of an ADMM iteration: ONE pass per iteration with a lane per PAIR of stages.
//
// Lane l owns the even stage e = 2l and the odd stage o = 2l + 1 (horizons with N + 1 <= 64: at most 32 pairs, one warp).
// Everything an iteration does outside the two sweeps over the even stages is then local to a lane or one lane away:
//
//   expand   x~_o = t_o - D_o^-1 (E_o x~_e + E_{o+1}' x~_{e+2})                 rows e, e+2 from the sweeps, t_o from the last pass
//   update   A1 of both stages (relaxation, row states, duals)                    needs the inputs of stage e-1   <- lane l-1
//   rhs      A2 of both stages from the values just computed (no reload)          needs d_{e-1} <- lane l-1,  the rate sums of e+2 <- lane l+1
//            t_o = D_o^-1 b_o                                                     (kept in the solution row of o for the next expand)
//   fixup    b'_e = b_e - E_e t_{e-1} - E_{e+1}' t_o                              needs t_{e-1} <- lane l-1
//
// i.e. the update of iteration i and the right-hand side of iteration i+1 are ONE pass over the stage records instead of
// four parity passes (rhs odd / even, update odd / even), the two stages of a lane are two independent instruction
// streams the compiler interleaves (the phases are latency bound: one warp per scheduler), and what crosses lanes are 14
// doubles per lane through warp shuffles instead of a round trip through shared memory and a barrier.
// The arithmetic (operands and order of every sum) is that of admm_update_vals / admm_rhs_vals in mpc_core.h; the
// parts are separate functions so that the host emulation (tests/emu) can run them lane after lane with the exchanged
// values taken from the neighbours' contexts.  The first right-hand side after a (re)factorisation is formed by the
// general parity passes (admm_rhs_stage_oe); the right-hand side formed by the last pass of a solve is not used.
#pragma once

namespace mpc {

MPC_HD int pair_lanes(int N) { return N / 2 + 1; }          // = number of even stages

struct PairCtx {
  double xe[6], xo[6];     // x~ of the lane's even / odd stage
  double xn[6];            // x~ of the next even stage (zeros beyond the horizon)
  double de[4], dd[4];     // d = rho_eq c - y_dyn of the dynamics rows of the even / odd stage, after the update
  double Ge[5], Go[5];     // t0 + t1 of the five soft groups, after the update
  double be[6], bo[6];     // sigma x - q of the relaxed x; be becomes b_e in pair_rhs
  double t[6];             // t_o
};

// part 1: rows from the sweeps, x~ of the odd stage
MPC_HD void pair_expand(const View& w, const Params& p, const IterConst& c, const OEView& oe, int l, PairCtx& cx) {
  const int N = w.N, e = 2 * l, o = e + 1;
  row_load(w.nx(e), cx.xe);
  if (e + 2 <= N) row_load(w.nx(e + 2), cx.xn);
  else {
#pragma unroll
    for (int j = 0; j < 6; ++j) cx.xn[j] = 0.0;
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) cx.xo[j] = 0.0;
  if (o <= N) {
    double di[OE_SYM], t[6], v[6], y[6], u[6];
    row_load(w.nx(o), t);
    cross_mul(w.rec(e) + R_LIN, p.dt, c.rho_eq, c.kap, o < N, cx.xe, v);
    if (o + 1 <= N) {
      cross_mul_t(w.rec(o) + R_LIN, p.dt, c.rho_eq, c.kap, o + 1 < N, cx.xn, y);
#pragma unroll
      for (int j = 0; j < 6; ++j) v[j] += y[j];
    }
    sym_load(oe.dinv + OE_SYM * l, di);
    symv6(di, v, u);
#pragma unroll
    for (int j = 0; j < 6; ++j) cx.xo[j] = t[j] - u[j];
  }
}

// A1 of stage k followed by the stage-local part of A2 from the new values.  xt = x~ of the stage, xn = of stage k+1,
// (ua, ud) = x~ inputs of stage k-1.  `in` and `out` are the same record (or a register copy of its state): every word is
// read before it is written and never read again, so the restrict qualifiers only tell the compiler that the two stages of
// a lane do not alias.  `cst` is where the constants of the stage (R_LIN, R_Q) are read (always the record).
// Leaves the new state in `out` and returns the group sums G[g] = t0 + t1; what A2 further needs of the stage - d (stage_d)
// and sigma x - q (stage_base) - is a function of the new state alone.
// The constants come as `const C&` with C = IterConst or volatile IterConst: the register form reads them from shared memory
// where they are used (volatile: one load per use, never hoisted into long-lived registers); multi-use scalars are copied
// to locals first.
template <class C>
// (inx, outx): where the x / u entries (R_XU) and the dynamics duals (R_YE) are read and written - the same record as
// (in, out) or, in the register form with partial residency, the shared-memory record while (in, out) is the register copy.
MPC_HD void pair_stage(const double* __restrict__ in, double* __restrict__ out, const double* inx, double* outx, const double* cst, double* hdr,
                       const Params& p, const C& c, int N, int k, const double* xt, const double* xn, double ua, double ud, double* G) {
  // BRANCH-FREE over the groups and the terminal stage: the terminal stage has no inputs, no dynamics rows and only the velocity
  // group, but its record has the slots of a regular stage and they hold exact zeros (cold_start_stage), which are a fixed point of
  // the update (z = clip(0) = 0 inside the bounds, n = 0 + alpha (0 - 0) = 0, t = 0, G = 0): computing them costs the one terminal
  // lane nothing extra and lets the compiler schedule the five groups, the dynamics rows and the relaxation as ONE basic block
  // (with a branch per group the phases were ~12 dependent fp64 operations at a time: `wait` 1.5 stall cycles per instruction).
  // Only the dynamics duals and the input entries need a select.
  const bool reg = k < N;
  const double c_ra = c.ra, c_alpha = c.alpha, c_rho = c.rho, c_sigma = c.sigma;
  {
    const double* lin = cst + R_LIN;
    const double z0 = xn[0] - (xt[0] + lin[0] * xt[2] + lin[1] * xt[3]);
    const double z1 = xn[1] - (xt[1] + lin[2] * xt[2] + lin[3] * xt[3]);
    const double z2 = xn[2] - (xt[2] + lin[4] * xt[5]);
    const double z3 = xn[3] - (xt[3] + p.dt * xt[4]);
    const double y0 = fma(c_ra, z0 - lin[5], inx[R_YE + 0]), y1 = fma(c_ra, z1 - lin[6], inx[R_YE + 1]);
    const double y2 = fma(c_ra, z2, inx[R_YE + 2]), y3 = fma(c_ra, z3, inx[R_YE + 3]);
    outx[R_YE + 0] = reg ? y0 : 0.0; outx[R_YE + 1] = reg ? y1 : 0.0; outx[R_YE + 2] = reg ? y2 : 0.0; outx[R_YE + 3] = reg ? y3 : 0.0;
  }
  if (k == 0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) hdr[H_YI + r] = fma(c_ra, xt[r] - hdr[H_X0 + r], hdr[H_YI + r]);
  }
  const double off0 = k == 0 ? c.up0 : 0.0, off1 = k == 0 ? c.up1 : 0.0;
  const double gt[5] = {xt[3], reg ? xt[4] : 0.0, reg ? xt[5] : 0.0, reg ? xt[4] - ua : 0.0, reg ? xt[5] - ud : 0.0};
  const double offs[5] = {0.0, 0.0, 0.0, off0, off1};
#pragma unroll
  for (int g = 0; g < 5; ++g) {
    const double st = in[R_ST + g], sv = in[R_S + g];
    const double v0 = in[R_V + 3 * g], v1 = in[R_V + 3 * g + 1], v2 = in[R_V + 3 * g + 2];
    const double hi = c.hi[g] + offs[g], lo = c.lo[g] + offs[g], ms = c.mssinv[g];
    const double z0 = dmin2(v0, hi), z1 = dmax2(v1, lo), z2 = dmax2(v2, 0.0);
    const double n0 = fma(c_alpha, (gt[g] - st) - z0, v0);
    const double n1 = fma(c_alpha, (gt[g] + st) - z1, v1);
    const double n2 = fma(c_alpha, st - z2, v2);
    const double sn = fma(c_alpha, st - sv, sv);
    out[R_V + 3 * g] = n0; out[R_V + 3 * g + 1] = n1; out[R_V + 3 * g + 2] = n2;
    out[R_S + g] = sn;
    const double y0 = dmin2(n0, hi), y1 = dmax2(n1, lo), y2 = dmax2(n2, 0.0);
    const double t0 = c_rho * (y0 + (y0 - n0)), t1 = c_rho * (y1 + (y1 - n1)), t2 = c_rho * (y2 + (y2 - n2));
    out[R_ST + g] = (c_sigma * sn + ((t1 - t0) + t2)) * ms;
    G[g] = t0 + t1;
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const double xo = inx[R_XU + j];
    const double xw = fma(c_alpha, xt[j] - xo, xo);
    outx[R_XU + j] = (j < 4 || reg) ? xw : 0.0;
  }
}
// d = rho_eq c - y of the dynamics rows of stage k (zeros for the terminal stage), from the state `st`
template <class C>
MPC_HD void stage_d(const double* st, const double* cst, const C& c, bool reg, double* d) {
  const double re = c.rho_eq;
  d[0] = reg ? re * cst[R_LIN + 5] - st[R_YE + 0] : 0.0;
  d[1] = reg ? re * cst[R_LIN + 6] - st[R_YE + 1] : 0.0;
  d[2] = reg ? -st[R_YE + 2] : 0.0;
  d[3] = reg ? -st[R_YE + 3] : 0.0;
}
// sigma x - q of stage k
template <class C>
MPC_HD void stage_base(const double* st, const double* cst, const C& c, bool reg, double* base) {
  const double sg = c.sigma;
#pragma unroll
  for (int j = 0; j < 6; ++j) base[j] = (j < 4 || reg) ? sg * st[R_XU + j] - (j < 4 ? cst[R_Q + j] : 0.0) : 0.0;
}

// part 2: both stages of the lane.  (ua, ud): x~ inputs of stage e-1 (lane l-1; zeros for l = 0)
MPC_HD void pair_update(const View& w, const Params& p, const IterConst& c, int l, PairCtx& cx, double ua, double ud) {
  const int N = w.N, e = 2 * l, o = e + 1;
  const bool ib = e > 0 && e < N;
  double* re = w.rec(e);
  pair_stage(re, re, re, re, re, w.hdr(), p, c, N, e, cx.xe, cx.xo, ib ? ua : 0.0, ib ? ud : 0.0, cx.Ge);
  stage_d(re, re, c, e < N, cx.de);
  stage_base(re, re, c, e < N, cx.be);
  if (o <= N) {
    double* ro = w.rec(o);
    pair_stage(ro, ro, ro, ro, ro, w.hdr(), p, c, N, o, cx.xo, cx.xn, o < N ? cx.xe[4] : 0.0, o < N ? cx.xe[5] : 0.0, cx.Go);
    stage_d(ro, ro, c, o < N, cx.dd);
    stage_base(ro, ro, c, o < N, cx.bo);
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r) cx.dd[r] = 0.0;
#pragma unroll
    for (int g = 0; g < 5; ++g) cx.Go[g] = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) cx.bo[j] = 0.0;
  }
}

// right-hand side of stage k from its own parts and its neighbours': dprev = d of stage k-1 (k = 0: the initial-state
// rows), rnext = rate-group sums (G[3], G[4]) of stage k+1
MPC_HD void pair_assemble(const double* lin, const Params& p, int N, int k, const double* dprev, const double* d, const double* G,
                          const double* rnext, const double* base, double* val) {
  // branch-free: for the terminal stage d = 0 (stage_d), G[1..4] = 0 and lin = 0, so the terms below vanish exactly
  const bool reg = k < N;
  double out[6] = {dprev[0], dprev[1], dprev[2], dprev[3], 0.0, 0.0};
  out[3] += G[0];
  out[0] -= d[0];
  out[1] -= d[1];
  out[2] -= lin[0] * d[0] + lin[2] * d[1] + d[2];
  out[3] -= lin[1] * d[0] + lin[3] * d[1] + d[3];
  out[4] = G[1] + G[3] - p.dt * d[3];
  out[5] = G[2] + G[4] - lin[4] * d[2];
  out[4] -= (k + 1 < N) ? rnext[0] : 0.0;
  out[5] -= (k + 1 < N) ? rnext[1] : 0.0;
#pragma unroll
  for (int j = 0; j < 6; ++j) val[j] = (j < 4 || reg) ? base[j] + out[j] : 0.0;
}

// part 3: b_e (kept in cx.be), b_o -> t_o (stored in the row of o, kept in cx.t).  dprev: d of stage e-1 (lane l-1),
// rnext: (G[3], G[4]) of stage e+2 (lane l+1)
MPC_HD void pair_rhs(const View& w, const Params& p, const IterConst& c, const OEView& oe, int l, PairCtx& cx, const double* dprev, const double* rnext) {
  const int N = w.N, e = 2 * l, o = e + 1;
  double dp[4];
  if (l == 0) {
    const double* h = w.hdr();
#pragma unroll
    for (int r = 0; r < 4; ++r) dp[r] = c.rho_eq * h[H_X0 + r] - h[H_YI + r];
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r) dp[r] = dprev[r];
  }
  double val[6];
  pair_assemble(w.rec(e) + R_LIN, p, N, e, dp, cx.de, cx.Ge, cx.Go + 3, cx.be, val);
#pragma unroll
  for (int j = 0; j < 6; ++j) cx.be[j] = val[j];
#pragma unroll
  for (int j = 0; j < 6; ++j) cx.t[j] = 0.0;
  if (o <= N) {
    double di[OE_SYM];
    pair_assemble(w.rec(o) + R_LIN, p, N, o, cx.de, cx.dd, cx.Go, rnext, cx.bo, val);
    sym_load(oe.dinv + OE_SYM * l, di);
    symv6(di, val, cx.t);
    row_store(w.nx(o), cx.t);
  }
}

// part 4: b'_e = b_e - E_e t_{e-1} - E_{e+1}' t_o.  tprev: t of stage e-1 (lane l-1)
MPC_HD void pair_fixup(const View& w, const Params& p, const IterConst& c, int l, PairCtx& cx, const double* tprev) {
  const int N = w.N, e = 2 * l;
  double y[6];
  if (e >= 1) {
    cross_mul(w.rec(e - 1) + R_LIN, p.dt, c.rho_eq, c.kap, e < N, tprev, y);
#pragma unroll
    for (int j = 0; j < 6; ++j) cx.be[j] -= y[j];
  }
  if (e + 1 <= N) {
    cross_mul_t(w.rec(e) + R_LIN, p.dt, c.rho_eq, c.kap, e + 1 < N, cx.t, y);
#pragma unroll
    for (int j = 0; j < 6; ++j) cx.be[j] -= y[j];
  }
  row_store(w.nx(e), cx.be);
}

// the whole pass, lane after lane (host emulation; `reverse` runs the lanes of every part in the other order: the parts
// only communicate through the contexts, so the result must not depend on it)
MPC_HD void pair_pass_seq(const View& w, const Params& p, const IterConst& c, const OEView& oe, PairCtx* cx, bool reverse) {
  const int L = pair_lanes(w.N);
  const double zero[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < L; ++i) { const int l = reverse ? L - 1 - i : i; pair_expand(w, p, c, oe, l, cx[l]); }
  for (int i = 0; i < L; ++i) {
    const int l = reverse ? L - 1 - i : i;
    pair_update(w, p, c, l, cx[l], l > 0 ? cx[l - 1].xo[4] : 0.0, l > 0 ? cx[l - 1].xo[5] : 0.0);
  }
  for (int i = 0; i < L; ++i) {
    const int l = reverse ? L - 1 - i : i;
    pair_rhs(w, p, c, oe, l, cx[l], l > 0 ? cx[l - 1].dd : zero, l + 1 < L ? cx[l + 1].Ge + 3 : zero);
  }
  for (int i = 0; i < L; ++i) { const int l = reverse ? L - 1 - i : i; pair_fixup(w, p, c, l, cx[l], l > 0 ? cx[l - 1].t : zero); }
}

}  // namespace mpc
