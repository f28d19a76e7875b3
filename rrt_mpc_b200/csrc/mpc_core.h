// mpc_core.h — per-problem MPC tracking-step mathematics for libcudampc (sm_100a).
//
// One tracking problem = the horizon QP of /root/reference/src/control/mpc_controller.py:47-117
// (n = 11N+5 variables, m = 19N+7 rows) solved by an OSQP-equivalent ADMM
// (mpc_controller.py:119-132 settings) plus polish.  Everything here is written as
// "stage-parallel" routines: routine(k) touches only stage k's record (and reads its neighbours'),
// so a warp runs them with lanes striding over stages, separated by barriers.  The only sequential
// parts are the banded LDL' factorisation and the two triangular sweeps ("the chain").
//
// The same header compiles for the host (tests/emu, a lane-by-lane emulation used to debug the
// index logic where no GPU is available).  The product path is the CUDA kernel in cudampc.cu.
//
// Formulation notes:
//  * unscaled problem (OSQP `scaling=0`): A's structural entries stay +-1, bounds stay shared constants;
//  * merged row state v = z + y/rho for inequality rows (z = clip(v), y = rho (v - z)); equality rows
//    keep z = b and store y;
//  * slack columns are eliminated inside the linear solve (they are leaves of the elimination tree),
//    leaving a banded SPD system of order 6N+4 and half-bandwidth 6 in (X,Y,psi,v,a,delta) stage order;
//  * polish = the same machinery with per-row weights {0, 1/delta} (reduced form of OSQP's polish KKT).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define MPC_HD __host__ __device__ __forceinline__
#else
#define MPC_HD inline
#endif

namespace mpc {

// ----------------------------------------------------------------------------------------------
// Parameters / settings as the kernel sees them (filled by the C-ABI from cudampc_params/settings)
// ----------------------------------------------------------------------------------------------
struct Params {
  double L, dt;
  double pq[4][4], pr[2][2], pqn[4][4];   // P blocks of the objective: Q + Q', R + R', Q_N + Q_N' (symmetric; cp.quad_form
                                          // accepts any PSD matrix, mpc_controller.py:74-75,112)
  double u_lo[2], u_hi[2];
  double v_lo, v_hi;
  double du_lo[2], du_hi[2];
  double w_v, w_u, w_du;             // slack weights (P entries are 2*w)
  int N;
};

struct Settings {
  double eps_abs, eps_rel;
  double rho0, alpha, sigma;
  double adaptive_rho_tolerance, rho_eq_factor, rho_min, rho_max, delta;
  int max_iter, check_termination, adaptive_rho, adaptive_rho_interval;
  int polish_passes, polish_refine_iter, warm_start;
  int polish_retry;   // resume ADMM at a tighter internal tolerance this many times if the polish is rejected
  int early_polish;        // try the polish at termination checks whose guessed active set repeated (0 = OSQP)
  int early_polish_start;  // first iteration at which an early polish may be tried
};

enum { STATUS_SOLVED = 1, STATUS_SOLVED_INACCURATE = 2, STATUS_MAX_ITER = -2, STATUS_UNSOLVED = -10 };

// ----------------------------------------------------------------------------------------------
// Workspace layout (doubles) of one problem
// ----------------------------------------------------------------------------------------------
// stage record (stride SR, odd => conflict-free for lanes striding over stages)
enum {
  R_XU = 0,    // X,Y,psi,v,a,delta        (terminal stage: first 4 only)
  R_S = 6,     // sv, su0, su1, sdu0, sdu1 (terminal: sv only)
  R_V = 11,    // merged row state, [3*g + r], g: 0 v,1 u0,2 u1,3 du0,4 du1 ; r: 0 hi,1 lo,2 s>=0
  R_YE = 26,   // duals of the 4 dynamics rows k -> k+1
  R_LIN = 30,  // a02,a03,a12,a13,b21,c0,c1
  R_Q = 37,    // linear cost of x_k  (-2 Q ref_k)
  R_ST = 41,   // s-tilde / reduced slack rhs scratch (5)
  R_PAD = 46,
  SR = 47
};
enum { H_X0 = 0, H_UPREV = 4, H_YI = 6, H_ACT = 10, HDR_FIXED = 10 };

MPC_HD int hdr_size(int N) { return HDR_FIXED + ((N + 2 + 1) >> 1); }
MPC_HD int nband(int N) { return 6 * (N + 1); }   // terminal stage carries two dummy unknowns (identity rows)

// Factor storage: one block of BLK doubles per stage (plus one padding block), 16-byte aligned so that a
// whole block is staged with 128-bit loads:
//   [0..14]  in-stage strictly lower L_k[j][jp], jp < j          -> IA(j,jp)
//   [15..20] diagonal (before factorisation) / 1/D (after)       -> ID(j)
//   [22..42] cross block L[(k,j)][(k-1,jp)], jp >= j (band = 6)  -> IC(j,jp)
// Both sweeps of stage k need exactly: in-stage part of block k plus one cross part (forward: block k+1,
// backward: block k).
enum { BLK = 44 };
#define MPC_IA(j, jp) ((j) * ((j) - 1) / 2 + (jp))
#define MPC_ID(j) (15 + (j))
#define MPC_IC(j, jp) (22 + 6 * (j) - (j) * ((j) - 1) / 2 + ((jp) - (j)))

// Twisted ("burn at both ends") organisation of the banded system: stages 0..m are eliminated top-down, stages
// N..m+1 bottom-up, both ending at the middle stage m = N/2.  The bottom half is STORED in reversed order
// (stage N-i at local index i, in-stage index 5-j): the reversed system has exactly the same band structure, so
// one instruction stream sweeps a top half in one lane and a bottom half in another, halving the chain length.
MPC_HD int mid_stage(int N) { return N / 2; }
MPC_HD int half_top(int N) { return mid_stage(N) + 1; }          // local stages of the top half (last = middle = border)
MPC_HD int half_bot(int N) { return N - mid_stage(N) + 1; }      // local stages of the bottom half (last = middle = border)
// rhs/solution rows: BXS = 6 doubles per local stage, rows -1 .. H of each half, 16-byte aligned so that the chain moves a
// row with three 128-bit accesses (shared-memory instructions, not bytes, are what the chain warp pays for)
enum { BXS = 6 };
MPC_HD int bx_offset(int N) { return (hdr_size(N) + (N + 1) * SR + 1) & ~1; }
MPC_HD int bx_doubles(int N) { return BXS * (half_top(N) + 2) + BXS * (half_bot(N) + 2) + 2; }
MPC_HD int band_offset(int N) { return bx_offset(N) + bx_doubles(N); }   // even => 16-byte aligned blocks
enum { PAD_MAX = 16 };                     // room for the bank-conflict pads of the bottom half (fpad, xpad)
MPC_HD int footprint(int N) {
  int f = band_offset(N) + BLK * (half_top(N) + 1) + BLK * (half_bot(N) + 1) + 22 + PAD_MAX + 2;
  while ((f & 3) != 2) ++f;               // F = 2 (mod 4): 16-byte aligned, and an odd number of 16-byte columns per problem
  return f;
}
// Pads that make the chain warp's accesses bank-conflict-free.  General form: lanes 0..P-1 read top halves and lanes
// P..2P-1 bottom halves of P problems F doubles apart (what tools/microbench/chain_bench.cu sweeps); the kernels use
// P = 1: lane 0 = top half, lane 1 = bottom half of the warp's own problem.  The bottom half's factor blocks are shifted
// by P 16-byte columns (mod 8) relative to the top half's, its rhs/solution rows by an odd number of 16-byte columns.
MPC_HD void layout_pads(int N, int P, int& fpad, int& xpad) {
  const int F = footprint(N);
  const int q = (F / 2) & 7;                                   // 16-byte columns per problem (odd)
  const int nat = ((BLK * (half_top(N) + 1)) / 2) & 7;         // natural column offset of the bottom half
  const int want = (P * q) & 7;
  fpad = 2 * ((want - nat) & 7);
  xpad = (((BXS * (half_top(N) + 2)) / 2) & 1) ? 0 : 2;         // bottom rows an odd number of 16-byte columns from the top rows
}
MPC_HD void layout_pads(int N, int& fpad, int& xpad) { layout_pads(N, 1, fpad, xpad); }   // one problem per chain warp
// warm-start state kept in HBM between calls: per stage xu(6) s(5) v(15) ye(4), + yi(4) + rho
MPC_HD int warm_size(int N) { return 30 * (N + 1) + 5; }

// one half of the twisted system as the chain code sees it (local stage index i = 0 .. H-1, border = H-1)
struct HalfView {
  double* bx0; double* blk0; int H;
  MPC_HD double* bx(int i) const { return bx0 + BXS * i; }       // i = -1 .. H
  MPC_HD double* blk(int i) const { return blk0 + BLK * i; }     // i = 0 .. H
};

struct View {
  double* base;
  int N;
  int fpad, xpad;    // bank-conflict pads of the bottom half (layout_pads)
  MPC_HD double* hdr() const { return base; }
  MPC_HD int* act() const { return reinterpret_cast<int*>(base + H_ACT); }  // [N+1] stage masks + [N+1]=init rows
  // records are stored even stages first, then the odd ones: the ADMM phases run one parity at a time, and lanes striding
  // over the stages of one parity then stay SR (odd) doubles apart - conflict-free
  MPC_HD double* rec(int k) const { return base + hdr_size(N) + ((k & 1) ? (N / 2 + 1) + (k >> 1) : (k >> 1)) * SR; }
  // natural (untwisted) rhs / solution rows of the odd-even block solve: row k of the 6(N+1)-vector, same parity split;
  // row N+1 is a row of zeros
  MPC_HD double* nx(int k) const { return bx_base() + BXS * ((k & 1) ? (N / 2 + 1) + (k >> 1) : (k >> 1)); }
  MPC_HD double* nx_zero() const { return bx_base() + BXS * (N + 1); }
  MPC_HD double* bx_base() const { return base + bx_offset(N); }
  MPC_HD double* scratch() const { return bx_base(); }                       // >= N+1 doubles, free before the first solve
  MPC_HD HalfView top() const { return HalfView{bx_base() + BXS, base + band_offset(N), half_top(N)}; }
  MPC_HD HalfView bottom() const {
    return HalfView{bx_base() + BXS * (half_top(N) + 2) + BXS + xpad, base + band_offset(N) + BLK * (half_top(N) + 1) + fpad, half_bot(N)};
  }
  MPC_HD double* mid() const { return base + band_offset(N) + BLK * (half_top(N) + 1) + BLK * (half_bot(N) + 1) + PAD_MAX; }
  // element (k, j) of the right-hand side / solution vector in the twisted storage
  MPC_HD double bx_get(int k, int j) const {
    const int m = mid_stage(N);
    return k <= m ? top().bx(k)[j] : bottom().bx(N - k)[5 - j];
  }
  MPC_HD void bx_set(int k, int j, double v) const {
    const int m = mid_stage(N);
    if (k <= m) top().bx(k)[j] = v; else bottom().bx(N - k)[5 - j] = v;
    if (k == m) bottom().bx(N - m)[5 - j] = 0.0;      // the bottom half's copy of the middle stage carries no rhs
  }
};

struct alignas(16) D2 { double x, y; };   // 128-bit shared-memory access (blocks are 16-byte aligned)

MPC_HD double clipd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }
MPC_HD double dmax(double a, double b) { return a > b ? a : b; }

// bounds of soft group g of stage k
MPC_HD void group_bounds(const Params& p, const double* hdr, int k, int g, double& lo, double& hi) {
  if (g == 0) { lo = p.v_lo; hi = p.v_hi; }
  else if (g <= 2) { lo = p.u_lo[g - 1]; hi = p.u_hi[g - 1]; }
  else {
    double off = (k == 0) ? hdr[H_UPREV + g - 3] : 0.0;
    lo = p.du_lo[g - 3] + off; hi = p.du_hi[g - 3] + off;
  }
}
MPC_HD double group_ps(const Params& p, int g) { return 2.0 * (g == 0 ? p.w_v : (g <= 2 ? p.w_u : p.w_du)); }
MPC_HD int ngroups(int N, int k) { return k < N ? 5 : 1; }

// ----------------------------------------------------------------------------------------------
// Linear-system mode: ADMM (uniform rho) or polish (activity bits, weights {0,1/delta})
// ----------------------------------------------------------------------------------------------
struct Mode {
  int polish;        // 0: ADMM, 1: polish
  double rho, rho_eq, reg;   // reg = sigma (ADMM) or delta (polish)
  double inv_delta;
};
MPC_HD Mode admm_mode(double rho, const Settings& s) {
  Mode m; m.polish = 0; m.rho = rho; m.rho_eq = s.rho_eq_factor * rho; m.reg = s.sigma; m.inv_delta = 0.0; return m;
}
MPC_HD Mode polish_mode(const Settings& s) {
  Mode m; m.polish = 1; m.rho = 0.0; m.rho_eq = 0.0; m.reg = s.delta; m.inv_delta = 1.0 / s.delta; return m;
}

struct GroupCoef { double kappa, mss_inv, csg; };
// activity bits of a group: bit0 hi-row, bit1 lo-row, bit2 s>=0 row
MPC_HD GroupCoef group_coef(const Mode& m, double ps, int bits) {
  GroupCoef c;
  if (!m.polish) {
    c.kappa = 2.0 * m.rho; c.mss_inv = 1.0 / (ps + m.reg + 3.0 * m.rho); c.csg = 0.0;
  } else {
    double a1 = (bits & 1) ? 1.0 : 0.0, a2 = (bits & 2) ? 1.0 : 0.0, a3 = (bits & 4) ? 1.0 : 0.0;
    double d = m.reg;
    double mssd = (ps + d) * d + (a1 + a2 + a3);          // m_ss * delta
    c.mss_inv = d / mssd;
    c.csg = (a2 - a1) * m.inv_delta;
    c.kappa = ((a1 + a2) * ((ps + d) * d + a3) + 4.0 * a1 * a2) / (d * mssd);   // cancellation-free Schur term
  }
  return c;
}
MPC_HD double eq_weight(const Mode& m, int active) { return m.polish ? (active ? m.inv_delta : 0.0) : m.rho_eq; }

// ----------------------------------------------------------------------------------------------
// np.unwrap over the window yaw column (mpc_controller.py:59-60), sequential, lane 0
// ----------------------------------------------------------------------------------------------
MPC_HD double np_mod(double a, double b) {
  double r = fmod(a, b);
  if (r != 0.0 && ((b < 0.0) != (r < 0.0))) r += b;
  return r;
}
// in: yaw[k] at stride `stride`; out: unwrapped values written to out[k] at stride ostride
MPC_HD void unwrap_yaw(const double* yaw, int stride, int count, double* out, int ostride) {
  const double PI = 3.141592653589793;
  double cum = 0.0, prev = yaw[0];
  out[0] = prev;
  for (int k = 1; k < count; ++k) {
    double cur = yaw[(size_t)k * stride];
    double dd = cur - prev;
    double ddmod = np_mod(dd + PI, 2.0 * PI) - PI;
    if (ddmod == -PI && dd > 0.0) ddmod = PI;
    double corr = ddmod - dd;
    if (fabs(dd) < PI) corr = 0.0;
    cum += corr;
    out[(size_t)k * ostride] = cur + cum;
    prev = cur;
  }
}

// ----------------------------------------------------------------------------------------------
// Linearisation of stage k (vehicle_model.py:24-45 as called from mpc_controller.py:65-70,108-109)
//   linearised at the unwrapped reference row max(k-1,0), ulin = 0.
//   lin7 = {a02,a03,a12,a13,b21,c0,c1};  A[2][3] = dt/L*tan(0) = 0, c2 = c3 = 0 exactly.
// ----------------------------------------------------------------------------------------------
MPC_HD void linearize_point(const Params& p, double X, double Y, double yaw, double v, double* lin7) {
  double c = cos(yaw), s = sin(yaw);
  double sec2 = 1.0 / (1.0 * 1.0 + 1e-9);            // cos(0)^2 + 1e-9  (vehicle_model.py:31)
  double a02 = -p.dt * v * s, a03 = p.dt * c, a12 = p.dt * v * c, a13 = p.dt * s;
  double b21 = p.dt * (v / p.L) * sec2;
  double fx0 = X + p.dt * v * cos(yaw + 0.0), fx1 = Y + p.dt * v * sin(yaw + 0.0);
  // c_aff = fx - A @ xlin  (row dot in index order, zeros included as exact no-ops)
  double ax0 = X + a02 * yaw + a03 * v;
  double ax1 = Y + a12 * yaw + a13 * v;
  lin7[0] = a02; lin7[1] = a03; lin7[2] = a12; lin7[3] = a13; lin7[4] = b21;
  lin7[5] = fx0 - ax0; lin7[6] = fx1 - ax1;
}

// ----------------------------------------------------------------------------------------------
// Row-value providers.  A "row provider" RP answers, for the rows owned by stage k:
//   dyn(k, r)  r=0..3   value attached to dynamics row r of stage k (k < N)
//   init(r)             value attached to the X_0 = x0 rows
//   grp(k, g)           the pair-combined value for the (x,u) part of soft group g of stage k
// gather_xu() forms  sum_i A_ij * value_i  for the six (x,u) unknowns of stage k.
// ----------------------------------------------------------------------------------------------
template <class RP>
MPC_HD void gather_xu(const View& w, const Params& p, int k, const RP& rp, double* out) {
  const int N = w.N;
  for (int j = 0; j < 6; ++j) out[j] = 0.0;
  if (k >= 1) { for (int j = 0; j < 4; ++j) out[j] += rp.dyn(k - 1, j); }
  else { for (int j = 0; j < 4; ++j) out[j] += rp.init(j); }
  out[3] += rp.grp(k, 0);
  if (k < N) {
    const double* lin = w.rec(k) + R_LIN;
    double d0 = rp.dyn(k, 0), d1 = rp.dyn(k, 1), d2 = rp.dyn(k, 2), d3 = rp.dyn(k, 3);
    out[0] -= d0;
    out[1] -= d1;
    out[2] -= lin[0] * d0 + lin[2] * d1 + d2;
    out[3] -= lin[1] * d0 + lin[3] * d1 + d3;
    out[4] -= p.dt * d3;
    out[5] -= lin[4] * d2;
    for (int i = 0; i < 2; ++i) {
      out[4 + i] += rp.grp(k, 1 + i) + rp.grp(k, 3 + i);
      if (k + 1 < N) out[4 + i] -= rp.grp(k + 1, 3 + i);
    }
  }
}

// (A x) for the dynamics rows of stage k from a vector accessor XV(k', j)
template <class XV>
MPC_HD void dyn_rows(const View& w, const Params& p, int k, const XV& xv, double* z) {
  const double* lin = w.rec(k) + R_LIN;
  double X = xv(k, 0), Y = xv(k, 1), ps = xv(k, 2), v = xv(k, 3), a = xv(k, 4), d = xv(k, 5);
  z[0] = xv(k + 1, 0) - (X + lin[0] * ps + lin[1] * v);
  z[1] = xv(k + 1, 1) - (Y + lin[2] * ps + lin[3] * v);
  z[2] = xv(k + 1, 2) - (ps + lin[4] * d);
  z[3] = xv(k + 1, 3) - (v + p.dt * a);
}
// g value of soft group g of stage k
template <class XV>
MPC_HD double group_g(int k, int g, const XV& xv) {
  if (g == 0) return xv(k, 3);
  if (g <= 2) return xv(k, 4 + g - 1);
  double cur = xv(k, 4 + g - 3);
  return k > 0 ? cur - xv(k - 1, 4 + g - 3) : cur;
}

// (P x)_j for the six (x,u) unknowns of a stage; W = pq or pqn
MPC_HD double px_entry(const Params& p, bool reg, int j, const double* x) {
  if (j < 4) {
    const double (*W)[4] = reg ? p.pq : p.pqn;
    return W[j][0] * x[0] + W[j][1] * x[1] + W[j][2] * x[2] + W[j][3] * x[3];
  }
  return p.pr[j - 4][0] * x[4] + p.pr[j - 4][1] * x[5];
}

struct StateXV { View w; MPC_HD double operator()(int k, int j) const { return w.rec(k)[R_XU + j]; } };
struct BxXV { View w; MPC_HD double operator()(int k, int j) const { return w.bx_get(k, j); } };

// ----------------------------------------------------------------------------------------------
// Problem setup
// ----------------------------------------------------------------------------------------------
// Reference window accessor: row k of the window is row min(start+k, len-1) of `base` ([len][4]), i.e. the
// tail padding of control_stage.py:101-105 / ref_builder.py:19-21; vscale = 0.6 in the relaxation retry
// (control_stage.py:46-47), else 1.
struct RefWin {
  const double* base; int start, len; double vscale;
  MPC_HD const double* row(int k) const { int i = start + k; if (i > len - 1) i = len - 1; return base + 4 * (size_t)i; }
  MPC_HD double v(int k) const { double x = row(k)[3]; return vscale == 1.0 ? x : x * vscale; }
};
// unwrap the window's yaw column (sequential)
MPC_HD void unwrap_window(const RefWin& rw, int count, double* out) {
  const double PI = 3.141592653589793;
  double cum = 0.0, prev = rw.row(0)[2];
  out[0] = prev;
  for (int k = 1; k < count; ++k) {
    double cur = rw.row(k)[2];
    double dd = cur - prev;
    double ddmod = np_mod(dd + PI, 2.0 * PI) - PI;
    if (ddmod == -PI && dd > 0.0) ddmod = PI;
    double corr = ddmod - dd;
    if (fabs(dd) < PI) corr = 0.0;
    cum += corr;
    out[k] = cur + cum;
    prev = cur;
  }
}
// stage k: linearise (at window row max(k-1,0), unwrapped yaw) + linear cost of x_k
MPC_HD void setup_stage(const View& w, const Params& p, int k, const RefWin& rw, const double* uyaw) {
  const int N = w.N;
  double* rc = w.rec(k);
  if (k < N) {
    int kl = k > 0 ? k - 1 : 0;
    const double* r = rw.row(kl);
    linearize_point(p, r[0], r[1], uyaw[kl], rw.v(kl), rc + R_LIN);
  } else {
    for (int j = 0; j < 7; ++j) rc[R_LIN + j] = 0.0;
  }
  const double (*W)[4] = k < N ? p.pq : p.pqn;
  const double* r = rw.row(k);
  const double xr[4] = {r[0], r[1], uyaw[k], rw.v(k)};
  for (int i = 0; i < 4; ++i) rc[R_Q + i] = -(W[i][0] * xr[0] + W[i][1] * xr[1] + W[i][2] * xr[2] + W[i][3] * xr[3]);
}

// cold start: x = 0, y = 0, z = clip(0, l, u)  (v = z for inequality rows)
MPC_HD void cold_start_stage(const View& w, const Params& p, int k) {
  double* rc = w.rec(k);
  for (int j = 0; j < 11; ++j) rc[R_XU + j] = 0.0;
  for (int g = 0; g < 5; ++g) {
    double lo = 0.0, hi = 0.0;
    if (g < ngroups(w.N, k)) group_bounds(p, w.hdr(), k, g, lo, hi);
    rc[R_V + 3 * g + 0] = fmin(0.0, hi);
    rc[R_V + 3 * g + 1] = fmax(0.0, lo);
    rc[R_V + 3 * g + 2] = 0.0;
  }
  for (int r = 0; r < 4; ++r) rc[R_YE + r] = 0.0;
  for (int j = 0; j < 5; ++j) rc[R_ST + j] = 0.0;
  if (k == 0) for (int r = 0; r < 4; ++r) w.hdr()[H_YI + r] = 0.0;
}

// ----------------------------------------------------------------------------------------------
// Band assembly: rows 6k..6k+5 of  M = P + reg I + sum_rows w_i a_i a_i' + sum_groups kappa g g'
//   written in the stage-blocked layout (MPC_IA / MPC_ID / MPC_IC)
// ----------------------------------------------------------------------------------------------
MPC_HD int act_group_bits(const View& w, int k, int g) { return (w.act()[k] >> (3 * g)) & 7; }
MPC_HD int act_dyn_bit(const View& w, int k, int r) { return (w.act()[k] >> (15 + r)) & 1; }
MPC_HD int act_init_bit(const View& w, int r) { return (w.act()[w.N + 1] >> r) & 1; }

// E = M[stage k rows][stage k-1 cols] (k >= 1): dynamics rows of stage k-1 and the rate groups of stage k
MPC_HD void stage_cross(const View& w, const Params& p, const Mode& m, int k, double E[6][6]) {
  for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) E[a][b] = 0.0;
  const double* lin = w.rec(k - 1) + R_LIN;
  double om[4];
  for (int r = 0; r < 4; ++r) om[r] = eq_weight(m, m.polish ? act_dyn_bit(w, k - 1, r) : 1);
  // row r of stage k-1: +1 at x_k[r], -(A,B) part on stage k-1  =>  E[r][:] = om[r] * a_r^{(k-1)}
  E[0][0] = -om[0]; E[0][2] = -om[0] * lin[0]; E[0][3] = -om[0] * lin[1];
  E[1][1] = -om[1]; E[1][2] = -om[1] * lin[2]; E[1][3] = -om[1] * lin[3];
  E[2][2] = -om[2]; E[2][5] = -om[2] * lin[4];
  E[3][3] = -om[3]; E[3][4] = -om[3] * p.dt;
  if (k < w.N) {
    for (int i = 0; i < 2; ++i) {
      GroupCoef cd = group_coef(m, group_ps(p, 3 + i), m.polish ? act_group_bits(w, k, 3 + i) : 0);
      E[4 + i][4 + i] -= cd.kappa;                            // (u_k - u_{k-1}) coupling
    }
  }
}
// D = diagonal block of stage k (lower part filled; dummy unknowns of the terminal stage are identity rows)
MPC_HD void stage_diag(const View& w, const Params& p, const Mode& m, int k, double D[6][6]) {
  const int N = w.N;
  for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) D[a][b] = 0.0;
  const double (*W)[4] = k < N ? p.pq : p.pqn;
  for (int a = 0; a < 4; ++a) {
    for (int b = 0; b < a; ++b) D[a][b] = W[a][b];
    D[a][a] = W[a][a] + m.reg;
  }
  if (k < N) { D[4][4] = p.pr[0][0] + m.reg; D[5][5] = p.pr[1][1] + m.reg; D[5][4] = p.pr[1][0]; }
  else { D[4][4] = 1.0; D[5][5] = 1.0; }
  // rows arriving at x_k: init rows (k == 0) or dynamics rows of stage k-1
  if (k == 0) {
    for (int r = 0; r < 4; ++r) D[r][r] += eq_weight(m, m.polish ? act_init_bit(w, r) : 1);
  } else {
    for (int r = 0; r < 4; ++r) D[r][r] += eq_weight(m, m.polish ? act_dyn_bit(w, k - 1, r) : 1);
  }
  if (k < N) {
    const double* lin = w.rec(k) + R_LIN;
    double om[4];
    for (int r = 0; r < 4; ++r) om[r] = eq_weight(m, m.polish ? act_dyn_bit(w, k, r) : 1);
    // a_0 = (-1,0,-a02,-a03,0,0), a_1 = (0,-1,-a12,-a13,0,0), a_2 = (0,0,-1,0,0,-b21), a_3 = (0,0,0,-1,-dt,0)
    double a0[6] = {-1.0, 0.0, -lin[0], -lin[1], 0.0, 0.0};
    double a1[6] = {0.0, -1.0, -lin[2], -lin[3], 0.0, 0.0};
    double a2[6] = {0.0, 0.0, -1.0, 0.0, 0.0, -lin[4]};
    double a3[6] = {0.0, 0.0, 0.0, -1.0, -p.dt, 0.0};
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b <= a; ++b)
        D[a][b] += om[0] * a0[a] * a0[b] + om[1] * a1[a] * a1[b] + om[2] * a2[a] * a2[b] + om[3] * a3[a] * a3[b];
  }
  // soft groups
  {
    GroupCoef c = group_coef(m, group_ps(p, 0), m.polish ? act_group_bits(w, k, 0) : 0);
    D[3][3] += c.kappa;
  }
  if (k < N) {
    for (int i = 0; i < 2; ++i) {
      GroupCoef cu = group_coef(m, group_ps(p, 1 + i), m.polish ? act_group_bits(w, k, 1 + i) : 0);
      GroupCoef cd = group_coef(m, group_ps(p, 3 + i), m.polish ? act_group_bits(w, k, 3 + i) : 0);
      D[4 + i][4 + i] += cu.kappa + cd.kappa;
      if (k + 1 < N) {
        GroupCoef cn = group_coef(m, group_ps(p, 3 + i), m.polish ? act_group_bits(w, k + 1, 3 + i) : 0);
        D[4 + i][4 + i] += cn.kappa;
      }
    }
  }
}

// Write stage k's blocks in the twisted storage.  Top half (k <= m): in-stage = D_k, cross = E_k (to stage k-1).
// Bottom half (k >= m), local index i = N-k, everything index-reversed: in-stage = rev(D_k) (zero for the border
// k = m, which only collects the bottom half's Schur complement), cross = rev(E_{k+1}') (to stage k+1).
MPC_HD void assemble_stage(const View& w, const Params& p, const Mode& m, int k) {
  const int N = w.N;
  const int ms = mid_stage(N);
  double D[6][6], E[6][6];
  stage_diag(w, p, m, k, D);
  if (k <= ms) {
    double* b = w.top().blk(k);
    if (k > 0) stage_cross(w, p, m, k, E);
    for (int j = 0; j < 6; ++j) {
      for (int jp = 0; jp < j; ++jp) b[MPC_IA(j, jp)] = D[j][jp];
      b[MPC_ID(j)] = D[j][j];
      for (int jp = j; jp < 6; ++jp) b[MPC_IC(j, jp)] = (k > 0) ? E[j][jp] : 0.0;
    }
    b[21] = 0.0; b[43] = 0.0;
  }
  if (k >= ms) {
    double* b = w.bottom().blk(N - k);
    if (k < N) stage_cross(w, p, m, k + 1, E);     // E = M[stage k+1 rows][stage k cols]
    const bool border = (k == ms);
    for (int j = 0; j < 6; ++j) {                  // local (reversed) indices j, jp  <->  global 5-j, 5-jp
      for (int jp = 0; jp < j; ++jp) b[MPC_IA(j, jp)] = border ? 0.0 : D[5 - jp][5 - j];
      b[MPC_ID(j)] = border ? 0.0 : D[5 - j][5 - j];
      // local cross entry (row j of this stage, col jp >= j of the previous local stage = global stage k+1):
      //   M[(k, 5-j)][(k+1, 5-jp)] = E[5-jp][5-j]
      for (int jp = j; jp < 6; ++jp) b[MPC_IC(j, jp)] = (k < N) ? E[5 - jp][5 - j] : 0.0;
    }
    b[21] = 0.0; b[43] = 0.0;
  }
}

// ----------------------------------------------------------------------------------------------
// Banded LDL' (unit lower L, half-bandwidth 6) as a right-looking block algorithm over the local stages of one
// half, everything of a stage staged in registers: factor the 6x6 in-stage block, form the cross block of the next
// stage, apply its Schur complement to the next stage's in-stage block.  Sequential over local stages [i0, i1),
// i1 <= H-1 (the border stage is never eliminated inside a half).
//   after: IA = L in-stage, ID = 1/D, IC = L cross
// ----------------------------------------------------------------------------------------------
MPC_HD void factor_half(const HalfView& h, int i0, int i1) {
  if (i1 > h.H - 1) i1 = h.H - 1;
  for (int k = i0; k < i1; ++k) {
    double* __restrict__ Bk = h.blk(k);
    double S[6][6], Lm[6][6], dinv[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
      for (int jp = 0; jp < j; ++jp) S[j][jp] = Bk[MPC_IA(j, jp)];
      S[j][j] = Bk[MPC_ID(j)];
    }
#pragma unroll
    for (int jp = 0; jp < 6; ++jp) {
      dinv[jp] = 1.0 / S[jp][jp];
#pragma unroll
      for (int j = jp + 1; j < 6; ++j) {
        const double l = S[j][jp] * dinv[jp];
        Lm[j][jp] = l;
#pragma unroll
        for (int j2 = jp + 1; j2 <= j; ++j2) S[j][j2] = fma(-l, S[j2][jp], S[j][j2]);   // S[j2][jp] still unscaled
      }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
      for (int jp = 0; jp < j; ++jp) Bk[MPC_IA(j, jp)] = Lm[j][jp];
      Bk[MPC_ID(j)] = dinv[j];
    }
    {
      double* __restrict__ Bn = h.blk(k + 1);
      double Wc[6][6], Cl[6][6];
#pragma unroll
      for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int jp = j; jp < 6; ++jp) {
          double sv = Bn[MPC_IC(j, jp)];
#pragma unroll
          for (int t = j; t < jp; ++t) sv = fma(-Wc[j][t], Lm[jp][t], sv);
          Wc[j][jp] = sv;
          Cl[j][jp] = sv * dinv[jp];
        }
#pragma unroll
      for (int j = 0; j < 6; ++j) {
#pragma unroll
        for (int jp = j; jp < 6; ++jp) Bn[MPC_IC(j, jp)] = Cl[j][jp];
#pragma unroll
        for (int j2 = 0; j2 <= j; ++j2) {
          double acc = 0.0;
#pragma unroll
          for (int t = j; t < 6; ++t) acc = fma(Wc[j][t], Cl[j2][t], acc);
          if (j2 < j) Bn[MPC_IA(j, j2)] -= acc; else Bn[MPC_ID(j)] -= acc;
        }
      }
    }
  }
}
// Middle stage: S_m = (top border in-stage block) + reversed(bottom border in-stage block); LDL' into mid();
// then clear the borders' in-stage parts so that the backward sweeps start from them with the plain stage step.
MPC_HD void factor_middle(const View& w) {
  HalfView T = w.top(), B = w.bottom();
  double* bt = T.blk(T.H - 1); double* bb = B.blk(B.H - 1); double* md = w.mid();
  double S[6][6];
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c <= r; ++c) {
      double top = r == c ? bt[MPC_ID(r)] : bt[MPC_IA(r, c)];
      const int rr = 5 - c, cc = 5 - r;                                    // reversed position (rr >= cc)
      double bot = rr == cc ? bb[MPC_ID(rr)] : bb[MPC_IA(rr, cc)];
      S[r][c] = top + bot;
    }
  double dinv[6];
  for (int jp = 0; jp < 6; ++jp) {
    dinv[jp] = 1.0 / S[jp][jp];
    for (int j = jp + 1; j < 6; ++j) {
      const double l = S[j][jp] * dinv[jp];
      for (int j2 = jp + 1; j2 <= j; ++j2) S[j][j2] -= l * S[j2][jp];
      md[MPC_IA(j, jp)] = l;
    }
    md[MPC_ID(jp)] = dinv[jp];
  }
  for (int i = 0; i < 21; ++i) { bt[i] = 0.0; bb[i] = 0.0; }
}
// x = S_m^{-1} a  (6x6 LDL' in mid())
MPC_HD void middle_solve(const View& w, const double* a, double* x) {
  const double* md = w.mid();
  double y[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double v = a[j];
#pragma unroll
    for (int jp = 0; jp < j; ++jp) v = fma(-md[MPC_IA(j, jp)], y[jp], v);
    y[j] = v;
  }
#pragma unroll
  for (int j = 5; j >= 0; --j) {
    double v = y[j] * md[MPC_ID(j)];
#pragma unroll
    for (int jp = j + 1; jp < 6; ++jp) v = fma(-md[MPC_IA(jp, j)], x[jp], v);
    x[j] = v;
  }
}
// whole twisted factorisation, sequential (host emulation / single-lane use)
MPC_HD void factor_band(const View& w) {
  HalfView T = w.top(), B = w.bottom();
  factor_half(T, 0, T.H - 1);
  factor_half(B, 0, B.H - 1);
  factor_middle(w);
}

// Triangular sweeps of one half in column-oriented (axpy) form: as soon as a pivot value is known it is pushed
// into the accumulators of the six rows that depend on it, so only ONE fma separates consecutive pivots and the
// other five issue in its shadow (the SM issues in order; a dot-product formulation would serialise six dependent
// fmas per row).  The factor block of a stage is staged in registers with 128-bit loads before any arithmetic.
struct ChainRegs { double la[22], lc[22], nb[6]; };

MPC_HD void chain_load_fwd(const HalfView& h, int k, ChainRegs& r) {
  const D2* __restrict__ pa = reinterpret_cast<const D2*>(h.blk(k));
  const D2* __restrict__ pc = reinterpret_cast<const D2*>(h.blk(k + 1) + 22);
#pragma unroll
  for (int i = 0; i < 11; ++i) { D2 u = pa[i], v = pc[i]; r.la[2 * i] = u.x; r.la[2 * i + 1] = u.y; r.lc[2 * i] = v.x; r.lc[2 * i + 1] = v.y; }
  const D2* __restrict__ bn = reinterpret_cast<const D2*>(h.bx(k + 1));
#pragma unroll
  for (int i = 0; i < 3; ++i) { D2 u = bn[i]; r.nb[2 * i] = u.x; r.nb[2 * i + 1] = u.y; }
}
MPC_HD void chain_load_bwd(const HalfView& h, int k, ChainRegs& r) {
  const D2* __restrict__ pa = reinterpret_cast<const D2*>(h.blk(k));
#pragma unroll
  for (int i = 0; i < 11; ++i) { D2 u = pa[i], v = pa[11 + i]; r.la[2 * i] = u.x; r.la[2 * i + 1] = u.y; r.lc[2 * i] = v.x; r.lc[2 * i + 1] = v.y; }
  const D2* __restrict__ bp = reinterpret_cast<const D2*>(h.bx(k - 1));
#pragma unroll
  for (int i = 0; i < 3; ++i) { D2 u = bp[i]; r.nb[2 * i] = u.x; r.nb[2 * i + 1] = u.y; }
}
MPC_HD void chain_math_fwd(const ChainRegs& r, double* a, double* out) {
  double nx[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) nx[j] = r.nb[j];
#pragma unroll
  for (int jp = 0; jp < 6; ++jp) {
    const double wv = a[jp];
    out[jp] = wv * r.la[MPC_ID(jp)];
#pragma unroll
    for (int j = jp + 1; j < 6; ++j) a[j] = fma(-r.la[MPC_IA(j, jp)], wv, a[j]);
#pragma unroll
    for (int j = 0; j <= jp; ++j) nx[j] = fma(-r.lc[MPC_IC(j, jp) - 22], wv, nx[j]);
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) a[j] = nx[j];
}
MPC_HD void chain_math_bwd(const ChainRegs& r, double* a, double* out) {
  double nx[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) nx[j] = r.nb[j];
#pragma unroll
  for (int jp = 5; jp >= 0; --jp) {
    const double xv = a[jp];
    out[jp] = xv;
#pragma unroll
    for (int j = jp - 1; j >= 0; --j) a[j] = fma(-r.la[MPC_IA(jp, j)], xv, a[j]);
#pragma unroll
    for (int j = 5; j >= jp; --j) nx[j] = fma(-r.lc[MPC_IC(jp, j) - 22], xv, nx[j]);
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) a[j] = nx[j];
}
MPC_HD void chain_store(double* bk, const double* out) {
  D2* q = reinterpret_cast<D2*>(bk);
#pragma unroll
  for (int i = 0; i < 3; ++i) { D2 u; u.x = out[2 * i]; u.y = out[2 * i + 1]; q[i] = u; }
}

// Rolling variant (MODE 2): the in-stage arithmetic (pivot chain) and the cross arithmetic (feeds the next stage) are
// separated and each part's registers are refilled for the next stage as soon as the part is done.  Same operations in
// the same order per accumulator as the interleaved form: bit-identical results.
MPC_HD void chain_load_in(const double* blk, double* la) {
  const D2* __restrict__ pa = reinterpret_cast<const D2*>(blk);
#pragma unroll
  for (int i = 0; i < 11; ++i) { D2 u = pa[i]; la[2 * i] = u.x; la[2 * i + 1] = u.y; }
}
MPC_HD void chain_load_cross(const double* blk, const double* bx, double* lc, double* nb) {
  const D2* __restrict__ pc = reinterpret_cast<const D2*>(blk + 22);
#pragma unroll
  for (int i = 0; i < 11; ++i) { D2 v = pc[i]; lc[2 * i] = v.x; lc[2 * i + 1] = v.y; }
  const D2* __restrict__ pb = reinterpret_cast<const D2*>(bx);
#pragma unroll
  for (int i = 0; i < 3; ++i) { D2 v = pb[i]; nb[2 * i] = v.x; nb[2 * i + 1] = v.y; }
}
MPC_HD void chain_in_fwd(const double* la, double* a, double* out) {
#pragma unroll
  for (int jp = 0; jp < 6; ++jp) {
    const double wv = a[jp];
    out[jp] = wv * la[MPC_ID(jp)];
#pragma unroll
    for (int j = jp + 1; j < 6; ++j) a[j] = fma(-la[MPC_IA(j, jp)], wv, a[j]);
  }
}
MPC_HD void chain_cross_fwd(const double* lc, const double* nb, double* a) {
  double nx[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) nx[j] = nb[j];
#pragma unroll
  for (int jp = 0; jp < 6; ++jp)
#pragma unroll
    for (int j = 0; j <= jp; ++j) nx[j] = fma(-lc[MPC_IC(j, jp) - 22], a[jp], nx[j]);
#pragma unroll
  for (int j = 0; j < 6; ++j) a[j] = nx[j];
}
MPC_HD void chain_in_bwd(const double* la, double* a, double* out) {
#pragma unroll
  for (int jp = 5; jp >= 0; --jp) {
    const double xv = a[jp];
    out[jp] = xv;
#pragma unroll
    for (int j = jp - 1; j >= 0; --j) a[j] = fma(-la[MPC_IA(jp, j)], xv, a[j]);
  }
}
MPC_HD void chain_cross_bwd(const double* lc, const double* nb, double* a) {
  double nx[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) nx[j] = nb[j];
#pragma unroll
  for (int jp = 5; jp >= 0; --jp)
#pragma unroll
    for (int j = 5; j >= jp; --j) nx[j] = fma(-lc[MPC_IC(jp, j) - 22], a[jp], nx[j]);
#pragma unroll
  for (int j = 0; j < 6; ++j) a[j] = nx[j];
}
// one stage of the rolling sweeps; `d` = +1 forward, -1 backward
#define MPC_ROLL_FWD(k) { chain_in_fwd(r0.la, a, out); chain_load_in(h.blk((k) + 1), r0.la); chain_cross_fwd(r0.lc, r0.nb, a); \
                          chain_load_cross(h.blk((k) + 2), h.bx((k) + 2), r0.lc, r0.nb); chain_store(h.bx(k), out); }
#define MPC_ROLL_BWD(k) { chain_in_bwd(r0.la, a, out); chain_load_in(h.blk((k) - 1), r0.la); chain_cross_bwd(r0.lc, r0.nb, a); \
                          chain_load_cross(h.blk((k) - 1), h.bx((k) - 2), r0.lc, r0.nb); chain_store(h.bx(k), out); }

// Forward sweep of one half over its local stages 0..H-2; returns the accumulators of the border stage in a[].
//   PIPE = 0: one register set (~80 registers), the factor block of a stage is loaded at the top of the stage;
//   PIPE = 1: software-pipelined by hand with two register sets, loop unrolled by two (needs > 255 registers: kept for the
//             microbenchmark only);
//   PIPE = 2: rolling refill (above), loop unrolled by two so that half of the refills have their consumer in the same
//             loop body and ptxas interleaves them with the arithmetic (~165 registers).  Used by the ADMM loop.
template <int PIPE>
MPC_HD void half_forward(const HalfView& h, double* a) {
  const int last = h.H - 2;                 // last eliminated local stage
  double out[6];
  ChainRegs r0;
  {
    const double* b0 = h.bx(0);
#pragma unroll
    for (int j = 0; j < 6; ++j) a[j] = b0[j];
  }
  if (PIPE == 2) {                          // (refills past the end read the border / padding block and row H: unused)
    if (last < 0) return;
    chain_load_in(h.blk(0), r0.la);
    chain_load_cross(h.blk(1), h.bx(1), r0.lc, r0.nb);
    int k = 0;
    for (; k + 1 <= last; k += 2) { MPC_ROLL_FWD(k); MPC_ROLL_FWD(k + 1); }
    if (k <= last) MPC_ROLL_FWD(k);
  } else if (PIPE == 1) {
    ChainRegs r1;
    if (last >= 0) chain_load_fwd(h, 0, r0);
    int k = 0;
    for (; k + 1 <= last; k += 2) {
      chain_load_fwd(h, k + 1, r1);
      chain_math_fwd(r0, a, out);
      chain_store(h.bx(k), out);
      if (k + 2 <= last) chain_load_fwd(h, k + 2, r0);
      chain_math_fwd(r1, a, out);
      chain_store(h.bx(k + 1), out);
    }
    if (k <= last) {
      chain_math_fwd(r0, a, out);
      chain_store(h.bx(k), out);
    }
  } else {
    for (int k = 0; k <= last; ++k) {
      chain_load_fwd(h, k, r0);
      chain_math_fwd(r0, a, out);
      chain_store(h.bx(k), out);
    }
  }
}
// Backward sweep of one half from the border (whose solution xm, in the half's local index order, is given and
// whose in-stage factor part is zero) down to local stage 0.
template <int PIPE>
MPC_HD void half_backward(const HalfView& h, const double* xm) {
  double a[6], out[6];
  ChainRegs r0;
#pragma unroll
  for (int j = 0; j < 6; ++j) a[j] = xm[j];
  const int first = h.H - 1;
  if (PIPE == 2) {                          // (refills below stage 0 read in-bounds words of this problem's workspace: unused)
    chain_load_in(h.blk(first), r0.la);
    chain_load_cross(h.blk(first), h.bx(first - 1), r0.lc, r0.nb);
    int k = first;
    for (; k - 1 >= 0; k -= 2) { MPC_ROLL_BWD(k); MPC_ROLL_BWD(k - 1); }
    if (k >= 0) MPC_ROLL_BWD(k);
  } else if (PIPE == 1) {
    ChainRegs r1;
    chain_load_bwd(h, first, r0);
    int k = first;
    for (; k - 1 >= 0; k -= 2) {
      chain_load_bwd(h, k - 1, r1);
      chain_math_bwd(r0, a, out);
      chain_store(h.bx(k), out);
      if (k - 2 >= 0) chain_load_bwd(h, k - 2, r0);
      chain_math_bwd(r1, a, out);
      chain_store(h.bx(k - 1), out);
    }
    if (k >= 0) {
      chain_math_bwd(r0, a, out);
      chain_store(h.bx(k), out);
    }
  } else {
    for (int k = first; k >= 0; --k) {
      chain_load_bwd(h, k, r0);
      chain_math_bwd(r0, a, out);
      chain_store(h.bx(k), out);
    }
  }
}
// whole twisted solve, sequential (host emulation / single-lane use): L D L' x = b in place on the bx storage
template <int MODE>
MPC_HD void chain_solve_mode(const View& w) {
  HalfView T = w.top(), B = w.bottom();
  double aT[6], aB[6], xm[6], xr[6];
  half_forward<MODE>(T, aT);
  half_forward<MODE>(B, aB);
  for (int j = 0; j < 6; ++j) aT[j] += aB[5 - j];
  middle_solve(w, aT, xm);
  for (int j = 0; j < 6; ++j) xr[j] = xm[5 - j];
  half_backward<MODE>(T, xm);
  half_backward<MODE>(B, xr);
}
MPC_HD void chain_solve(const View& w) { chain_solve_mode<0>(w); }
MPC_HD void chain_solve_rolling(const View& w) { chain_solve_mode<2>(w); }   // the kernel's hot-loop order of loads

// ----------------------------------------------------------------------------------------------
// ADMM step, split in two stage-parallel halves
//   A1(k): consume x-tilde (bx) and s-tilde (R_ST): relax x, update row states (v, ye, yi)
//   A2(k): from the new state form t = rho z - y per row, s-tilde, and the banded rhs -> bx
// ----------------------------------------------------------------------------------------------
MPC_HD void admm_update_stage(const View& w, const Params& p, const Settings& s, double rho, int k) {
  const int N = w.N;
  double* rc = w.rec(k);
  const double alpha = s.alpha, oma = 1.0 - s.alpha;
  const double rho_eq = s.rho_eq_factor * rho;
  BxXV xt{w};
  // soft groups
  const int ng = ngroups(N, k);
  for (int g = 0; g < ng; ++g) {
    double lo, hi; group_bounds(p, w.hdr(), k, g, lo, hi);
    double gt = group_g(k, g, xt);
    double st = rc[R_ST + g];
    double zt[3] = {gt - st, gt + st, st};
    double blo[3] = {-1e30, lo, 0.0}, bhi[3] = {hi, 1e30, 1e30};
    for (int r = 0; r < 3; ++r) {
      double v = rc[R_V + 3 * g + r];
      double z = clipd(v, blo[r], bhi[r]);
      double wv = alpha * zt[r] + oma * z;
      rc[R_V + 3 * g + r] = wv + (v - z);
    }
    rc[R_S + g] = alpha * st + oma * rc[R_S + g];
  }
  // equality rows
  if (k < N) {
    double zt[4]; dyn_rows(w, p, k, xt, zt);
    const double* lin = rc + R_LIN;
    double b[4] = {lin[5], lin[6], 0.0, 0.0};
    for (int r = 0; r < 4; ++r) rc[R_YE + r] += rho_eq * alpha * (zt[r] - b[r]);
  }
  if (k == 0) {
    double* h = w.hdr();
    for (int r = 0; r < 4; ++r) h[H_YI + r] += rho_eq * alpha * (xt(0, r) - h[H_X0 + r]);
  }
}
// second half of A1: relax x (kept separate because A1 reads neighbours' x-tilde from bx, not state)
MPC_HD void admm_relax_x_stage(const View& w, const Settings& s, int k) {
  double* rc = w.rec(k);
  const int nj = k < w.N ? 6 : 4;
  for (int j = 0; j < nj; ++j) rc[R_XU + j] = s.alpha * w.bx_get(k, j) + (1.0 - s.alpha) * rc[R_XU + j];
}

// row provider for the ADMM rhs: value = rho_i z_i - y_i
struct AdmmRP {
  View w; const Params* p; double rho, rho_eq;
  MPC_HD double dyn(int k, int r) const {
    const double* rc = w.rec(k);
    double b = r == 0 ? rc[R_LIN + 5] : (r == 1 ? rc[R_LIN + 6] : 0.0);
    return rho_eq * b - rc[R_YE + r];
  }
  MPC_HD double init(int r) const { return rho_eq * w.hdr()[H_X0 + r] - w.hdr()[H_YI + r]; }
  MPC_HD double t_row(int k, int g, int r) const {
    double lo, hi; group_bounds(*p, w.hdr(), k, g, lo, hi);
    double v = w.rec(k)[R_V + 3 * g + r];
    double z = r == 0 ? fmin(v, hi) : (r == 1 ? fmax(v, lo) : fmax(v, 0.0));
    return rho * (2.0 * z - v);
  }
  MPC_HD double grp(int k, int g) const { return t_row(k, g, 0) + t_row(k, g, 1); }
};

MPC_HD void admm_rhs_stage(const View& w, const Params& p, const Settings& s, double rho, int k) {
  const int N = w.N;
  double* rc = w.rec(k);
  AdmmRP rp{w, &p, rho, s.rho_eq_factor * rho};
  const int ng = ngroups(N, k);
  for (int g = 0; g < ng; ++g) {
    double t1 = rp.t_row(k, g, 0), t2 = rp.t_row(k, g, 1), t3 = rp.t_row(k, g, 2);
    double mss_inv = 1.0 / (group_ps(p, g) + s.sigma + 3.0 * rho);
    rc[R_ST + g] = (s.sigma * rc[R_S + g] + (-t1 + t2 + t3)) * mss_inv;
  }
  double out[6];
  gather_xu(w, p, k, rp, out);
  const int nj = k < N ? 6 : 4;
  for (int j = 0; j < nj; ++j) {
    double qj = j < 4 ? rc[R_Q + j] : 0.0;
    w.bx_set(k, j, s.sigma * rc[R_XU + j] - qj + out[j]);
  }
  for (int j = nj; j < 6; ++j) w.bx_set(k, j, 0.0);
}

// ----------------------------------------------------------------------------------------------
// Specialised ADMM phases (the per-iteration hot code).  Same mathematics as admm_update_stage /
// admm_relax_x_stage / admm_rhs_stage above (kept as the readable statement and used by nothing hot), with the
// loops over groups and rows unrolled at compile time, bounds hoisted, one-sided clips instead of two-sided,
// and v + alpha (z~ - z) in place of alpha z~ + (1-alpha) z + (v - z).
// ----------------------------------------------------------------------------------------------
}  // namespace mpc
#include "mpc_oe.h"
namespace mpc {

struct IterConst {
  double rho, rho_eq, alpha, sigma, ra;     // ra = rho_eq * alpha
  double lo[5], hi[5], mssinv[5];           // soft-group bounds (stage k > 0) and 1 / (2w + sigma + 3 rho)
  double ps_[5];                            // P entries of the slacks (2w)
  double up0, up1;                          // u_prev: shifts the rate bounds of stage 0
  double kap;                               // 2 rho: coupling of consecutive inputs through the rate groups
};
MPC_HD IterConst iter_const(const View& w, const Params& p, const Settings& s, double rho) {
  IterConst c;
  c.rho = rho; c.rho_eq = s.rho_eq_factor * rho; c.alpha = s.alpha; c.sigma = s.sigma; c.ra = c.rho_eq * s.alpha;
  c.kap = 2.0 * rho;
  c.lo[0] = p.v_lo; c.hi[0] = p.v_hi;
  c.lo[1] = p.u_lo[0]; c.hi[1] = p.u_hi[0]; c.lo[2] = p.u_lo[1]; c.hi[2] = p.u_hi[1];
  c.lo[3] = p.du_lo[0]; c.hi[3] = p.du_hi[0]; c.lo[4] = p.du_lo[1]; c.hi[4] = p.du_hi[1];
  c.mssinv[0] = 1.0 / (2.0 * p.w_v + s.sigma + 3.0 * rho);
  c.mssinv[1] = c.mssinv[2] = 1.0 / (2.0 * p.w_u + s.sigma + 3.0 * rho);
  c.mssinv[3] = c.mssinv[4] = 1.0 / (2.0 * p.w_du + s.sigma + 3.0 * rho);
  for (int g = 0; g < 5; ++g) c.ps_[g] = group_ps(p, g);
  c.up0 = w.hdr()[H_UPREV]; c.up1 = w.hdr()[H_UPREV + 1];
  return c;
}
MPC_HD double dmin2(double a, double b) { return a < b ? a : b; }
MPC_HD double dmax2(double a, double b) { return a > b ? a : b; }
// pointer to the six entries of stage k in the twisted rhs/solution storage; rev: stored index-reversed
MPC_HD double* bx_ptr(const View& w, int k, int& rev) {
  const int m = mid_stage(w.N);
  if (k <= m) { rev = 0; return w.top().bx(k); }
  rev = 1; return w.bottom().bx(w.N - k);
}
// branch-free: entry j sits at p[j * st] with (p, st) = (row start, +1) in the top half, (row end, -1) in the bottom half,
// so lanes on both sides of the middle stage stay converged
MPC_HD void bx_load6(const View& w, int k, double* x) {
  const int m = mid_stage(w.N);
  const bool rev = k > m;
  const double* top = w.top().bx(k);
  const double* bot = w.bottom().bx(w.N - k) + 5;
  const double* p = rev ? bot : top;
  const int st = rev ? -1 : 1;
#pragma unroll
  for (int j = 0; j < 6; ++j) x[j] = p[j * st];
}

// A1: consume x-tilde / s-tilde, relax x and s, update the row states.  xt = x-tilde of stage k, xn = of stage k+1 (read
// when k < N), (ua, ud) = the inputs of stage k-1 (0 for k = 0).
MPC_HD void admm_update_vals(const View& w, const Params& p, const IterConst& c, int k, const double* xt, const double* xn, double ua, double ud) {
  // Branch-free over the groups and the terminal stage (see mpc_pair.h: pair_stage): the unused slots of the terminal record hold
  // exact zeros, a fixed point of the update, so the one terminal lane computes them like a regular stage and the compiler gets the
  // five groups, the dynamics rows and the relaxation as ONE basic block.  Only the dynamics duals and the inputs need a select.
  const int N = w.N;
  double* rc = w.rec(k);
  const bool reg = k < N;
  const double off0 = k == 0 ? c.up0 : 0.0, off1 = k == 0 ? c.up1 : 0.0;
  const double gt[5] = {xt[3], reg ? xt[4] : 0.0, reg ? xt[5] : 0.0, reg ? xt[4] - ua : 0.0, reg ? xt[5] - ud : 0.0};
  const double offs[5] = {0.0, 0.0, 0.0, off0, off1};
#pragma unroll
  for (int g = 0; g < 5; ++g) {
    const double st = rc[R_ST + g], sv = rc[R_S + g];
    const double v0 = rc[R_V + 3 * g], v1 = rc[R_V + 3 * g + 1], v2 = rc[R_V + 3 * g + 2];
    const double z0 = dmin2(v0, c.hi[g] + offs[g]), z1 = dmax2(v1, c.lo[g] + offs[g]), z2 = dmax2(v2, 0.0);
    rc[R_V + 3 * g] = fma(c.alpha, (gt[g] - st) - z0, v0);
    rc[R_V + 3 * g + 1] = fma(c.alpha, (gt[g] + st) - z1, v1);
    rc[R_V + 3 * g + 2] = fma(c.alpha, st - z2, v2);
    rc[R_S + g] = fma(c.alpha, st - sv, sv);
  }
  {
    const double* lin = rc + R_LIN;
    const double z0 = xn[0] - (xt[0] + lin[0] * xt[2] + lin[1] * xt[3]);
    const double z1 = xn[1] - (xt[1] + lin[2] * xt[2] + lin[3] * xt[3]);
    const double z2 = xn[2] - (xt[2] + lin[4] * xt[5]);
    const double z3 = xn[3] - (xt[3] + p.dt * xt[4]);
    const double y0 = fma(c.ra, z0 - lin[5], rc[R_YE + 0]), y1 = fma(c.ra, z1 - lin[6], rc[R_YE + 1]);
    const double y2 = fma(c.ra, z2, rc[R_YE + 2]), y3 = fma(c.ra, z3, rc[R_YE + 3]);
    rc[R_YE + 0] = reg ? y0 : 0.0; rc[R_YE + 1] = reg ? y1 : 0.0; rc[R_YE + 2] = reg ? y2 : 0.0; rc[R_YE + 3] = reg ? y3 : 0.0;
  }
  if (k == 0) {
    double* h = w.hdr();
#pragma unroll
    for (int r = 0; r < 4; ++r) h[H_YI + r] = fma(c.ra, xt[r] - h[H_X0 + r], h[H_YI + r]);
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const double xo = rc[R_XU + j];
    const double xw = fma(c.alpha, xt[j] - xo, xo);
    rc[R_XU + j] = (j < 4 || reg) ? xw : 0.0;
  }
}

// A2: t = rho z - y of every row from the new state, s-tilde, right-hand side of stage k -> val[6]
MPC_HD void admm_rhs_vals(const View& w, const Params& p, const IterConst& c, int k, double* val) {
  // branch-free like A1: for the terminal stage G[1..4], its dynamics duals and lin are exact zeros, so its terms vanish; the
  // neighbours' records are read through clamped indices and selected
  const int N = w.N;
  double* rc = w.rec(k);
  const bool reg = k < N;
  const double off0 = k == 0 ? c.up0 : 0.0, off1 = k == 0 ? c.up1 : 0.0;
  const double offs[5] = {0.0, 0.0, 0.0, off0, off1};
  double G[5];
#pragma unroll
  for (int g = 0; g < 5; ++g) {
    const double v0 = rc[R_V + 3 * g], v1 = rc[R_V + 3 * g + 1], v2 = rc[R_V + 3 * g + 2];
    const double z0 = dmin2(v0, c.hi[g] + offs[g]), z1 = dmax2(v1, c.lo[g] + offs[g]), z2 = dmax2(v2, 0.0);
    const double t0 = c.rho * (z0 + (z0 - v0)), t1 = c.rho * (z1 + (z1 - v1)), t2 = c.rho * (z2 + (z2 - v2));
    rc[R_ST + g] = (c.sigma * rc[R_S + g] + ((t1 - t0) + t2)) * c.mssinv[g];
    G[g] = t0 + t1;
  }
  double out[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  {
    const double* rp = w.rec(k >= 1 ? k - 1 : 0);
    const double* h = w.hdr();
    const double p0 = c.rho_eq * rp[R_LIN + 5] - rp[R_YE + 0], p1 = c.rho_eq * rp[R_LIN + 6] - rp[R_YE + 1];
    const double p2 = -rp[R_YE + 2], p3 = -rp[R_YE + 3];
    const double i0_ = c.rho_eq * h[H_X0 + 0] - h[H_YI + 0], i1_ = c.rho_eq * h[H_X0 + 1] - h[H_YI + 1];
    const double i2_ = c.rho_eq * h[H_X0 + 2] - h[H_YI + 2], i3_ = c.rho_eq * h[H_X0 + 3] - h[H_YI + 3];
    out[0] = k >= 1 ? p0 : i0_; out[1] = k >= 1 ? p1 : i1_; out[2] = k >= 1 ? p2 : i2_; out[3] = k >= 1 ? p3 : i3_;
  }
  out[3] += G[0];
  {
    const double* lin = rc + R_LIN;
    const double d0 = reg ? c.rho_eq * lin[5] - rc[R_YE + 0] : 0.0, d1 = reg ? c.rho_eq * lin[6] - rc[R_YE + 1] : 0.0;
    const double d2 = reg ? -rc[R_YE + 2] : 0.0, d3 = reg ? -rc[R_YE + 3] : 0.0;
    out[0] -= d0;
    out[1] -= d1;
    out[2] -= lin[0] * d0 + lin[2] * d1 + d2;
    out[3] -= lin[1] * d0 + lin[3] * d1 + d3;
    out[4] = G[1] + G[3] - p.dt * d3;
    out[5] = G[2] + G[4] - lin[4] * d2;
    const double* rn = w.rec(k + 1 <= N ? k + 1 : N);                  // rate groups of stage k+1 (zeros at the terminal stage)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double v0 = rn[R_V + 3 * (3 + i)], v1 = rn[R_V + 3 * (3 + i) + 1];
      const double z0 = dmin2(v0, c.hi[3 + i]), z1 = dmax2(v1, c.lo[3 + i]);
      const double rsum = c.rho * (z0 + (z0 - v0)) + c.rho * (z1 + (z1 - v1));
      out[4 + i] -= (k + 1 < N) ? rsum : 0.0;
    }
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) val[j] = (j < 4 || reg) ? c.sigma * rc[R_XU + j] - (j < 4 ? rc[R_Q + j] : 0.0) + out[j] : 0.0;
}

// The two phases as the iteration loop runs them, ONE PARITY OF STAGES AT A TIME, with the stage-parallel parts of the
// odd-even block solve (mpc_oe.h) fused in:
//   rhs,    odd  k: b_k -> t_k = D_k^-1 b_k                                   (rows of the even neighbours are not touched)
//   rhs,    even k: b_k -> b'_k = b_k - E_k t_{k-1} - E_{k+1}' t_{k+1}        (after the odd pass)
//   update, odd  k: x~_k = t_k - D_k^-1 (E_k x~_{k-1} + E_{k+1}' x~_{k+1}), then A1(k)   (after the sweeps over the even stages)
//   update, even k: A1(k)                                                     (after the odd pass: it reads x~_{k-1}, x~_{k+1})
MPC_HD void admm_rhs_stage_oe(const View& w, const Params& p, const IterConst& c, const OEView& oe, int k) {
  const int N = w.N;
  double val[6];
  admm_rhs_vals(w, p, c, k, val);
  if (k & 1) {
    double di[OE_SYM], t[6];
    sym_load(oe.dinv + OE_SYM * (k >> 1), di);
    symv6(di, val, t);
    row_store(w.nx(k), t);
  } else {
    double t[6], y[6];
    if (k >= 1) {
      row_load(w.nx(k - 1), t);
      cross_mul(w.rec(k - 1) + R_LIN, p.dt, c.rho_eq, c.kap, k < N, t, y);
#pragma unroll
      for (int j = 0; j < 6; ++j) val[j] -= y[j];
    }
    if (k + 1 <= N) {
      row_load(w.nx(k + 1), t);
      cross_mul_t(w.rec(k) + R_LIN, p.dt, c.rho_eq, c.kap, k + 1 < N, t, y);
#pragma unroll
      for (int j = 0; j < 6; ++j) val[j] -= y[j];
    }
    row_store(w.nx(k), val);
  }
}
MPC_HD void admm_update_stage_oe(const View& w, const Params& p, const IterConst& c, const OEView& oe, int k) {
  const int N = w.N;
  double xt[6], xn[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, xp[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (k >= 1) row_load(w.nx(k - 1), xp);
  if (k + 1 <= N) row_load(w.nx(k + 1), xn);
  if (k & 1) {
    double di[OE_SYM], t[6], v[6], y[6], u[6];
    row_load(w.nx(k), t);
    cross_mul(w.rec(k - 1) + R_LIN, p.dt, c.rho_eq, c.kap, k < N, xp, v);
    if (k + 1 <= N) {
      cross_mul_t(w.rec(k) + R_LIN, p.dt, c.rho_eq, c.kap, k + 1 < N, xn, y);
#pragma unroll
      for (int j = 0; j < 6; ++j) v[j] += y[j];
    }
    sym_load(oe.dinv + OE_SYM * (k >> 1), di);
    symv6(di, v, u);
#pragma unroll
    for (int j = 0; j < 6; ++j) xt[j] = t[j] - u[j];
    row_store(w.nx(k), xt);
  } else {
    row_load(w.nx(k), xt);
  }
  const bool inputs_before = k > 0 && k < N;
  admm_update_vals(w, p, c, k, xt, xn, inputs_before ? xp[4] : 0.0, inputs_before ? xp[5] : 0.0);
}

// Short horizons (N+1 <= 32: every stage has its own lane of one warp).  The heavy parts of the two phases do not depend
// on the other parity - A2 reads the neighbours' STATE, A1 reads x-tilde of all stages once it exists - so they run as ONE
// pass over all stages; only the two light block-solve steps remain parity passes:
//   rhs   : all k: b_k (odd k: stored as t_k = D_k^-1 b_k)   ->   even k: b'_k = b_k - E_k t_{k-1} - E_{k+1}' t_{k+1}
//   update: odd k: x~_k = t_k - D_k^-1 (E_k x~_{k-1} + E_{k+1}' x~_{k+1})   ->   all k: A1
MPC_HD void admm_rhs_stage_short(const View& w, const Params& p, const IterConst& c, const OEView& oe, int k) {
  double val[6];
  admm_rhs_vals(w, p, c, k, val);
  if (k & 1) {
    double di[OE_SYM], t[6];
    sym_load(oe.dinv + OE_SYM * (k >> 1), di);
    symv6(di, val, t);
    row_store(w.nx(k), t);
  } else {
    row_store(w.nx(k), val);
  }
}
MPC_HD void oe_even_fixup(const View& w, const Params& p, const IterConst& c, int k) {
  const int N = w.N;
  double val[6], t[6], y[6];
  row_load(w.nx(k), val);
  if (k >= 1) {
    row_load(w.nx(k - 1), t);
    cross_mul(w.rec(k - 1) + R_LIN, p.dt, c.rho_eq, c.kap, k < N, t, y);
#pragma unroll
    for (int j = 0; j < 6; ++j) val[j] -= y[j];
  }
  if (k + 1 <= N) {
    row_load(w.nx(k + 1), t);
    cross_mul_t(w.rec(k) + R_LIN, p.dt, c.rho_eq, c.kap, k + 1 < N, t, y);
#pragma unroll
    for (int j = 0; j < 6; ++j) val[j] -= y[j];
  }
  row_store(w.nx(k), val);
}
MPC_HD void oe_expand_odd(const View& w, const Params& p, const IterConst& c, const OEView& oe, int k) {
  const int N = w.N;
  double xp[6], xn[6], di[OE_SYM], t[6], v[6], y[6], u[6];
  row_load(w.nx(k - 1), xp);
  row_load(w.nx(k), t);
  cross_mul(w.rec(k - 1) + R_LIN, p.dt, c.rho_eq, c.kap, k < N, xp, v);
  if (k + 1 <= N) {
    row_load(w.nx(k + 1), xn);
    cross_mul_t(w.rec(k) + R_LIN, p.dt, c.rho_eq, c.kap, k + 1 < N, xn, y);
#pragma unroll
    for (int j = 0; j < 6; ++j) v[j] += y[j];
  }
  sym_load(oe.dinv + OE_SYM * (k >> 1), di);
  symv6(di, v, u);
#pragma unroll
  for (int j = 0; j < 6; ++j) t[j] -= u[j];
  row_store(w.nx(k), t);
}
MPC_HD void admm_update_stage_short(const View& w, const Params& p, const IterConst& c, int k) {
  const int N = w.N;
  double xt[6], xn[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, xp[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  row_load(w.nx(k), xt);
  if (k >= 1) row_load(w.nx(k - 1), xp);
  if (k + 1 <= N) row_load(w.nx(k + 1), xn);
  const bool inputs_before = k > 0 && k < N;
  admm_update_vals(w, p, c, k, xt, xn, inputs_before ? xp[4] : 0.0, inputs_before ? xp[5] : 0.0);
}

}  // namespace mpc
#include "mpc_pair.h"
#include "mpc_reg.h"
namespace mpc {

// ----------------------------------------------------------------------------------------------
// Residuals of the current iterate (x, z = clip(v), y): per-stage partial maxima
//   r[0] |Ax - z|, r[1] |Ax|, r[2] |z|, r[3] |Px + q + A'y|, r[4] |Px|, r[5] |A'y|, r[6] |q|
// mode 0: ADMM state (R_V holds v);  mode 1: polished state (R_V holds y, z := clip(Ax))
// ----------------------------------------------------------------------------------------------
struct DualRP {       // value = y_i
  View w; const Params* p; double rho; int ymode;
  MPC_HD double dyn(int k, int r) const { return w.rec(k)[R_YE + r]; }
  MPC_HD double init(int r) const { return w.hdr()[H_YI + r]; }
  MPC_HD double y_row(int k, int g, int r) const {
    double v = w.rec(k)[R_V + 3 * g + r];
    if (ymode) return v;
    double lo, hi; group_bounds(*p, w.hdr(), k, g, lo, hi);
    double z = r == 0 ? fmin(v, hi) : (r == 1 ? fmax(v, lo) : fmax(v, 0.0));
    return rho * (v - z);
  }
  MPC_HD double grp(int k, int g) const { return y_row(k, g, 0) + y_row(k, g, 1); }
};

MPC_HD void residual_stage(const View& w, const Params& p, double rho, int ymode, int k, double* r) {
  const int N = w.N;
  const double* rc = w.rec(k);
  StateXV xs{w};
  DualRP rp{w, &p, rho, ymode};
  const int ng = ngroups(N, k);
  // primal
  for (int g = 0; g < ng; ++g) {
    double lo, hi; group_bounds(p, w.hdr(), k, g, lo, hi);
    double gv = group_g(k, g, xs), sv = rc[R_S + g];
    double ax[3] = {gv - sv, gv + sv, sv};
    double blo[3] = {-1e30, lo, 0.0}, bhi[3] = {hi, 1e30, 1e30};
    for (int q = 0; q < 3; ++q) {
      double z = ymode ? clipd(ax[q], blo[q], bhi[q]) : clipd(rc[R_V + 3 * g + q], blo[q], bhi[q]);
      r[0] = dmax(r[0], fabs(ax[q] - z)); r[1] = dmax(r[1], fabs(ax[q])); r[2] = dmax(r[2], fabs(z));
    }
  }
  if (k < N) {
    double ax[4]; dyn_rows(w, p, k, xs, ax);
    double b[4] = {rc[R_LIN + 5], rc[R_LIN + 6], 0.0, 0.0};
    for (int q = 0; q < 4; ++q) {
      r[0] = dmax(r[0], fabs(ax[q] - b[q])); r[1] = dmax(r[1], fabs(ax[q])); r[2] = dmax(r[2], fabs(b[q]));
    }
  }
  if (k == 0) {
    const double* h = w.hdr();
    for (int q = 0; q < 4; ++q) {
      double ax = xs(0, q), b = h[H_X0 + q];
      r[0] = dmax(r[0], fabs(ax - b)); r[1] = dmax(r[1], fabs(ax)); r[2] = dmax(r[2], fabs(b));
    }
  }
  // dual
  double aty[6]; gather_xu(w, p, k, rp, aty);
  const int nj = k < N ? 6 : 4;
  for (int j = 0; j < nj; ++j) {
    double px = px_entry(p, k < N, j, rc + R_XU);
    double qj = j < 4 ? rc[R_Q + j] : 0.0;
    r[3] = dmax(r[3], fabs(px + qj + aty[j])); r[4] = dmax(r[4], fabs(px)); r[5] = dmax(r[5], fabs(aty[j]));
    r[6] = dmax(r[6], fabs(qj));
  }
  for (int g = 0; g < ng; ++g) {
    double px = group_ps(p, g) * rc[R_S + g];
    double ay = -rp.y_row(k, g, 0) + rp.y_row(k, g, 1) + rp.y_row(k, g, 2);
    r[3] = dmax(r[3], fabs(px + ay)); r[4] = dmax(r[4], fabs(px)); r[5] = dmax(r[5], fabs(ay));
  }
}

// rho change: keep (z, y) fixed, re-express v = z + y / rho_new
MPC_HD void rescale_v_stage(const View& w, const Params& p, double rho_old, double rho_new, int k) {
  double* rc = w.rec(k);
  const int ng = ngroups(w.N, k);
  const double f = rho_old / rho_new;
  for (int g = 0; g < ng; ++g) {
    double lo, hi; group_bounds(p, w.hdr(), k, g, lo, hi);
    double blo[3] = {-1e30, lo, 0.0}, bhi[3] = {hi, 1e30, 1e30};
    for (int r = 0; r < 3; ++r) {
      double v = rc[R_V + 3 * g + r];
      double z = clipd(v, blo[r], bhi[r]);
      rc[R_V + 3 * g + r] = z + (v - z) * f;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Polish
// ----------------------------------------------------------------------------------------------
// pass 1 (from_admm = 1): activity from the ADMM pair (z, y): upper-active  u - z < y  <=> v > hi, etc.;
//                         then R_V is overwritten by y (inactive rows: exactly 0).
// later passes:           activity from  Ax + y  (primal-dual active set), y kept only on active rows.
// returns 1 if the activity mask of this stage changed.
MPC_HD int polish_activity_stage(const View& w, const Params& p, double rho, int from_admm, int k) {
  const int N = w.N;
  double* rc = w.rec(k);
  StateXV xs{w};
  int bits = 0;
  const int ng = ngroups(N, k);
  for (int g = 0; g < ng; ++g) {
    double lo, hi; group_bounds(p, w.hdr(), k, g, lo, hi);
    double blo[3] = {-1e30, lo, 0.0}, bhi[3] = {hi, 1e30, 1e30};
    double gv = group_g(k, g, xs), sv = rc[R_S + g];
    double ax[3] = {gv - sv, gv + sv, sv};
    for (int r = 0; r < 3; ++r) {
      double v = rc[R_V + 3 * g + r];
      double y;
      int a;
      if (from_admm) {
        double z = clipd(v, blo[r], bhi[r]);
        y = rho * (v - z);
        a = (r == 0) ? (bhi[r] - z < y) : (z - blo[r] < -y);
      } else {
        // primal-dual active set on the polished pair; rows that are active with a zero multiplier
        // (|margin| below round-off of the regularised solve) keep their previous bit
        y = v;
        double margin = (r == 0) ? (y - (bhi[r] - ax[r])) : (-y - (ax[r] - blo[r]));
        a = margin > 0.0;
        if (fabs(margin) <= 1e-7) a = (w.act()[k] >> (3 * g + r)) & 1;
      }
      rc[R_V + 3 * g + r] = a ? y : 0.0;
      bits |= a << (3 * g + r);
    }
  }
  if (k < N) for (int r = 0; r < 4; ++r) { int a = rc[R_YE + r] != 0.0; bits |= a << (15 + r); }
  int changed = (w.act()[k] != bits) || from_admm;
  w.act()[k] = bits;
  if (k == 0) {
    int ib = 0;
    for (int r = 0; r < 4; ++r) ib |= (w.hdr()[H_YI + r] != 0.0) << r;
    changed |= (w.act()[N + 1] != ib);
    w.act()[N + 1] = ib;
  }
  return changed;
}

// Active set the ADMM pair (z, y) suggests right now (upper-active u - z < y <=> v > hi, ...), WITHOUT touching the
// state: stores it in act[] and reports whether it differs from the stored one.
MPC_HD int activity_probe_stage(const View& w, const Params& p, int k) {
  const int N = w.N;
  const double* rc = w.rec(k);
  const int old = w.act()[k];
  int bits = 0;
  const int ng = ngroups(N, k);
  const double tol = 1e-6;       // rows closer than this to their bound (weakly active, round-off decides) keep their bit
  for (int g = 0; g < ng; ++g) {
    double lo, hi; group_bounds(p, w.hdr(), k, g, lo, hi);
    const double m0 = rc[R_V + 3 * g] - hi, m1 = lo - rc[R_V + 3 * g + 1], m2 = -rc[R_V + 3 * g + 2];
    const int b0 = fabs(m0) <= tol ? (old >> (3 * g)) & 1 : (m0 > 0.0);
    const int b1 = fabs(m1) <= tol ? (old >> (3 * g + 1)) & 1 : (m1 > 0.0);
    const int b2 = fabs(m2) <= tol ? (old >> (3 * g + 2)) & 1 : (m2 > 0.0);
    bits |= (b0 << (3 * g)) | (b1 << (3 * g + 1)) | (b2 << (3 * g + 2));
  }
  if (k < N) bits |= 0xF << 15;      // equality rows are always active
  int changed = (old != bits);
  w.act()[k] = bits;
  if (k == 0) {
    changed |= (w.act()[N + 1] != 0xF);
    w.act()[N + 1] = 0xF;
  }
  return changed;
}

// row provider for the polish step: value_i = act_i * (e2_i / delta - y_i),  e2_i = b_i - (A x)_i
struct PolishRP {
  View w; const Params* p; double inv_delta; int zero_sol;   // zero_sol: first step, sol = 0 (state not yet reset)
  MPC_HD double dyn(int k, int r) const {
    if (!act_dyn_bit(w, k, r)) return 0.0;
    const double* rc = w.rec(k);
    double b = r == 0 ? rc[R_LIN + 5] : (r == 1 ? rc[R_LIN + 6] : 0.0);
    double z[4]; StateXV xs{w}; dyn_rows(w, *p, k, xs, z);
    return (b - z[r]) * inv_delta - rc[R_YE + r];
  }
  MPC_HD double init(int r) const {
    if (!act_init_bit(w, r)) return 0.0;
    const double* h = w.hdr();
    return (h[H_X0 + r] - w.rec(0)[R_XU + r]) * inv_delta - h[H_YI + r];
  }
  MPC_HD double row(int k, int g, int r) const {
    if (!((act_group_bits(w, k, g) >> r) & 1)) return 0.0;
    double lo, hi; group_bounds(*p, w.hdr(), k, g, lo, hi);
    StateXV xs{w};
    const double* rc = w.rec(k);
    double gv = group_g(k, g, xs), sv = rc[R_S + g];
    double ax = r == 0 ? gv - sv : (r == 1 ? gv + sv : sv);
    double b = r == 0 ? hi : (r == 1 ? lo : 0.0);
    return (b - ax) * inv_delta - rc[R_V + 3 * g + r];
  }
  // reduced slack rhs and pair-combined value after eliminating the slack
  MPC_HD void group_vals(int k, int g, const Mode& m, double& rr_s, double& gval) const {
    double v1 = row(k, g, 0), v2 = row(k, g, 1), v3 = row(k, g, 2);
    double ps = group_ps(*p, g);
    GroupCoef c = group_coef(m, ps, act_group_bits(w, k, g));
    rr_s = -ps * w.rec(k)[R_S + g] + (-v1 + v2 + v3);
    gval = v1 + v2 - c.csg * c.mss_inv * rr_s;
  }
  Mode mode;
  MPC_HD double grp(int k, int g) const { double a, b; group_vals(k, g, mode, a, b); return b; }
};

// S1(k): residual of the un-regularised polish KKT at the current (x, y) folded into the banded rhs
MPC_HD void polish_rhs_stage(const View& w, const Params& p, const Mode& m, int k) {
  const int N = w.N;
  double* rc = w.rec(k);
  PolishRP rp{w, &p, m.inv_delta, 0, m};
  const int ng = ngroups(N, k);
  for (int g = 0; g < ng; ++g) {
    double rr_s, gval; rp.group_vals(k, g, m, rr_s, gval);
    rc[R_ST + g] = rr_s;
  }
  double out[6];
  gather_xu(w, p, k, rp, out);
  const int nj = k < N ? 6 : 4;
  for (int j = 0; j < nj; ++j) {
    double qj = j < 4 ? rc[R_Q + j] : 0.0;
    w.bx_set(k, j, -qj - px_entry(p, k < N, j, rc + R_XU) + out[j]);
  }
  for (int j = nj; j < 6; ++j) w.bx_set(k, j, 0.0);
}
// S3a(k): ds and dy from the banded correction; y += dy (row-owned); ds left in R_ST
MPC_HD void polish_dual_stage(const View& w, const Params& p, const Mode& m, int k) {
  const int N = w.N;
  double* rc = w.rec(k);
  BxXV dx{w};
  StateXV xs{w};
  const int ng = ngroups(N, k);
  for (int g = 0; g < ng; ++g) {
    double lo, hi; group_bounds(p, w.hdr(), k, g, lo, hi);
    int bits = act_group_bits(w, k, g);
    GroupCoef c = group_coef(m, group_ps(p, g), bits);
    double gd = group_g(k, g, dx);
    double ds = (rc[R_ST + g] - c.csg * gd) * c.mss_inv;
    rc[R_ST + g] = ds;
    double gv = group_g(k, g, xs), sv = rc[R_S + g];
    double ax[3] = {gv - sv, gv + sv, sv};
    double ad[3] = {gd - ds, gd + ds, ds};
    double b[3] = {hi, lo, 0.0};
    for (int r = 0; r < 3; ++r)
      if ((bits >> r) & 1) rc[R_V + 3 * g + r] += m.inv_delta * (ad[r] - (b[r] - ax[r]));
  }
  if (k < N) {
    double ax[4], ad[4]; dyn_rows(w, p, k, xs, ax); dyn_rows(w, p, k, dx, ad);
    double b[4] = {rc[R_LIN + 5], rc[R_LIN + 6], 0.0, 0.0};
    // dyn_rows on the correction has no affine part: ad is linear in dx by construction
    for (int r = 0; r < 4; ++r)
      if (act_dyn_bit(w, k, r)) rc[R_YE + r] += m.inv_delta * (ad[r] - (b[r] - ax[r]));
  }
  if (k == 0) {
    double* h = w.hdr();
    for (int r = 0; r < 4; ++r)
      if (act_init_bit(w, r)) h[H_YI + r] += m.inv_delta * (dx(0, r) - (h[H_X0 + r] - xs(0, r)));
  }
}
// S3b(k): x += dx, s += ds
MPC_HD void polish_primal_stage(const View& w, int k) {
  double* rc = w.rec(k);
  const int nj = k < w.N ? 6 : 4;
  for (int j = 0; j < nj; ++j) rc[R_XU + j] += w.bx_get(k, j);
  const int ng = ngroups(w.N, k);
  for (int g = 0; g < ng; ++g) rc[R_S + g] += rc[R_ST + g];
}
// zero the primal/dual solution before the first polish step (sol = 0); activity masks are kept
MPC_HD void polish_zero_stage(const View& w, int k) {
  double* rc = w.rec(k);
  for (int j = 0; j < 11; ++j) rc[R_XU + j] = 0.0;
  for (int j = 0; j < 15; ++j) rc[R_V + j] = 0.0;
  for (int r = 0; r < 4; ++r) rc[R_YE + r] = 0.0;
  if (k == 0) for (int r = 0; r < 4; ++r) w.hdr()[H_YI + r] = 0.0;
}

// ----------------------------------------------------------------------------------------------
// Specialised polish / residual phases: same mathematics as polish_rhs_stage, polish_dual_stage and
// residual_stage above (kept as the readable statement), unrolled at compile time, with the slack-elimination
// coefficients tabulated per (group type, number of active rows) instead of divided out per use.
// ----------------------------------------------------------------------------------------------
struct PolConst {
  double invd;
  double lo[5], hi[5], ps[5];
  double mssinv[5][4];      // delta / ((ps + delta) delta + na), na = number of active rows of the group
  double up0, up1;
};
MPC_HD PolConst pol_const(const View& w, const Params& p, const Settings& s) {
  PolConst c;
  c.invd = 1.0 / s.delta;
  c.lo[0] = p.v_lo; c.hi[0] = p.v_hi;
  c.lo[1] = p.u_lo[0]; c.hi[1] = p.u_hi[0]; c.lo[2] = p.u_lo[1]; c.hi[2] = p.u_hi[1];
  c.lo[3] = p.du_lo[0]; c.hi[3] = p.du_hi[0]; c.lo[4] = p.du_lo[1]; c.hi[4] = p.du_hi[1];
  for (int g = 0; g < 5; ++g) {
    c.ps[g] = group_ps(p, g);
    for (int na = 0; na < 4; ++na) c.mssinv[g][na] = s.delta / ((c.ps[g] + s.delta) * s.delta + na);
  }
  c.up0 = w.hdr()[H_UPREV]; c.up1 = w.hdr()[H_UPREV + 1];
  return c;
}
MPC_HD double mss_pick(const PolConst& c, int g, int bits) {
  const int na = (bits & 1) + ((bits >> 1) & 1) + ((bits >> 2) & 1);
  return na == 0 ? c.mssinv[g][0] : (na == 1 ? c.mssinv[g][1] : (na == 2 ? c.mssinv[g][2] : c.mssinv[g][3]));
}
// value_i = act_i (e2_i / delta - y_i) of the three rows of one soft group; returns the pair-combined value after
// eliminating the slack and the reduced slack rhs
MPC_HD double pol_group_val(const PolConst& c, int g, int bits, double gv, double sv, double lo, double hi,
                            double y0, double y1, double y2, double& rr_s) {
  const double v0 = (bits & 1) ? fma(hi - (gv - sv), c.invd, -y0) : 0.0;
  const double v1 = (bits & 2) ? fma(lo - (gv + sv), c.invd, -y1) : 0.0;
  const double v2 = (bits & 4) ? fma(-sv, c.invd, -y2) : 0.0;
  rr_s = fma(-c.ps[g], sv, (v1 - v0) + v2);
  const double csg = (double)(((bits >> 1) & 1) - (bits & 1)) * c.invd;
  return (v0 + v1) - csg * mss_pick(c, g, bits) * rr_s;
}
MPC_HD void dyn_rows6(const double* lin, double dt, const double* x, const double* xn, double* z) {
  z[0] = xn[0] - (x[0] + lin[0] * x[2] + lin[1] * x[3]);
  z[1] = xn[1] - (x[1] + lin[2] * x[2] + lin[3] * x[3]);
  z[2] = xn[2] - (x[2] + lin[4] * x[5]);
  z[3] = xn[3] - (x[3] + dt * x[4]);
}

MPC_HD void polish_rhs_fast(const View& w, const Params& p, const PolConst& c, int k) {
  const int N = w.N;
  double* rc = w.rec(k);
  const bool reg = k < N;
  const int act = w.act()[k];
  double x[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) x[j] = rc[R_XU + j];
  double out[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  double ua = 0.0, ud = 0.0;
  if (k >= 1) {
    const double* rp = w.rec(k - 1);
    double xp[6], z[4];
#pragma unroll
    for (int j = 0; j < 6; ++j) xp[j] = rp[R_XU + j];
    ua = xp[4]; ud = xp[5];
    dyn_rows6(rp + R_LIN, p.dt, xp, x, z);
    const int ap = w.act()[k - 1];
    const double b[4] = {rp[R_LIN + 5], rp[R_LIN + 6], 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < 4; ++r) out[r] = ((ap >> (15 + r)) & 1) ? fma(b[r] - z[r], c.invd, -rp[R_YE + r]) : 0.0;
  } else {
    const double* h = w.hdr();
    const int ai = w.act()[N + 1];
#pragma unroll
    for (int r = 0; r < 4; ++r) out[r] = ((ai >> r) & 1) ? fma(h[H_X0 + r] - x[r], c.invd, -h[H_YI + r]) : 0.0;
  }
  const double off0 = k == 0 ? c.up0 : 0.0, off1 = k == 0 ? c.up1 : 0.0;
  const double gv[5] = {x[3], x[4], x[5], x[4] - (reg ? ua : 0.0), x[5] - (reg ? ud : 0.0)};
  const double offs[5] = {0.0, 0.0, 0.0, off0, off1};
  double G[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int g = 0; g < 5; ++g) {
    if (g == 0 || reg) {
      double rr_s;
      G[g] = pol_group_val(c, g, (act >> (3 * g)) & 7, gv[g], rc[R_S + g], c.lo[g] + offs[g], c.hi[g] + offs[g],
                           rc[R_V + 3 * g], rc[R_V + 3 * g + 1], rc[R_V + 3 * g + 2], rr_s);
      rc[R_ST + g] = rr_s;
    }
  }
  out[3] += G[0];
  if (reg) {
    const double* rn = w.rec(k + 1);
    double xn[6], z[4];
#pragma unroll
    for (int j = 0; j < 6; ++j) xn[j] = rn[R_XU + j];
    const double* lin = rc + R_LIN;
    dyn_rows6(lin, p.dt, x, xn, z);
    const double b[4] = {lin[5], lin[6], 0.0, 0.0};
    double d[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) d[r] = ((act >> (15 + r)) & 1) ? fma(b[r] - z[r], c.invd, -rc[R_YE + r]) : 0.0;
    out[0] -= d[0];
    out[1] -= d[1];
    out[2] -= lin[0] * d[0] + lin[2] * d[1] + d[2];
    out[3] -= lin[1] * d[0] + lin[3] * d[1] + d[3];
    out[4] = G[1] + G[3] - p.dt * d[3];
    out[5] = G[2] + G[4] - lin[4] * d[2];
    if (k + 1 < N) {
      const int an = w.act()[k + 1];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        double rr_s;
        out[4 + i] -= pol_group_val(c, 3 + i, (an >> (3 * (3 + i))) & 7, xn[4 + i] - x[4 + i], rn[R_S + 3 + i], c.lo[3 + i], c.hi[3 + i],
                                    rn[R_V + 3 * (3 + i)], rn[R_V + 3 * (3 + i) + 1], rn[R_V + 3 * (3 + i) + 2], rr_s);
      }
    }
  }
  int rev; double* bp = bx_ptr(w, k, rev);
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double val = 0.0;
    if (j < 4 || reg) val = -(j < 4 ? rc[R_Q + j] : 0.0) - px_entry(p, reg, j, x) + out[j];
    bp[rev ? 5 - j : j] = val;
  }
  if (k == mid_stage(N)) {
    double* bb = w.bottom().bx(N - k);
#pragma unroll
    for (int j = 0; j < 6; ++j) bb[j] = 0.0;
  }
}

MPC_HD void polish_dual_fast(const View& w, const Params& p, const PolConst& c, int k) {
  const int N = w.N;
  double* rc = w.rec(k);
  const bool reg = k < N;
  const int act = w.act()[k];
  double x[6], dx[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) x[j] = rc[R_XU + j];
  bx_load6(w, k, dx);
  double ua = 0.0, ud = 0.0, dua_ = 0.0, dud_ = 0.0;
  if (k >= 1 && reg) {
    const double* rp = w.rec(k - 1);
    ua = rp[R_XU + 4]; ud = rp[R_XU + 5];
    double dp[6]; bx_load6(w, k - 1, dp); dua_ = dp[4]; dud_ = dp[5];
  }
  const double off0 = k == 0 ? c.up0 : 0.0, off1 = k == 0 ? c.up1 : 0.0;
  const double offs[5] = {0.0, 0.0, 0.0, off0, off1};
  const double gv[5] = {x[3], x[4], x[5], x[4] - ua, x[5] - ud};
  const double gd[5] = {dx[3], dx[4], dx[5], dx[4] - dua_, dx[5] - dud_};
#pragma unroll
  for (int g = 0; g < 5; ++g) {
    if (g == 0 || reg) {
      const int bits = (act >> (3 * g)) & 7;
      const double csg = (double)(((bits >> 1) & 1) - (bits & 1)) * c.invd;
      const double ds = (rc[R_ST + g] - csg * gd[g]) * mss_pick(c, g, bits);
      rc[R_ST + g] = ds;
      const double sv = rc[R_S + g];
      if (bits & 1) rc[R_V + 3 * g] += c.invd * ((gd[g] - ds) - ((c.hi[g] + offs[g]) - (gv[g] - sv)));
      if (bits & 2) rc[R_V + 3 * g + 1] += c.invd * ((gd[g] + ds) - ((c.lo[g] + offs[g]) - (gv[g] + sv)));
      if (bits & 4) rc[R_V + 3 * g + 2] += c.invd * (ds - (0.0 - sv));
    }
  }
  if (reg) {
    const double* rn = w.rec(k + 1);
    double xn[6], dn[6], ax[4], ad[4];
#pragma unroll
    for (int j = 0; j < 6; ++j) xn[j] = rn[R_XU + j];
    bx_load6(w, k + 1, dn);
    const double* lin = rc + R_LIN;
    dyn_rows6(lin, p.dt, x, xn, ax);
    dyn_rows6(lin, p.dt, dx, dn, ad);
    const double b[4] = {lin[5], lin[6], 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if ((act >> (15 + r)) & 1) rc[R_YE + r] += c.invd * (ad[r] - (b[r] - ax[r]));
  }
  if (k == 0) {
    double* h = w.hdr();
    const int ai = w.act()[N + 1];
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if ((ai >> r) & 1) h[H_YI + r] += c.invd * (dx[r] - (h[H_X0 + r] - x[r]));
  }
}

// residual partial maxima (same seven entries as residual_stage); ymode 0: ADMM state, 1: polished pair
MPC_HD void residual_fast(const View& w, const Params& p, const IterConst& c, int ymode, int k, double* r) {
  const int N = w.N;
  const double* rc = w.rec(k);
  const bool reg = k < N;
  double x[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) x[j] = rc[R_XU + j];
  double aty[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  double ua = 0.0, ud = 0.0;
  if (k >= 1) {
    const double* rp = w.rec(k - 1);
    ua = rp[R_XU + 4]; ud = rp[R_XU + 5];
#pragma unroll
    for (int q = 0; q < 4; ++q) aty[q] = rp[R_YE + q];
  } else {
    const double* h = w.hdr();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      aty[q] = h[H_YI + q];
      const double ax = x[q], b = h[H_X0 + q];
      r[0] = dmax(r[0], fabs(ax - b)); r[1] = dmax(r[1], fabs(ax)); r[2] = dmax(r[2], fabs(b));
    }
  }
  const double off0 = k == 0 ? c.up0 : 0.0, off1 = k == 0 ? c.up1 : 0.0;
  const double offs[5] = {0.0, 0.0, 0.0, off0, off1};
  const double gv[5] = {x[3], x[4], x[5], x[4] - (reg ? ua : 0.0), x[5] - (reg ? ud : 0.0)};
  double Gy[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int g = 0; g < 5; ++g) {
    if (g == 0 || reg) {
      const double hi = c.hi[g] + offs[g], lo = c.lo[g] + offs[g];
      const double sv = rc[R_S + g];
      const double v0 = rc[R_V + 3 * g], v1 = rc[R_V + 3 * g + 1], v2 = rc[R_V + 3 * g + 2];
      const double ax0 = gv[g] - sv, ax1 = gv[g] + sv, ax2 = sv;
      double z0, z1, z2, y0, y1, y2;
      if (ymode) { z0 = dmin2(ax0, hi); z1 = dmax2(ax1, lo); z2 = dmax2(ax2, 0.0); y0 = v0; y1 = v1; y2 = v2; }
      else {
        z0 = dmin2(v0, hi); z1 = dmax2(v1, lo); z2 = dmax2(v2, 0.0);
        y0 = c.rho * (v0 - z0); y1 = c.rho * (v1 - z1); y2 = c.rho * (v2 - z2);
      }
      r[0] = dmax(r[0], dmax(fabs(ax0 - z0), dmax(fabs(ax1 - z1), fabs(ax2 - z2))));
      r[1] = dmax(r[1], dmax(fabs(ax0), dmax(fabs(ax1), fabs(ax2))));
      r[2] = dmax(r[2], dmax(fabs(z0), dmax(fabs(z1), fabs(z2))));
      Gy[g] = y0 + y1;
      const double px = c.ps_[g] * sv, ay = (y1 - y0) + y2;
      r[3] = dmax(r[3], fabs(px + ay)); r[4] = dmax(r[4], fabs(px)); r[5] = dmax(r[5], fabs(ay));
    }
  }
  aty[3] += Gy[0];
  if (reg) {
    const double* rn = w.rec(k + 1);
    const double* lin = rc + R_LIN;
    double xn[4], ax[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xn[j] = rn[R_XU + j];
    ax[0] = xn[0] - (x[0] + lin[0] * x[2] + lin[1] * x[3]);
    ax[1] = xn[1] - (x[1] + lin[2] * x[2] + lin[3] * x[3]);
    ax[2] = xn[2] - (x[2] + lin[4] * x[5]);
    ax[3] = xn[3] - (x[3] + p.dt * x[4]);
    const double b[4] = {lin[5], lin[6], 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      r[0] = dmax(r[0], fabs(ax[q] - b[q])); r[1] = dmax(r[1], fabs(ax[q])); r[2] = dmax(r[2], fabs(b[q]));
    }
    const double d0 = rc[R_YE + 0], d1 = rc[R_YE + 1], d2 = rc[R_YE + 2], d3 = rc[R_YE + 3];
    aty[0] -= d0;
    aty[1] -= d1;
    aty[2] -= lin[0] * d0 + lin[2] * d1 + d2;
    aty[3] -= lin[1] * d0 + lin[3] * d1 + d3;
    aty[4] = Gy[1] + Gy[3] - p.dt * d3;
    aty[5] = Gy[2] + Gy[4] - lin[4] * d2;
    if (k + 1 < N) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double v0 = rn[R_V + 3 * (3 + i)], v1 = rn[R_V + 3 * (3 + i) + 1];
        double y0, y1;
        if (ymode) { y0 = v0; y1 = v1; }
        else { y0 = c.rho * (v0 - dmin2(v0, c.hi[3 + i])); y1 = c.rho * (v1 - dmax2(v1, c.lo[3 + i])); }
        aty[4 + i] -= y0 + y1;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    if (j < 4 || reg) {
      const double px = px_entry(p, reg, j, x), qj = j < 4 ? rc[R_Q + j] : 0.0;
      r[3] = dmax(r[3], fabs(px + qj + aty[j])); r[4] = dmax(r[4], fabs(px)); r[5] = dmax(r[5], fabs(aty[j]));
      r[6] = dmax(r[6], fabs(qj));
    }
  }
}

// save / restore the iterate of one stage to the HBM warm-start slot ([stage][30] + tail)
MPC_HD void save_stage(const View& w, int k, double* g) {
  const double* rc = w.rec(k);
  double* o = g + 30 * k;
  for (int j = 0; j < 30; ++j) o[j] = rc[R_XU + j];
}
MPC_HD void load_stage(const View& w, int k, const double* g) {
  double* rc = w.rec(k);
  const double* o = g + 30 * k;
  for (int j = 0; j < 30; ++j) rc[R_XU + j] = o[j];
}

}  // namespace mpc
