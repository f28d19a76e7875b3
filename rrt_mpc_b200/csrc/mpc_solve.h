// mpc_solve.h — per-problem driver of the MPC tracking step (OSQP-equivalent ADMM + polish),
// templated on an execution policy `Exec` that says how stage-parallel phases, reductions and the
// sequential chain operations are run (CUDA warp in cudampc.cu; lane emulation in tests/emu).
//
// Follows /root/reference/src/control/mpc_controller.py:39-141 (what is solved, what is returned)
// with the solver restated from the published OSQP algorithm (see oracle/ for the CPU statement).
#pragma once
#include "mpc_core.h"

namespace mpc {

struct ProblemIO {
  const double* x0;      // [4]
  RefWin ref;            // window rows [x,y,yaw,v], time-major (mpc_controller.py:42)
  const double* u_prev;  // [2]
  double* warm;          // HBM slot, warm_size(N) doubles (ADMM iterate; also polish back-up)
  double* scratch;       // HBM slot, warm_size(N) doubles (polished-solution back-up between passes)
  double* fsave;         // HBM slot of the resident group, oe_doubles(N) doubles: the ADMM factor while a polish uses its place (or null)
  double* u0;            // [2]
  double* Xp;            // [4][N+1] state-major (visualization.py:244-245)
  double* Up;            // [2][N]
  int* status;           // OSQP status code
  int* iters;
  double* pri_res;
  double* dua_res;
  int* info;             // [4]: rho updates, factorisations, polish passes accepted, chain solves
};

struct Residuals { double pri, dua, eps_p, eps_d, sp, sd, nz, nq; };

template <class Exec>
MPC_HD Residuals compute_residuals(Exec& ex, const View& w, const Params& p, const Settings& s, const IterConst& ic, int ymode) {
  double r[7];
  ex.reduce_max(w.N + 1, r, 7, [&](int k, double* rl) { residual_fast(w, p, ic, ymode, k, rl); });
  Residuals o;
  o.pri = r[0]; o.dua = r[3];
  o.eps_p = s.eps_abs + s.eps_rel * dmax(r[1], r[2]);
  o.eps_d = s.eps_abs + s.eps_rel * dmax(dmax(r[4], r[5]), r[6]);
  o.sp = r[0] / (dmax(r[1], r[2]) + 1e-10);
  o.sd = r[3] / (dmax(dmax(r[4], r[5]), r[6]) + 1e-10);
  o.nz = dmax(r[1], r[2]); o.nq = dmax(dmax(r[4], r[5]), r[6]);
  return o;
}

// FORM of the ADMM phases.  FORM_GENERAL: one parity of stages at a time (any horizon, any number of warps per problem);
// FORM_SHORT: horizons with N+1 <= 32, one lane per stage (mpc_core.h admm_rhs_stage_short); FORM_PAIR: horizons with
// N+1 <= 64, one lane per pair of stages, update and next right-hand side in ONE pass (mpc_pair.h); FORM_REG: N+1 <= 64, two
// warps per problem, every stage record in the registers of its lane for a block of iterations (mpc_reg.h).  Separate
// instantiations, so a kernel carries one form.
enum { FORM_GENERAL = 0, FORM_SHORT = 1, FORM_PAIR = 2, FORM_REG = 3 };
template <class Exec>
MPC_HD void solve_problem_reg(Exec& ex, const View& w, const Params& p, const Settings& s, const ProblemIO& io);   // mpc_drv.h

template <int FORM, class Exec>
MPC_HD void solve_problem(Exec& ex, const View& w, const Params& p, const Settings& s, const ProblemIO& io) {
  if constexpr (FORM == FORM_REG) { solve_problem_reg(ex, w, p, s, io); return; }
  const int N = w.N;
  const int NS = N + 1;
  int n_rho = 0, n_fac = 0, n_solve = 0, n_pol = 0;

  // ---- load + linearise ------------------------------------------------------------------
  ex.tag(0);
  ex.single([&]() {
    double* h = w.hdr();
    for (int i = 0; i < 4; ++i) h[H_X0 + i] = io.x0[i];
    for (int i = 0; i < 2; ++i) h[H_UPREV + i] = io.u_prev ? io.u_prev[i] : 0.0;
    for (int i = 0; i < N + 2; ++i) w.act()[i] = 0;
    unwrap_window(io.ref, NS, w.scratch());               // scratch: bx area holds the unwrapped yaw column
  });
  ex.stages(NS, [&](int k) { setup_stage(w, p, k, io.ref, w.scratch()); });

  // ---- initial iterate -------------------------------------------------------------------
  double rho = s.rho0;
  // A slot is warm only if a previous solve left a usable iterate in it: the rho word doubles as the valid marker
  // (cudampc_create zeroes the buffer; a solve that ended on non-finite residuals stores 0).  Anything else starts cold.
  const double rho_slot = (s.warm_start && io.warm) ? io.warm[30 * NS + 4] : 0.0;
  if (rho_slot >= s.rho_min && rho_slot <= s.rho_max) {
    ex.stages(NS, [&](int k) { load_stage(w, k, io.warm); });
    ex.single([&]() { for (int r = 0; r < 4; ++r) w.hdr()[H_YI + r] = io.warm[30 * NS + r]; });
    rho = rho_slot;
  } else {
    ex.stages(NS, [&](int k) { cold_start_stage(w, p, k); });
  }

  // The ADMM matrix is (re)assembled and factorised at ONE place, the top of the iteration loop, whenever need_factor
  // is set (start, rho update, resume after a rejected polish); likewise the right-hand side phase.  One copy of that
  // code instead of four keeps the kernel's hot instruction footprint small.
  Mode mode = admm_mode(rho, s);
  IterConst ic = iter_const(w, p, s, rho);
  const OEView oe = oe_view(w);            // odd-even block solve of the ADMM iterations (mpc_oe.h)
  bool need_factor = true, need_rhs = true;

  // ---- ADMM + polish --------------------------------------------------------------------------
  // OSQP: iterate until the residual test passes, then polish once.  Two opt-in extensions:
  //  * polish_retry: a polish that is rejected or does not end on a KKT point (active set not identified) restores
  //    the ADMM iterate, tightens the internal tolerance 10x and iterates on;
  //  * early_polish: at a termination check whose guessed active set equals the one of the previous check, the
  //    polish is tried although the residual test has not passed yet; if it ends on a KKT point of a settled active
  //    set (sufficient for optimality of this strictly convex QP) the solve is finished, otherwise ADMM resumes.
  int status = STATUS_UNSOLVED;
  int it = 0;
  Residuals res; res.pri = res.dua = 1e300; res.eps_p = res.eps_d = 0.0; res.sp = res.sd = 0.0; res.nz = res.nq = 0.0;
  double pri = 1e300, dua = 1e300;
  Settings se = s;                         // effective tolerances (tightened by polish_retry)
  int retries = s.polish_retry;
  const bool can_polish = s.polish_passes > 0 && io.warm && io.scratch;

  auto save_iterate = [&]() {
    ex.stages(NS, [&](int k) { save_stage(w, k, io.warm); });
    ex.single([&]() {
      for (int r = 0; r < 4; ++r) io.warm[30 * NS + r] = w.hdr()[H_YI + r];
      io.warm[30 * NS + 4] = (res.pri < 1e300 && res.dua < 1e300) ? rho : 0.0;      // NaN / inf residuals: slot invalid
    });
  };
  auto restore_iterate = [&](const double* src) {
    ex.stages(NS, [&](int k) { load_stage(w, k, src); });
    ex.single([&]() { for (int r = 0; r < 4; ++r) w.hdr()[H_YI + r] = src[30 * NS + r]; });
  };
  // Polish from the ADMM iterate in shared memory (which must already be saved in io.warm).  Returns true if it ended on
  // a KKT point of a settled active set ("clean").  On return the state holds the last accepted polished pair
  // (n_pol > 0) or the restored ADMM iterate (n_pol == 0); pri/dua are updated accordingly.
  auto polish = [&](double pri0, double dua0) -> bool {
    const Mode pm = polish_mode(s);
    const PolConst pc = pol_const(w, p, s);
    bool settled = false;
    int acc = 0, rejected = 0;
    double pp = pri0, dd = dua0;
    for (int pass = 0; pass < s.polish_passes; ++pass) {
      ex.tag(9);
      if (pass > 0) {
        ex.stages(NS, [&](int k) { save_stage(w, k, io.scratch); });
        ex.single([&]() { for (int r = 0; r < 4; ++r) io.scratch[30 * NS + r] = w.hdr()[H_YI + r]; });
      }
      int changed = ex.any(NS, [&](int k) { return polish_activity_stage(w, p, rho, pass == 0, k); });
      if (pass > 0 && !changed) { settled = true; break; }
      ex.stages(NS, [&](int k) { polish_zero_stage(w, k); });
      ex.stages(NS, [&](int k) { assemble_stage(w, p, pm, k); });
      ex.tag(10); ex.factor(w); ++n_fac;
      ex.tag(11);
      for (int step = 0; step <= s.polish_refine_iter; ++step) {
        ex.stages(NS, [&](int k) { polish_rhs_fast(w, p, pc, k); });
        ex.solve(w); ++n_solve;
        ex.stages(NS, [&](int k) { polish_dual_fast(w, p, pc, k); });
        ex.stages(NS, [&](int k) { polish_primal_stage(w, k); });
      }
      ex.tag(12);
      Residuals rp = compute_residuals(ex, w, p, s, ic, 1);
      bool ok;
      if (pass == 0) {
        ok = (rp.pri < pp && rp.dua < dd) || (rp.pri < pp && dd < 1e-10) || (rp.dua < dd && pp < 1e-10);
      } else {
        ok = rp.pri <= dmax(10.0 * pp, 1e-9 * dmax(1.0, res.nz)) && rp.dua <= dmax(10.0 * dd, 1e-9 * dmax(1.0, res.nq));
      }
      if (ok) { pp = rp.pri; dd = rp.dua; acc = pass + 1; }
      else { rejected = pass == 0 ? 1 : 2; break; }
    }
    if (rejected == 2) {               // a later pass was rejected: back to the previous accepted polished pair
      ex.stages(NS, [&](int k) { load_stage(w, k, io.scratch); });
      ex.single([&]() { for (int r = 0; r < 4; ++r) w.hdr()[H_YI + r] = io.scratch[30 * NS + r]; });
    }
    if (acc > 0) { pri = pp; dua = dd; }
    n_pol = acc;
    return acc > 0 && (settled || s.polish_passes == 1) && pp <= 1e-9 * dmax(1.0, res.nz) && dd <= 1e-9 * dmax(1.0, res.nq);
  };
  auto resume_admm = [&]() {               // the polish replaced the factor and (maybe) the state
    n_pol = 0;
    ex.tag(13);
    ex.stages(NS, [&](int k) { load_stage(w, k, io.warm); });
    ex.single([&]() { for (int r = 0; r < 4; ++r) w.hdr()[H_YI + r] = io.warm[30 * NS + r]; });
    if (io.fsave) { ex.stages(oe_doubles(N), [&](int i) { oe_restore_word(w, i, io.fsave); }); need_rhs = true; }   // rho is unchanged: the saved factor is valid
    else need_factor = true;
  };

  // NOTE: every multi-statement lambda above has exactly ONE call site below, so that it is inlined and the solver
  // state it captures stays in registers (a second call site makes nvcc outline it and spill the captures to local memory).
  bool finished = false, need_restore = false;
  // what follows the update of iteration `it`: termination check, early / final polish, rho adaptation
  auto after_update = [&]() {
    const bool last = it >= s.max_iter;
    const bool check = last || ((s.check_termination > 0) && (it % s.check_termination == 0));
    const bool adapt = !last && s.adaptive_rho && (s.adaptive_rho_interval > 0) && (it % s.adaptive_rho_interval == 0);
    if (check || adapt) {
      ex.tag(3); res = compute_residuals(ex, w, p, se, ic, 0);
      pri = res.pri; dua = res.dua;
      if (check) {
        const bool converged = res.pri <= res.eps_p && res.dua <= res.eps_d;
        if (converged && status == STATUS_UNSOLVED) status = STATUS_SOLVED;
        if (last && status == STATUS_UNSOLVED) {
          // iteration limit: OSQP's "solved inaccurate" test with 10x looser tolerances (user eps, not the tightened ones)
          const double ep10 = 10.0 * s.eps_abs + 10.0 * s.eps_rel * res.nz, ed10 = 10.0 * s.eps_abs + 10.0 * s.eps_rel * res.nq;
          status = (res.pri <= ep10 && res.dua <= ed10) ? STATUS_SOLVED_INACCURATE : STATUS_MAX_ITER;
        }
        bool attempt = converged || (last && status == STATUS_SOLVED);
        if (!attempt && !last && s.early_polish && can_polish) {
          ex.tag(15);
          const int changed = ex.any(NS, [&](int k) { return activity_probe_stage(w, p, k); });
          attempt = (!changed || s.early_polish >= 2) && it >= s.early_polish_start;
        }
        if (attempt || last) {
          ex.tag(8);
          if (io.warm) save_iterate();
          if (attempt && can_polish) {
            // the polish factorises in the place of the ADMM factor: keep a copy unless this is certainly the last polish
            if (io.fsave && !last && !(converged && retries <= 0)) ex.stages(oe_doubles(N), [&](int i) { oe_save_word(w, i, io.fsave); });
            const bool clean = polish(res.pri, res.dua);
            if (clean) { status = STATUS_SOLVED; finished = true; }
            else if (last || (converged && retries <= 0)) {                      // keep what the polish gave (OSQP behaviour)
              finished = true;
              if (n_pol == 0) need_restore = true;
            }
            else {
              if (converged) { --retries; se.eps_abs *= 0.1; se.eps_rel *= 0.1; }
              resume_admm();
            }
          } else {
            finished = true;
          }
        }
      }
      if (adapt && !finished) {
        double rho_new = rho * sqrt(res.sp / (res.sd + 1e-10));
        rho_new = fmin(fmax(rho_new, s.rho_min), s.rho_max);
        if (rho_new > rho * s.adaptive_rho_tolerance || rho_new < rho / s.adaptive_rho_tolerance) {
          ex.stages(NS, [&](int k) { rescale_v_stage(w, p, rho, rho_new, k); });
          rho = rho_new; ++n_rho;
          need_factor = true;
        }
      }
    }
  };
  // One ADMM iteration = right-hand side (odd stages, then even), sweeps over the even stages, update (odd, then even).
  // The two parities of a phase share ONE copy of the phase's code: the iteration loop has to stay inside the instruction
  // cache (32 KB of L1.5 per SM) also when the resident problems are at different places of the solve.
  while (!finished) {
    if (need_factor) {
      mode = admm_mode(rho, s);
      ex.tag(5); ex.oe_factor(w, p, mode, oe); ++n_fac;
      ic = iter_const(w, p, s, rho);
      need_factor = false;
      need_rhs = true;
    }
    ex.tag(6);
    if constexpr (FORM == FORM_SHORT) {
      ex.stages(NS, [&](int k) { admm_rhs_stage_short(w, p, ic, oe, k); });
      ex.stages_par(NS, 0, [&](int k) { oe_even_fixup(w, p, ic, k); });
    } else if (FORM != FORM_PAIR || need_rhs) {      // pair form: only the first right-hand side after a (re)factorisation
#pragma unroll 1
      for (int par = 1; par >= 0; --par) ex.stages_par(NS, par, [&](int k) { admm_rhs_stage_oe(w, p, ic, oe, k); });
      need_rhs = false;
    }
    ++it;
    ex.tag(16);
    ex.oe_forward(w, oe);
    ex.tag(17);
    ex.stages_par(NS, 0, [&](int k) { oe_diag_stage(w, oe, k); });
    ex.tag(18);
    ex.oe_backward(w, oe); ++n_solve;
    ex.tag(2);
    if constexpr (FORM == FORM_PAIR) {
      ex.pair_pass(w, p, ic, oe);          // update of this iteration + right-hand side of the next one
    } else if constexpr (FORM == FORM_SHORT) {
      ex.stages_par(NS, 1, [&](int k) { oe_expand_odd(w, p, ic, oe, k); });
      ex.stages(NS, [&](int k) { admm_update_stage_short(w, p, ic, k); });
    } else {
#pragma unroll 1
      for (int par = 1; par >= 0; --par) ex.stages_par(NS, par, [&](int k) { admm_update_stage_oe(w, p, ic, oe, k); });
    }
#ifdef MPC_POISON      // dev builds (tools/poison_check.py): after the update every right-hand side / solution row is dead until the next
                       // right-hand side pass rewrites it - a phase that read one earlier (a missing barrier between the warps of a
                       // group, a wrong row index) would now read a NaN and the solve could not reach its optimum
    ex.stages(NS, [&](int k) { double* r_ = w.nx(k); const double nan_ = (w.N - w.N) / (double)(w.N - w.N); for (int j = 0; j < BXS; ++j) r_[j] = nan_; });
#endif
    after_update();
  }
  ex.tag(7);
  if (need_restore) restore_iterate(io.warm);   // polish rejected outright: the answer is the ADMM iterate

  // ---- outputs (mpc_controller.py:141: U[:,0], X (4,N+1), U (2,N)) --------------------------
  ex.stages(NS, [&](int k) {
    const double* rc = w.rec(k);
    for (int j = 0; j < 4; ++j) io.Xp[j * NS + k] = rc[R_XU + j];
    if (k < N) { io.Up[k] = rc[R_XU + 4]; io.Up[N + k] = rc[R_XU + 5]; }
    if (k == 0) {
      io.u0[0] = rc[R_XU + 4]; io.u0[1] = rc[R_XU + 5];
      *io.status = status; *io.iters = it;
      if (io.pri_res) *io.pri_res = pri;
      if (io.dua_res) *io.dua_res = dua;
      if (io.info) { io.info[0] = n_rho; io.info[1] = n_fac; io.info[2] = n_pol; io.info[3] = n_solve; }
    }
  });
}

}  // namespace mpc

#include "mpc_drv.h"
