/* mpc_oracle.c — CPU oracle of the MPC tracking step (TEST INFRASTRUCTURE; never linked into or called by
 * the product path — only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load it).
 *
 * A plain-C restatement of the reference's algorithm for this path:
 *   f_discrete / linearize ............ /root/reference/src/control/vehicle_model.py:11-45
 *   horizon QP (literal form) .......... /root/reference/src/control/mpc_controller.py:47-117
 *   solver call, status mapping ........ /root/reference/src/control/mpc_controller.py:119-141
 *   relaxation retry, closed loop ...... /root/reference/src/pipeline/control_stage.py:33-56,74-157
 *
 * The arithmetic of the solve lives in third-party packages that are NOT under /root/reference and cannot be
 * installed offline: cvxpy (>=1.4,<2.0) and osqp (>=0.6.5, bundling qdldl), requirements.txt:6-7.  What follows
 * restates the published OSQP algorithm (Stellato, Banjac, Goulart, Bemporad, Boyd: "OSQP: an operator
 * splitting solver for quadratic programs", Math. Prog. Comp. 12, 2020): Ruiz equilibration + cost scaling,
 * rho vector (x1e3 on equality rows), quasi-definite KKT solved by a sparse LDL', over-relaxed ADMM, unscaled
 * residual termination, adaptive rho (fixed iteration interval), and polish with iterative refinement.
 * PARITY UNPINNED for the solver part: the reference's tests hold no solver-dependent number
 * (tests/test_mpc_controller.py:7-17).  The optimum itself is certified independently
 * (oracle/mpc_numpy.py: solve_kkt_newton) and the linearisation is pinned against the real reference
 * function through tests/golden/.
 *
 * The QP is kept in generic sparse (CSC) form and factorised by a generic sparse LDL' — deliberately a
 * different route from the CUDA kernel's slack-eliminated banded solve.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OSQP_INFTY 1e30
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4
#define RHO_TOL 1e-4

typedef struct {
  double wheelbase_px, dt;
  int32_t horizon, _pad;
  double q[16], r[4], q_terminal[16];
  double u_bounds[4];  /* a_lo, a_hi, d_lo, d_hi */
  double v_bounds[2];
  double du_bounds[4]; /* da_lo, da_hi, dd_lo, dd_hi */
  double slack_velocity, slack_input, slack_rate;
} oracle_params;

typedef struct {
  double eps_abs, eps_rel, rho, alpha, sigma, adaptive_rho_tolerance, rho_eq_factor, rho_min, rho_max, delta;
  int32_t max_iter, check_termination, adaptive_rho, adaptive_rho_interval;
  int32_t polish_passes, polish_refine_iter, scaling, z0_projected;
} oracle_settings;

enum { ST_SOLVED = 1, ST_INACCURATE = 2, ST_MAX_ITER = -2, ST_UNSOLVED = -10 };

/* ------------------------------------------------------------------------------------------------ */
/* vehicle model (vehicle_model.py:11-45)                                                           */
/* ------------------------------------------------------------------------------------------------ */
void oracle_f_discrete(const double* x, const double* u, double dt, double L, double* out) {
  double yaw = x[2], v = x[3], beta = 0.0;
  out[0] = x[0] + dt * v * cos(yaw + beta);
  out[1] = x[1] + dt * v * sin(yaw + beta);
  out[2] = yaw + dt * (v / L) * tan(u[1]);
  out[3] = v + dt * u[0];
}

void oracle_linearize(const double* x, const double* u, double dt, double L, double* A, double* B, double* fx) {
  double yaw = x[2], v = x[3], delta = u[1];
  double c = cos(yaw), s = sin(yaw), tan_d = tan(delta);
  double sec2_d = 1.0 / (cos(delta) * cos(delta) + 1e-9);
  memset(A, 0, 16 * sizeof(double));
  memset(B, 0, 8 * sizeof(double));
  A[0] = A[5] = A[10] = A[15] = 1.0;
  A[2] = -dt * v * s; A[3] = dt * c; A[6] = dt * v * c; A[7] = dt * s;
  A[11] = dt * (1.0 / L) * tan_d;
  B[6] = dt;
  B[5] = dt * (v / L) * sec2_d;
  oracle_f_discrete(x, u, dt, L, fx);
}

/* np.unwrap of the yaw column of a (rows,4) window copy (mpc_controller.py:59-60) */
static double np_mod(double a, double b) {
  double r = fmod(a, b);
  if (r != 0.0 && ((b < 0.0) != (r < 0.0))) r += b;
  return r;
}
static void unwrap_col(double* ref, int rows) {
  const double PI = 3.141592653589793;
  double cum = 0.0, prev = ref[2];
  for (int k = 1; k < rows; ++k) {
    double cur = ref[4 * k + 2], dd = cur - prev;
    double ddmod = np_mod(dd + PI, 2.0 * PI) - PI;
    if (ddmod == -PI && dd > 0.0) ddmod = PI;
    double corr = ddmod - dd;
    if (fabs(dd) < PI) corr = 0.0;
    cum += corr;
    ref[4 * k + 2] = cur + cum;
    prev = cur;
  }
}

/* (A_k,B_k,c_k), k<N, as MPCController.solve computes them (mpc_controller.py:65-70,108-109) */
void oracle_linearize_window(const oracle_params* p, const double* ref_in, double* refu, double* As, double* Bs, double* cs) {
  int N = p->horizon;
  memcpy(refu, ref_in, sizeof(double) * 4 * (N + 1));
  unwrap_col(refu, N + 1);
  const double* xlin = refu;
  double ulin[2] = {0.0, 0.0};
  for (int k = 0; k < N; ++k) {
    double* A = As + 16 * k; double* B = Bs + 8 * k; double fx[4];
    oracle_linearize(xlin, ulin, p->dt, p->wheelbase_px, A, B, fx);
    for (int i = 0; i < 4; ++i) {
      double ax = 0.0, bu = 0.0;
      for (int j = 0; j < 4; ++j) ax += A[4 * i + j] * xlin[j];
      for (int j = 0; j < 2; ++j) bu += B[2 * i + j] * ulin[j];
      cs[4 * k + i] = fx[i] - ax - bu;
    }
    xlin = refu + 4 * k;
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* tiny sparse toolkit: triplets -> CSC                                                             */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { int m, n, nz, cap; int *i, *j; double* x; } trip;
typedef struct { int m, n; int *p, *i; double* x; } csc;

static void trip_init(trip* t, int m, int n, int cap) {
  t->m = m; t->n = n; t->nz = 0; t->cap = cap;
  t->i = (int*)malloc(sizeof(int) * cap); t->j = (int*)malloc(sizeof(int) * cap); t->x = (double*)malloc(sizeof(double) * cap);
}
static void trip_add(trip* t, int i, int j, double x) {
  if (t->nz == t->cap) {
    t->cap *= 2;
    t->i = (int*)realloc(t->i, sizeof(int) * t->cap); t->j = (int*)realloc(t->j, sizeof(int) * t->cap);
    t->x = (double*)realloc(t->x, sizeof(double) * t->cap);
  }
  t->i[t->nz] = i; t->j[t->nz] = j; t->x[t->nz] = x; t->nz++;
}
static void trip_free(trip* t) { free(t->i); free(t->j); free(t->x); }
static void csc_free(csc* a) { free(a->p); free(a->i); free(a->x); a->p = a->i = NULL; a->x = NULL; }

/* compress (entries within a column sorted by row, duplicates summed) */
static void trip_to_csc(const trip* t, csc* a) {
  int n = t->n;
  a->m = t->m; a->n = n;
  a->p = (int*)calloc(n + 1, sizeof(int));
  a->i = (int*)malloc(sizeof(int) * (t->nz > 0 ? t->nz : 1));
  a->x = (double*)malloc(sizeof(double) * (t->nz > 0 ? t->nz : 1));
  for (int k = 0; k < t->nz; ++k) a->p[t->j[k] + 1]++;
  for (int j = 0; j < n; ++j) a->p[j + 1] += a->p[j];
  int* next = (int*)malloc(sizeof(int) * (n + 1));
  memcpy(next, a->p, sizeof(int) * (n + 1));
  for (int k = 0; k < t->nz; ++k) { int q = next[t->j[k]]++; a->i[q] = t->i[k]; a->x[q] = t->x[k]; }
  free(next);
  /* sort each column by row (insertion sort; columns are short) and merge duplicates */
  int w = 0;
  int* np_ = (int*)calloc(n + 1, sizeof(int));
  for (int j = 0; j < n; ++j) {
    int s = a->p[j], e = a->p[j + 1];
    for (int q = s + 1; q < e; ++q) {
      int ri = a->i[q]; double rx = a->x[q]; int z = q - 1;
      while (z >= s && a->i[z] > ri) { a->i[z + 1] = a->i[z]; a->x[z + 1] = a->x[z]; --z; }
      a->i[z + 1] = ri; a->x[z + 1] = rx;
    }
    int start = w;
    for (int q = s; q < e; ++q) {
      if (w > start && a->i[w - 1] == a->i[q]) a->x[w - 1] += a->x[q];
      else { a->i[w] = a->i[q]; a->x[w] = a->x[q]; ++w; }
    }
    np_[j + 1] = w;
  }
  memcpy(a->p, np_, sizeof(int) * (n + 1));
  free(np_);
}

static void csc_mv(const csc* a, const double* x, double* y) { /* y = A x */
  for (int i = 0; i < a->m; ++i) y[i] = 0.0;
  for (int j = 0; j < a->n; ++j) for (int q = a->p[j]; q < a->p[j + 1]; ++q) y[a->i[q]] += a->x[q] * x[j];
}
static void csc_mtv(const csc* a, const double* x, double* y) { /* y = A' x */
  for (int j = 0; j < a->n; ++j) { double s = 0.0; for (int q = a->p[j]; q < a->p[j + 1]; ++q) s += a->x[q] * x[a->i[q]]; y[j] = s; }
}
/* y = P x with P stored as upper triangle */
static void sym_mv(const csc* p, const double* x, double* y) {
  for (int i = 0; i < p->n; ++i) y[i] = 0.0;
  for (int j = 0; j < p->n; ++j)
    for (int q = p->p[j]; q < p->p[j + 1]; ++q) {
      int i = p->i[q];
      y[i] += p->x[q] * x[j];
      if (i != j) y[j] += p->x[q] * x[i];
    }
}
static double ninf(const double* v, int n) { double m = 0.0; for (int i = 0; i < n; ++i) { double a = fabs(v[i]); if (a > m) m = a; } return m; }

/* ------------------------------------------------------------------------------------------------ */
/* sparse LDL' of a symmetric quasi-definite matrix given by its upper triangle (CSC, sorted)        */
/* up-looking: row k of L is the solution of a sparse triangular system whose pattern is the reach   */
/* of column k's entries in the elimination tree.                                                   */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { int n; int *parent, *lp, *li, *lnz, *flag, *pattern; double *lx, *d, *dinv, *y; } ldl;

static void ldl_free(ldl* f) {
  free(f->parent); free(f->lp); free(f->li); free(f->lnz); free(f->flag); free(f->pattern);
  free(f->lx); free(f->d); free(f->dinv); free(f->y);
  memset(f, 0, sizeof *f);
}
static void ldl_symbolic(ldl* f, const csc* K) {
  int n = K->n;
  memset(f, 0, sizeof *f);
  f->n = n;
  f->parent = (int*)malloc(sizeof(int) * n); f->lnz = (int*)calloc(n, sizeof(int));
  f->flag = (int*)malloc(sizeof(int) * n); f->pattern = (int*)malloc(sizeof(int) * n);
  f->lp = (int*)malloc(sizeof(int) * (n + 1));
  for (int k = 0; k < n; ++k) {
    f->parent[k] = -1; f->flag[k] = k;
    for (int q = K->p[k]; q < K->p[k + 1]; ++q) {
      int i = K->i[q];
      if (i >= k) continue;
      for (; f->flag[i] != k; i = f->parent[i]) {
        if (f->parent[i] == -1) f->parent[i] = k;
        f->lnz[i]++;
        f->flag[i] = k;
      }
    }
  }
  f->lp[0] = 0;
  for (int k = 0; k < n; ++k) f->lp[k + 1] = f->lp[k] + f->lnz[k];
  int nl = f->lp[n] > 0 ? f->lp[n] : 1;
  f->li = (int*)malloc(sizeof(int) * nl); f->lx = (double*)malloc(sizeof(double) * nl);
  f->d = (double*)malloc(sizeof(double) * n); f->dinv = (double*)malloc(sizeof(double) * n);
  f->y = (double*)calloc(n, sizeof(double));
}
static int ldl_numeric(ldl* f, const csc* K) {
  int n = f->n;
  for (int k = 0; k < n; ++k) f->lnz[k] = 0;
  for (int k = 0; k < n; ++k) {
    int top = n;
    f->flag[k] = k;
    f->y[k] = 0.0;
    for (int q = K->p[k]; q < K->p[k + 1]; ++q) {
      int i = K->i[q];
      if (i > k) continue;
      f->y[i] += K->x[q];
      int len = 0;
      for (; f->flag[i] != k; i = f->parent[i]) { f->pattern[len++] = i; f->flag[i] = k; }
      while (len > 0) f->pattern[--top] = f->pattern[--len];
    }
    double dk = f->y[k];
    f->y[k] = 0.0;
    for (; top < n; ++top) {
      int i = f->pattern[top];
      double yi = f->y[i];
      f->y[i] = 0.0;
      int e = f->lp[i] + f->lnz[i];
      for (int q = f->lp[i]; q < e; ++q) f->y[f->li[q]] -= f->lx[q] * yi;
      double lki = yi * f->dinv[i];
      dk -= lki * yi;
      f->li[e] = k; f->lx[e] = lki; f->lnz[i]++;
    }
    if (dk == 0.0) return -1;
    f->d[k] = dk; f->dinv[k] = 1.0 / dk;
  }
  return 0;
}
static void ldl_solve(const ldl* f, double* x) {
  int n = f->n;
  for (int j = 0; j < n; ++j) { double xj = x[j]; for (int q = f->lp[j]; q < f->lp[j + 1]; ++q) x[f->li[q]] -= f->lx[q] * xj; }
  for (int j = 0; j < n; ++j) x[j] *= f->dinv[j];
  for (int j = n - 1; j >= 0; --j) { double s = x[j]; for (int q = f->lp[j]; q < f->lp[j + 1]; ++q) s -= f->lx[q] * x[f->li[q]]; x[j] = s; }
}

/* ------------------------------------------------------------------------------------------------ */
/* QP assembly, literal standard form of mpc_controller.py:47-117, stage-interleaved ordering        */
/*   z = [x_k(4) u_k(2) sv_k su_k(2) sdu_k(2)]_{k<N} , [x_N(4) sv_N];  n = 11N+5                       */
/*   rows = init(4), then per stage k<N: dyn(4) v(3) u(6) du(6), then v_N(3);  m = 19N+7               */
/* ------------------------------------------------------------------------------------------------ */
static int ix(int k, int i) { return 11 * k + i; }
static int iu(int k, int i) { return 11 * k + 4 + i; }
static int isv(int N, int k) { return k < N ? 11 * k + 6 : 11 * N + 4; }
static int isu(int k, int i) { return 11 * k + 7 + i; }
static int isdu(int k, int i) { return 11 * k + 9 + i; }
static int r_dyn(int k, int i) { return 4 + 19 * k + i; }
static int r_v(int N, int k, int j) { return 4 + 19 * k + (k < N ? 4 : 0) + j; }
static int r_u(int k, int i, int j) { return 4 + 19 * k + 7 + 3 * i + j; }
static int r_du(int k, int i, int j) { return 4 + 19 * k + 13 + 3 * i + j; }

typedef struct { int n, m, N; csc P, A; double *q, *l, *u; double* refu; double *As, *Bs, *cs; } qp_t;

static void qp_free(qp_t* Q) { csc_free(&Q->P); csc_free(&Q->A); free(Q->q); free(Q->l); free(Q->u); free(Q->refu); free(Q->As); free(Q->Bs); free(Q->cs); }

static void qp_build(const oracle_params* p, const double* x0, const double* ref, const double* u_prev, qp_t* Q) {
  int N = p->horizon, n = 11 * N + 5, m = 19 * N + 7;
  Q->n = n; Q->m = m; Q->N = N;
  Q->refu = (double*)malloc(sizeof(double) * 4 * (N + 1));
  Q->As = (double*)malloc(sizeof(double) * 16 * N); Q->Bs = (double*)malloc(sizeof(double) * 8 * N); Q->cs = (double*)malloc(sizeof(double) * 4 * N);
  oracle_linearize_window(p, ref, Q->refu, Q->As, Q->Bs, Q->cs);
  Q->q = (double*)calloc(n, sizeof(double));
  Q->l = (double*)malloc(sizeof(double) * m); Q->u = (double*)malloc(sizeof(double) * m);
  for (int i = 0; i < m; ++i) { Q->l[i] = -OSQP_INFTY; Q->u[i] = OSQP_INFTY; }
  double up0 = u_prev ? u_prev[0] : 0.0, up1 = u_prev ? u_prev[1] : 0.0;
  trip tp, ta;
  trip_init(&tp, n, n, 32 * (N + 1)); trip_init(&ta, m, n, 48 * (N + 1));
  /* cost: quad_form(e,W) = e'We = 1/2 e'(W+W')e ; P = W+W' (upper triangle), q = -(W+W') ref */
  for (int k = 0; k <= N; ++k) {
    const double* W = k < N ? p->q : p->q_terminal;
    for (int i = 0; i < 4; ++i) {
      double qi = 0.0;
      for (int j = 0; j < 4; ++j) {
        double s = W[4 * i + j] + W[4 * j + i];
        qi -= s * Q->refu[4 * k + j];
        if (j >= i && s != 0.0) trip_add(&tp, ix(k, i), ix(k, j), s);
      }
      Q->q[ix(k, i)] = qi;
    }
    trip_add(&tp, isv(N, k), isv(N, k), 2.0 * p->slack_velocity);
  }
  for (int k = 0; k < N; ++k)
    for (int i = 0; i < 2; ++i) {
      for (int j = i; j < 2; ++j) { double s = p->r[2 * i + j] + p->r[2 * j + i]; if (s != 0.0) trip_add(&tp, iu(k, i), iu(k, j), s); }
      trip_add(&tp, isu(k, i), isu(k, i), 2.0 * p->slack_input);
      trip_add(&tp, isdu(k, i), isdu(k, i), 2.0 * p->slack_rate);
    }
  /* X_0 = x0 */
  for (int i = 0; i < 4; ++i) { trip_add(&ta, i, ix(0, i), 1.0); Q->l[i] = Q->u[i] = x0[i]; }
  for (int k = 0; k < N; ++k) {
    for (int i = 0; i < 4; ++i) { /* X_{k+1} - A X_k - B U_k = c */
      int r = r_dyn(k, i);
      trip_add(&ta, r, ix(k + 1, i), 1.0);
      for (int j = 0; j < 4; ++j) if (Q->As[16 * k + 4 * i + j] != 0.0) trip_add(&ta, r, ix(k, j), -Q->As[16 * k + 4 * i + j]);
      for (int j = 0; j < 2; ++j) if (Q->Bs[8 * k + 2 * i + j] != 0.0) trip_add(&ta, r, iu(k, j), -Q->Bs[8 * k + 2 * i + j]);
      Q->l[r] = Q->u[r] = Q->cs[4 * k + i];
    }
    for (int i = 0; i < 2; ++i) {
      int r;
      r = r_u(k, i, 0); trip_add(&ta, r, iu(k, i), 1.0); trip_add(&ta, r, isu(k, i), -1.0); Q->u[r] = p->u_bounds[2 * i + 1];
      r = r_u(k, i, 1); trip_add(&ta, r, iu(k, i), 1.0); trip_add(&ta, r, isu(k, i), 1.0); Q->l[r] = p->u_bounds[2 * i];
      r = r_u(k, i, 2); trip_add(&ta, r, isu(k, i), 1.0); Q->l[r] = 0.0;
      double off = k == 0 ? (i == 0 ? up0 : up1) : 0.0;
      r = r_du(k, i, 0); trip_add(&ta, r, iu(k, i), 1.0); trip_add(&ta, r, isdu(k, i), -1.0); if (k > 0) trip_add(&ta, r, iu(k - 1, i), -1.0);
      Q->u[r] = p->du_bounds[2 * i + 1] + off;
      r = r_du(k, i, 1); trip_add(&ta, r, iu(k, i), 1.0); trip_add(&ta, r, isdu(k, i), 1.0); if (k > 0) trip_add(&ta, r, iu(k - 1, i), -1.0);
      Q->l[r] = p->du_bounds[2 * i] + off;
      r = r_du(k, i, 2); trip_add(&ta, r, isdu(k, i), 1.0); Q->l[r] = 0.0;
    }
  }
  for (int k = 0; k <= N; ++k) {
    int r;
    r = r_v(N, k, 0); trip_add(&ta, r, ix(k, 3), 1.0); trip_add(&ta, r, isv(N, k), -1.0); Q->u[r] = p->v_bounds[1];
    r = r_v(N, k, 1); trip_add(&ta, r, ix(k, 3), 1.0); trip_add(&ta, r, isv(N, k), 1.0); Q->l[r] = p->v_bounds[0];
    r = r_v(N, k, 2); trip_add(&ta, r, isv(N, k), 1.0); Q->l[r] = 0.0;
  }
  trip_to_csc(&tp, &Q->P); trip_to_csc(&ta, &Q->A);
  trip_free(&tp); trip_free(&ta);
}

/* ------------------------------------------------------------------------------------------------ */
/* OSQP restatement                                                                                 */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
  int n, m;
  csc P, A;              /* scaled */
  double *q, *l, *u;     /* scaled */
  double *D, *E, *Dinv, *Einv, c, cinv;
  int* perm_var; int* perm_row;   /* position of variable / row in the KKT ordering */
  csc K; int* kdiag_row;          /* KKT upper triangle in the permuted order; index of the -1/rho diagonal entry of each row */
  int* kdiag_var;
  ldl F; int symbolic_done;
} work_t;

static double limit_scaling(double v) { return v < MIN_SCALING ? 1.0 : (v > MAX_SCALING ? MAX_SCALING : v); }

static void scale_problem(work_t* W, int iters) {
  int n = W->n, m = W->m;
  double* dt = (double*)malloc(sizeof(double) * n); double* et = (double*)malloc(sizeof(double) * m);
  for (int it = 0; it < iters; ++it) {
    for (int j = 0; j < n; ++j) dt[j] = 0.0;
    for (int i = 0; i < m; ++i) et[i] = 0.0;
    for (int j = 0; j < n; ++j) {      /* column norms of [P; A] (P symmetric, stored upper) */
      for (int q = W->P.p[j]; q < W->P.p[j + 1]; ++q) {
        double a = fabs(W->P.x[q]); int i = W->P.i[q];
        if (a > dt[j]) dt[j] = a;
        if (a > dt[i]) dt[i] = a;
      }
      for (int q = W->A.p[j]; q < W->A.p[j + 1]; ++q) {
        double a = fabs(W->A.x[q]); int i = W->A.i[q];
        if (a > dt[j]) dt[j] = a;
        if (a > et[i]) et[i] = a;
      }
    }
    for (int j = 0; j < n; ++j) dt[j] = 1.0 / sqrt(limit_scaling(dt[j]));
    for (int i = 0; i < m; ++i) et[i] = 1.0 / sqrt(limit_scaling(et[i]));
    for (int j = 0; j < n; ++j) {
      for (int q = W->P.p[j]; q < W->P.p[j + 1]; ++q) W->P.x[q] *= dt[j] * dt[W->P.i[q]];
      for (int q = W->A.p[j]; q < W->A.p[j + 1]; ++q) W->A.x[q] *= dt[j] * et[W->A.i[q]];
      W->q[j] *= dt[j]; W->D[j] *= dt[j];
    }
    for (int i = 0; i < m; ++i) W->E[i] *= et[i];
    /* cost scaling: 1 / max(mean column norm of P, ||q||_inf) */
    double sum = 0.0;
    for (int j = 0; j < n; ++j) dt[j] = 0.0;
    for (int j = 0; j < n; ++j)
      for (int q = W->P.p[j]; q < W->P.p[j + 1]; ++q) {
        double a = fabs(W->P.x[q]); int i = W->P.i[q];
        if (a > dt[j]) dt[j] = a;
        if (a > dt[i]) dt[i] = a;
      }
    for (int j = 0; j < n; ++j) sum += dt[j];
    double cn = sum / n, qn = ninf(W->q, n);
    if (qn > cn) cn = qn;
    cn = limit_scaling(cn);
    double ct = 1.0 / cn;
    for (int q = 0; q < W->P.p[n]; ++q) W->P.x[q] *= ct;
    for (int j = 0; j < n; ++j) W->q[j] *= ct;
    W->c *= ct;
  }
  free(dt); free(et);
}

static void csc_copy(const csc* a, csc* b) {
  b->m = a->m; b->n = a->n;
  int nz = a->p[a->n];
  b->p = (int*)malloc(sizeof(int) * (a->n + 1)); memcpy(b->p, a->p, sizeof(int) * (a->n + 1));
  b->i = (int*)malloc(sizeof(int) * (nz > 0 ? nz : 1)); memcpy(b->i, a->i, sizeof(int) * nz);
  b->x = (double*)malloc(sizeof(double) * (nz > 0 ? nz : 1)); memcpy(b->x, a->x, sizeof(double) * nz);
}

/* KKT ordering: stage by stage, variables of stage k then the rows that belong to stage k (keeps the band narrow) */
static void kkt_ordering(int N, int n, int m, int* pv, int* pr) {
  int pos = 0;
  for (int i = 0; i < 4; ++i) pr[i] = -1;
  for (int k = 0; k <= N; ++k) {
    int nv = k < N ? 11 : 5;
    for (int i = 0; i < nv; ++i) pv[11 * k + i] = pos++;
    if (k == 0) for (int i = 0; i < 4; ++i) pr[i] = pos++;
    if (k < N) for (int i = 0; i < 19; ++i) pr[4 + 19 * k + i] = pos++;
    else for (int i = 0; i < 3; ++i) pr[4 + 19 * N + i] = pos++;
  }
  (void)n; (void)m;
}

/* K = [[P + reg I, A_sel'], [A_sel, -diag(w)]] restricted to rows with sel[i] (NULL = all), upper triangle, permuted */
static void kkt_build(work_t* W, double reg, const double* wdiag, const int* sel) {
  int n = W->n, m = W->m;
  trip t; trip_init(&t, n + m, n + m, W->P.p[n] + W->A.p[n] + n + m + 8);
  for (int j = 0; j < n; ++j) {
    int pj = W->perm_var[j];
    int has_diag = 0;
    for (int q = W->P.p[j]; q < W->P.p[j + 1]; ++q) {
      int pi = W->perm_var[W->P.i[q]];
      double v = W->P.x[q];
      if (W->P.i[q] == j) { v += reg; has_diag = 1; }
      if (pi <= pj) trip_add(&t, pi, pj, v); else trip_add(&t, pj, pi, v);
    }
    if (!has_diag) trip_add(&t, pj, pj, reg);
    for (int q = W->A.p[j]; q < W->A.p[j + 1]; ++q) {
      int r = W->A.i[q];
      if (sel && !sel[r]) continue;
      int pi = W->perm_row[r];
      if (pi <= pj) trip_add(&t, pi, pj, W->A.x[q]); else trip_add(&t, pj, pi, W->A.x[q]);
    }
  }
  for (int i = 0; i < m; ++i) {
    int pi = W->perm_row[i];
    if (sel && !sel[i]) trip_add(&t, pi, pi, 1.0);       /* deselected row: decoupled dummy unknown */
    else trip_add(&t, pi, pi, -wdiag[i]);
  }
  if (W->K.p) csc_free(&W->K);
  trip_to_csc(&t, &W->K);
  trip_free(&t);
}

static int kkt_factor(work_t* W) {
  if (W->symbolic_done) ldl_free(&W->F);
  ldl_symbolic(&W->F, &W->K);
  W->symbolic_done = 1;
  return ldl_numeric(&W->F, &W->K);
}

typedef struct { double pri, dua, eps_p, eps_d, sp, sd; } resid_t;

static void residuals(const work_t* W, const oracle_settings* s, const double* x, const double* z, const double* y,
                      double* t_m, double* t_m2, double* t_n, double* t_n2, resid_t* r) {
  int n = W->n, m = W->m;
  csc_mv(&W->A, x, t_m);                 /* Ax (scaled) */
  double nAx_s = ninf(t_m, m), nz_s = ninf(z, m);
  double pri_s = 0.0, pri = 0.0, nAx = 0.0, nz = 0.0;
  for (int i = 0; i < m; ++i) {
    double d = t_m[i] - z[i];
    if (fabs(d) > pri_s) pri_s = fabs(d);
    double du = W->Einv[i] * d; if (fabs(du) > pri) pri = fabs(du);
    double a = fabs(W->Einv[i] * t_m[i]); if (a > nAx) nAx = a;
    double b = fabs(W->Einv[i] * z[i]); if (b > nz) nz = b;
  }
  sym_mv(&W->P, x, t_n);                 /* Px */
  csc_mtv(&W->A, y, t_n2);               /* A'y */
  double dua_s = 0.0, dua = 0.0, nPx = 0.0, nAty = 0.0, nq = 0.0;
  double nPx_s = ninf(t_n, n), nAty_s = ninf(t_n2, n), nq_s = ninf(W->q, n);
  for (int j = 0; j < n; ++j) {
    double d = t_n[j] + W->q[j] + t_n2[j];
    if (fabs(d) > dua_s) dua_s = fabs(d);
    double du = W->Dinv[j] * d; if (fabs(du) > dua) dua = fabs(du);
    double a = fabs(W->Dinv[j] * t_n[j]); if (a > nPx) nPx = a;
    double b = fabs(W->Dinv[j] * t_n2[j]); if (b > nAty) nAty = b;
    double c = fabs(W->Dinv[j] * W->q[j]); if (c > nq) nq = c;
  }
  (void)t_m2;
  r->pri = pri; r->dua = W->cinv * dua;
  r->eps_p = s->eps_abs + s->eps_rel * fmax(nAx, nz);
  r->eps_d = s->eps_abs + s->eps_rel * W->cinv * fmax(fmax(nPx, nAty), nq);
  r->sp = pri_s / (fmax(nAx_s, nz_s) + 1e-10);
  r->sd = dua_s / (fmax(fmax(nPx_s, nAty_s), nq_s) + 1e-10);
}

static double clip(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* solves the QP; outputs unscaled x (n), y (m), z (m); returns status */
static int osqp_solve(const qp_t* Q, const oracle_settings* s, double* xo, double* yo, double* zo, int* iters_out,
                      double* pri_out, double* dua_out, int* info) {
  int n = Q->n, m = Q->m, N = Q->N;
  work_t W; memset(&W, 0, sizeof W);
  W.n = n; W.m = m;
  csc_copy(&Q->P, &W.P); csc_copy(&Q->A, &W.A);
  W.q = (double*)malloc(sizeof(double) * n); memcpy(W.q, Q->q, sizeof(double) * n);
  W.l = (double*)malloc(sizeof(double) * m); W.u = (double*)malloc(sizeof(double) * m);
  W.D = (double*)malloc(sizeof(double) * n); W.E = (double*)malloc(sizeof(double) * m);
  W.Dinv = (double*)malloc(sizeof(double) * n); W.Einv = (double*)malloc(sizeof(double) * m);
  for (int j = 0; j < n; ++j) W.D[j] = 1.0;
  for (int i = 0; i < m; ++i) W.E[i] = 1.0;
  W.c = 1.0;
  if (s->scaling > 0) scale_problem(&W, s->scaling);
  for (int j = 0; j < n; ++j) W.Dinv[j] = 1.0 / W.D[j];
  for (int i = 0; i < m; ++i) {
    W.Einv[i] = 1.0 / W.E[i];
    W.l[i] = Q->l[i] <= -OSQP_INFTY ? -OSQP_INFTY : W.E[i] * Q->l[i];
    W.u[i] = Q->u[i] >= OSQP_INFTY ? OSQP_INFTY : W.E[i] * Q->u[i];
  }
  W.cinv = 1.0 / W.c;
  W.perm_var = (int*)malloc(sizeof(int) * n); W.perm_row = (int*)malloc(sizeof(int) * m);
  kkt_ordering(N, n, m, W.perm_var, W.perm_row);

  double* rhov = (double*)malloc(sizeof(double) * m); double* rinv = (double*)malloc(sizeof(double) * m);
  int* is_eq = (int*)malloc(sizeof(int) * m);
  for (int i = 0; i < m; ++i) is_eq[i] = (W.u[i] - W.l[i]) < RHO_TOL;
  double rho = s->rho;
#define SET_RHO()                                                                                      \
  for (int i = 0; i < m; ++i) {                                                                       \
    int fr = (W.l[i] <= -OSQP_INFTY) && (W.u[i] >= OSQP_INFTY);                                        \
    rhov[i] = fr ? s->rho_min : (is_eq[i] ? s->rho_eq_factor * rho : rho);                             \
    rinv[i] = 1.0 / rhov[i];                                                                           \
  }
  SET_RHO();
  kkt_build(&W, s->sigma, rinv, NULL);
  int n_fac = 0, n_rho = 0, n_pol = 0;
  kkt_factor(&W); ++n_fac;

  double* x = (double*)calloc(n, sizeof(double)); double* y = (double*)calloc(m, sizeof(double)); double* z = (double*)calloc(m, sizeof(double));
  double* rhs = (double*)malloc(sizeof(double) * (n + m));
  double* xt = (double*)malloc(sizeof(double) * n); double* zt = (double*)malloc(sizeof(double) * m);
  double* t_m = (double*)malloc(sizeof(double) * m); double* t_m2 = (double*)malloc(sizeof(double) * m);
  double* t_n = (double*)malloc(sizeof(double) * n); double* t_n2 = (double*)malloc(sizeof(double) * n);
  if (s->z0_projected) for (int i = 0; i < m; ++i) z[i] = clip(0.0, W.l[i], W.u[i]);

  int status = ST_UNSOLVED, it = 0;
  resid_t r; memset(&r, 0, sizeof r); r.pri = r.dua = 1e300;
  while (it < s->max_iter) {
    ++it;
    /* [[P+sigma I, A'],[A, -1/rho]] [xt; nu] = [sigma x - q; z - y/rho] */
    for (int j = 0; j < n; ++j) rhs[W.perm_var[j]] = s->sigma * x[j] - W.q[j];
    for (int i = 0; i < m; ++i) rhs[W.perm_row[i]] = z[i] - rinv[i] * y[i];
    ldl_solve(&W.F, rhs);
    for (int j = 0; j < n; ++j) xt[j] = rhs[W.perm_var[j]];
    for (int i = 0; i < m; ++i) zt[i] = z[i] + rinv[i] * (rhs[W.perm_row[i]] - y[i]);
    for (int j = 0; j < n; ++j) x[j] = s->alpha * xt[j] + (1.0 - s->alpha) * x[j];
    for (int i = 0; i < m; ++i) {
      double w = s->alpha * zt[i] + (1.0 - s->alpha) * z[i];
      double zn = clip(w + rinv[i] * y[i], W.l[i], W.u[i]);
      y[i] += rhov[i] * (w - zn);
      z[i] = zn;
    }
    int check = s->check_termination > 0 && (it % s->check_termination == 0);
    int adapt = s->adaptive_rho && s->adaptive_rho_interval > 0 && (it % s->adaptive_rho_interval == 0);
    if (check || adapt) {
      residuals(&W, s, x, z, y, t_m, t_m2, t_n, t_n2, &r);
      if (check && r.pri <= r.eps_p && r.dua <= r.eps_d) { status = ST_SOLVED; break; }
      if (adapt) {
        double rn = rho * sqrt(r.sp / (r.sd + 1e-10));
        rn = fmin(fmax(rn, s->rho_min), s->rho_max);
        if (rn > rho * s->adaptive_rho_tolerance || rn < rho / s->adaptive_rho_tolerance) {
          rho = rn; ++n_rho;
          SET_RHO();
          kkt_build(&W, s->sigma, rinv, NULL);
          kkt_factor(&W); ++n_fac;
        }
      }
    }
  }
  if (status != ST_SOLVED) {
    residuals(&W, s, x, z, y, t_m, t_m2, t_n, t_n2, &r);
    if (r.pri <= r.eps_p && r.dua <= r.eps_d) status = ST_SOLVED;
    else {
      double ep10 = 10.0 * s->eps_abs + 10.0 * (r.eps_p - s->eps_abs), ed10 = 10.0 * s->eps_abs + 10.0 * (r.eps_d - s->eps_abs);
      status = (r.pri <= ep10 && r.dua <= ed10) ? ST_INACCURATE : ST_MAX_ITER;
    }
  }
  double pri = r.pri, dua = r.dua;

  /* polish: pass 1 is OSQP's; further passes re-identify the active set from Ax + y (primal-dual active set) */
  if (status == ST_SOLVED && s->polish_passes > 0) {
    int* low = (int*)malloc(sizeof(int) * m); int* upp = (int*)malloc(sizeof(int) * m); int* sel = (int*)malloc(sizeof(int) * m);
    int* plow = (int*)calloc(m, sizeof(int)); int* pupp = (int*)calloc(m, sizeof(int));
    double* dd = (double*)malloc(sizeof(double) * m);
    double* sol = (double*)malloc(sizeof(double) * (n + m)); double* res = (double*)malloc(sizeof(double) * (n + m));
    double* xp = (double*)malloc(sizeof(double) * n); double* yp = (double*)malloc(sizeof(double) * m); double* zp = (double*)malloc(sizeof(double) * m);
    double* zref = (double*)malloc(sizeof(double) * m);
    memcpy(zref, z, sizeof(double) * m);
    for (int pass = 0; pass < s->polish_passes; ++pass) {
      int changed = 0;
      for (int i = 0; i < m; ++i) {
        double ml = -y[i] - (zref[i] - W.l[i]), mu = y[i] - (W.u[i] - zref[i]);   /* active iff margin > 0 */
        int lo = ml > 0.0, up = mu > 0.0;
        if (pass > 0) { if (fabs(ml) <= 1e-7) lo = plow[i]; if (fabs(mu) <= 1e-7) up = pupp[i]; }
        low[i] = lo; upp[i] = up; sel[i] = lo | up;
        if (pass == 0 || lo != plow[i] || up != pupp[i]) changed = 1;
        dd[i] = s->delta;
      }
      if (pass > 0 && !changed) break;
      memcpy(plow, low, sizeof(int) * m); memcpy(pupp, upp, sizeof(int) * m);
      kkt_build(&W, s->delta, dd, sel);
      if (kkt_factor(&W) != 0) break;
      ++n_fac;
      /* rhs = [-q; b_act]; sol = Khat^{-1} rhs; refine against the un-regularised K */
      for (int j = 0; j < n; ++j) xp[j] = 0.0;
      for (int i = 0; i < m; ++i) yp[i] = 0.0;
      for (int step = 0; step <= s->polish_refine_iter; ++step) {
        sym_mv(&W.P, xp, t_n); csc_mtv(&W.A, yp, t_n2); csc_mv(&W.A, xp, t_m);
        for (int j = 0; j < n; ++j) res[W.perm_var[j]] = -W.q[j] - t_n[j] - t_n2[j];
        for (int i = 0; i < m; ++i) res[W.perm_row[i]] = sel[i] ? ((low[i] ? W.l[i] : W.u[i]) - t_m[i]) : 0.0;
        ldl_solve(&W.F, res);
        for (int j = 0; j < n; ++j) xp[j] += res[W.perm_var[j]];
        for (int i = 0; i < m; ++i) if (sel[i]) yp[i] += res[W.perm_row[i]];
      }
      csc_mv(&W.A, xp, zref);            /* unprojected Ax: used for the next activity test */
      for (int i = 0; i < m; ++i) zp[i] = clip(zref[i], W.l[i], W.u[i]);
      double pri_p = 0.0;
      for (int i = 0; i < m; ++i) { double d = fabs(W.Einv[i] * (zref[i] - zp[i])); if (d > pri_p) pri_p = d; }
      sym_mv(&W.P, xp, t_n); csc_mtv(&W.A, yp, t_n2);
      double dua_p = 0.0;
      for (int j = 0; j < n; ++j) { double d = fabs(W.Dinv[j] * (t_n[j] + W.q[j] + t_n2[j])); if (d > dua_p) dua_p = d; }
      dua_p *= W.cinv;
      int ok;
      if (pass == 0) ok = (pri_p < pri && dua_p < dua) || (pri_p < pri && dua < 1e-10) || (dua_p < dua && pri < 1e-10);
      else ok = pri_p <= fmax(10.0 * pri, 1e-9) && dua_p <= fmax(10.0 * dua, 1e-9);
      if (!ok) break;
      memcpy(x, xp, sizeof(double) * n); memcpy(y, yp, sizeof(double) * m); memcpy(z, zp, sizeof(double) * m);
      pri = pri_p; dua = dua_p; n_pol = pass + 1;
    }
    free(low); free(upp); free(sel); free(plow); free(pupp); free(dd); free(sol); free(res); free(xp); free(yp); free(zp); free(zref);
  }

  for (int j = 0; j < n; ++j) xo[j] = W.D[j] * x[j];
  if (yo) for (int i = 0; i < m; ++i) yo[i] = W.cinv * W.E[i] * y[i];
  if (zo) for (int i = 0; i < m; ++i) zo[i] = W.Einv[i] * z[i];
  *iters_out = it; *pri_out = pri; *dua_out = dua;
  if (info) { info[0] = n_rho; info[1] = n_fac; info[2] = n_pol; info[3] = it; }

  free(x); free(y); free(z); free(rhs); free(xt); free(zt); free(t_m); free(t_m2); free(t_n); free(t_n2);
  free(rhov); free(rinv); free(is_eq);
  csc_free(&W.P); csc_free(&W.A); if (W.K.p) csc_free(&W.K);
  if (W.symbolic_done) ldl_free(&W.F);
  free(W.q); free(W.l); free(W.u); free(W.D); free(W.E); free(W.Dinv); free(W.Einv); free(W.perm_var); free(W.perm_row);
  return status;
}

/* ------------------------------------------------------------------------------------------------ */
/* public entry points (ctypes)                                                                     */
/* ------------------------------------------------------------------------------------------------ */
void oracle_default_settings(oracle_settings* s) {
  memset(s, 0, sizeof *s);
  s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->rho = 0.1; s->alpha = 1.6;   /* mpc_controller.py:121-131 */
  s->sigma = 1e-6; s->adaptive_rho_tolerance = 5.0; s->rho_eq_factor = 1e3; s->rho_min = 1e-6; s->rho_max = 1e6; s->delta = 1e-6;
  s->max_iter = 60000; s->check_termination = 25; s->adaptive_rho = 1; s->adaptive_rho_interval = 50;
  s->polish_passes = 1; s->polish_refine_iter = 3; s->scaling = 10; s->z0_projected = 0;
}

/* MPCController.solve (mpc_controller.py:39-141): status in {1,2} <=> reference returns arrays, else (None,)*3 */
int oracle_solve(const oracle_params* p, const oracle_settings* s, const double* x0, const double* ref, const double* u_prev,
                 double* u0, double* Xp, double* Up, int* iters, double* pri, double* dua, int* info) {
  qp_t Q; memset(&Q, 0, sizeof Q);
  qp_build(p, x0, ref, u_prev, &Q);
  int N = Q.N;
  double* z = (double*)malloc(sizeof(double) * Q.n);
  int status = osqp_solve(&Q, s, z, NULL, NULL, iters, pri, dua, info);
  for (int k = 0; k <= N; ++k) for (int i = 0; i < 4; ++i) Xp[i * (N + 1) + k] = z[ix(k, i)];
  for (int k = 0; k < N; ++k) for (int i = 0; i < 2; ++i) Up[i * N + k] = z[iu(k, i)];
  u0[0] = z[iu(0, 0)]; u0[1] = z[iu(0, 1)];
  free(z);
  qp_free(&Q);
  return status;
}

int oracle_solve_batch(const oracle_params* p, const oracle_settings* s, int B, const double* x0, const double* ref,
                       const double* u_prev, double* u0, double* Xp, double* Up, int* status, int* iters, double* pri,
                       double* dua, int* info) {
  int N = p->horizon;
  for (int b = 0; b < B; ++b)
    status[b] = oracle_solve(p, s, x0 + 4 * (size_t)b, ref + (size_t)4 * (N + 1) * b, u_prev ? u_prev + 2 * (size_t)b : NULL,
                             u0 + 2 * (size_t)b, Xp + (size_t)4 * (N + 1) * b, Up + (size_t)2 * N * b, iters + b, pri + b,
                             dua + b, info ? info + 4 * (size_t)b : NULL);
  return 0;
}

/* QP matrices for cross-checks: dense copies (row-major) */
int oracle_qp_dense(const oracle_params* p, const double* x0, const double* ref, const double* u_prev, double* Pd,
                    double* q, double* Ad, double* l, double* u) {
  qp_t Q; memset(&Q, 0, sizeof Q);
  qp_build(p, x0, ref, u_prev, &Q);
  int n = Q.n, m = Q.m;
  memset(Pd, 0, sizeof(double) * n * n); memset(Ad, 0, sizeof(double) * m * n);
  for (int j = 0; j < n; ++j) {
    for (int t = Q.P.p[j]; t < Q.P.p[j + 1]; ++t) { Pd[(size_t)Q.P.i[t] * n + j] = Q.P.x[t]; Pd[(size_t)j * n + Q.P.i[t]] = Q.P.x[t]; }
    for (int t = Q.A.p[j]; t < Q.A.p[j + 1]; ++t) Ad[(size_t)Q.A.i[t] * n + j] = Q.A.x[t];
  }
  memcpy(q, Q.q, sizeof(double) * n); memcpy(l, Q.l, sizeof(double) * m); memcpy(u, Q.u, sizeof(double) * m);
  qp_free(&Q);
  return 0;
}

/* TrajectoryTracker.track for one vehicle (control_stage.py:74-157); returns number of states written.
 * flags: bit0 goal reached, bit1 aborted, bit2 the relaxation retry (control_stage.py:45-56) ran at least once.  ref_global (len,4) from build_reference. */
int oracle_track(const oracle_params* p, const oracle_settings* s, const double* ref_global, int len, const double* state0,
                 const double* goal, int sim_steps, double* states, double* controls, int* step_status, int* step_iters, int* flags) {
  int N = p->horizon;
  double state[4], u_prev[2] = {0.0, 0.0};
  memcpy(state, state0, sizeof state);
  double* win = (double*)malloc(sizeof(double) * 4 * (N + 1));
  double* Xp = (double*)malloc(sizeof(double) * 4 * (N + 1)); double* Up = (double*)malloc(sizeof(double) * 2 * N);
  int path_idx = 0, nst = 0; *flags = 0;
  for (int step = 0; step < sim_steps; ++step) {
    for (int k = 0; k <= N; ++k) { int r = path_idx + k; if (r > len - 1) r = len - 1; memcpy(win + 4 * k, ref_global + 4 * (size_t)r, 4 * sizeof(double)); }
    double u0[2], pri, dua; int it;
    int st = oracle_solve(p, s, state, win, u_prev, u0, Xp, Up, &it, &pri, &dua, NULL);
    if (st != ST_SOLVED && st != ST_INACCURATE) {      /* control_stage.py:45-56 */
      oracle_params pr = *p;
      pr.du_bounds[0] -= 5.0; pr.du_bounds[1] += 5.0; pr.du_bounds[2] -= 0.05; pr.du_bounds[3] += 0.05;
      for (int k = 0; k <= N; ++k) win[4 * k + 3] *= 0.6;
      st = oracle_solve(&pr, s, state, win, u_prev, u0, Xp, Up, &it, &pri, &dua, NULL);
      *flags |= 4;                                       /* the relaxation retry ran at least once */
    }
    if (step_status) step_status[step] = st;
    if (step_iters) step_iters[step] = it;
    if (st != ST_SOLVED && st != ST_INACCURATE) { *flags |= 2; break; }
    double nx[4];
    oracle_f_discrete(state, u0, p->dt, p->wheelbase_px, nx);
    memcpy(state, nx, sizeof state);
    memcpy(states + 4 * (size_t)step, state, sizeof state);
    if (controls) { controls[2 * step] = u0[0]; controls[2 * step + 1] = u0[1]; }
    u_prev[0] = u0[0]; u_prev[1] = u0[1];
    nst = step + 1;
    if (path_idx < len - 2) {
      double dx = state[0] - ref_global[4 * (size_t)path_idx], dy = state[1] - ref_global[4 * (size_t)path_idx + 1];
      if (dx * dx + dy * dy > 25.0) path_idx += 1;
    }
    if (hypot(state[0] - goal[0], state[1] - goal[1]) < 8.0) { *flags |= 1; break; }
  }
  free(win); free(Xp); free(Up);
  return nst;
}
