/* gko.c -- CPU oracle for the gokalman hot path (see gko.h).  TEST INFRASTRUCTURE ONLY.
 *
 * Every function cites the reference lines it restates.  Products are written in the order the
 * reference writes them ((F*P)*F^T, not F*(P*F^T)); inverses are explicit LU inverses followed by a
 * multiply, exactly where the reference calls mat64.Dense.Inverse; mat64.SymDense values are
 * modelled by mirroring the upper triangle (SymDense only reads the upper triangle).
 */
#include "gko.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gko_linalg.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define GKO_MAXC 16

struct gko_filter {
  int kind;
  int n, m, c, q;
  int need_ctrl;       /* !IsNil(G): vanilla.go:39, information.go:52, squareroot.go:48 */
  int prediction_only; /* vanilla.go:61 */
  int step;
  /* model */
  double *F, *G, *H, *Q, *R; /* R is m_r x m_r */
  int m_r;
  /* information.go:38-50: precomputed once, NOT refreshed by SetNoise (information.go:136-138) */
  double *Finv, *Qinv, *Rinv;
  int rinv_dim;
  /* squareroot.go:100-114: lower Cholesky factors, refreshed by every SetNoise */
  double *sqrtQ, *sqrtR;
  /* previous / initial estimate.  Meaning by kind:
   *   vanilla, hybrid : x = state, A = covariance (sym)
   *   information     : x = info state i, A = info matrix I (sym)
   *   sqrt            : x = state, A = stddev S
   *   srif            : x = b, A = R                                           */
  double *x, *A, *x_init, *A_init;
  /* replay noise (BatchNoise-style index by k, noise.go:73-86) */
  int replay_steps, replay_mv;
  double *replay_w, *replay_v;
  double* replay_w2; /* optional: the vector the SECOND Process(k) call of Vanilla.Update returns (vanilla.go:195) --
                        AWGN draws a fresh sample on every call (noise.go:127-131), BatchNoise repeats vector k */
  /* NLDKF */
  double *Phi, *Htilde, *Gamma;
  int has_htilde, ekf, locked, snc;
  double* sqrt_inv_noise; /* srif.go:48: stores L = chol(R_meas), not its inverse */
  int non_tri_r;
  /* scratch */
  double* ws;
  size_t ws_len;
};

static double* dalloc(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }
static void dcopy(double* d, const double* s, size_t n) { memcpy(d, s, n * sizeof(double)); }

static int is_nil(const double* M, int len) { /* helper.go:50-63 */
  if (!M) return 1;
  for (int i = 0; i < len; ++i)
    if (M[i] != 0.0) return 0;
  return 1;
}

static void sym_from_upper(double* dst, const double* src, int n) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) dst[i * n + j] = (j >= i) ? src[i * n + j] : src[j * n + i];
}

static gko_filter* base_new(int kind, int n, int m, int c, int q) {
  if (n < 1 || n > GKO_MAXN || m < 0 || m > GKO_MAXM || c < 0 || c > GKO_MAXC) return NULL;
  gko_filter* f = (gko_filter*)calloc(1, sizeof(gko_filter));
  f->kind = kind;
  f->n = n;
  f->m = m;
  f->c = c;
  f->q = q;
  size_t nn = (size_t)n * n;
  f->F = dalloc(nn);
  f->G = dalloc((size_t)n * GKO_MAXC);
  f->H = dalloc((size_t)GKO_MAXM * n);
  f->Q = dalloc(nn);
  f->R = dalloc((size_t)GKO_MAXM * GKO_MAXM);
  f->Finv = dalloc(nn);
  f->Qinv = dalloc(nn);
  f->Rinv = dalloc((size_t)GKO_MAXM * GKO_MAXM);
  f->sqrtQ = dalloc(nn);
  f->sqrtR = dalloc((size_t)GKO_MAXM * GKO_MAXM);
  f->x = dalloc(n);
  f->A = dalloc(nn);
  f->x_init = dalloc(n);
  f->A_init = dalloc(nn);
  f->Phi = dalloc(nn);
  f->Htilde = dalloc((size_t)GKO_MAXM * n);
  f->Gamma = dalloc(nn);
  f->sqrt_inv_noise = dalloc((size_t)GKO_MAXM * GKO_MAXM);
  int d = 2 * n + GKO_MAXM + 2;
  f->ws_len = (size_t)24 * d * d;
  f->ws = dalloc(f->ws_len);
  f->m_r = m;
  return f;
}

void gko_free(gko_filter* f) {
  if (!f) return;
  free(f->F); free(f->G); free(f->H); free(f->Q); free(f->R);
  free(f->Finv); free(f->Qinv); free(f->Rinv); free(f->sqrtQ); free(f->sqrtR);
  free(f->x); free(f->A); free(f->x_init); free(f->A_init);
  free(f->replay_w); free(f->replay_v); free(f->replay_w2);
  free(f->Phi); free(f->Htilde); free(f->Gamma); free(f->sqrt_inv_noise);
  free(f->ws);
  free(f);
}

static void set_lti_model(gko_filter* f, const double* F, const double* G, const double* H,
                          const double* Q, const double* R) {
  int n = f->n, m = f->m, c = f->c;
  dcopy(f->F, F, (size_t)n * n);
  if (G && c > 0) dcopy(f->G, G, (size_t)n * c);
  f->need_ctrl = !(G == NULL || c == 0 || is_nil(G, n * c));
  dcopy(f->H, H, (size_t)m * n);
  sym_from_upper(f->Q, Q, n);
  sym_from_upper(f->R, R, m);
  f->m_r = m;
}

/* ---- noise.go ------------------------------------------------------------------------------- */

/* Noise.Process(k): Noiseless -> zeros (noise.go:36-38); replay -> stored vector k, out of range is
 * the reference's panic (noise.go:73-78). Returns 0 or GKO_ERR_NOISE_RANGE. */
static int noise_process(const gko_filter* f, int k, double* w) {
  int n = f->n;
  if (f->replay_w) {
    if (k >= f->replay_steps) return GKO_ERR_NOISE_RANGE;
    dcopy(w, f->replay_w + (size_t)k * n, n);
  } else {
    for (int i = 0; i < n; ++i) w[i] = 0.0;
  }
  return 0;
}

static int noise_process_again(const gko_filter* f, int k, double* w) {
  if (f->replay_w2) {
    if (k >= f->replay_steps) return GKO_ERR_NOISE_RANGE;
    dcopy(w, f->replay_w2 + (size_t)k * f->n, f->n);
    return 0;
  }
  return noise_process(f, k, w);
}

static int noise_measurement(const gko_filter* f, int k, double* v) {
  int m = f->m;
  if (f->replay_v) {
    if (k >= f->replay_steps) return GKO_ERR_NOISE_RANGE;
    for (int i = 0; i < m; ++i) v[i] = (i < f->replay_mv) ? f->replay_v[(size_t)k * f->replay_mv + i] : 0.0;
  } else {
    for (int i = 0; i < m; ++i) v[i] = 0.0;
  }
  return 0;
}

void gko_set_replay_second_draw(gko_filter* f, const double* w2) { /* after gko_set_replay: [replay_steps][n] */
  free(f->replay_w2);
  f->replay_w2 = NULL;
  if (w2) {
    f->replay_w2 = dalloc((size_t)f->replay_steps * f->n);
    dcopy(f->replay_w2, w2, (size_t)f->replay_steps * f->n);
  }
}

void gko_set_replay(gko_filter* f, int steps, const double* w, const double* v, int m_v) {
  free(f->replay_w);
  free(f->replay_v);
  free(f->replay_w2);
  f->replay_w2 = NULL;
  f->replay_w = f->replay_v = NULL;
  f->replay_steps = steps;
  f->replay_mv = m_v;
  if (w) {
    f->replay_w = dalloc((size_t)steps * f->n);
    dcopy(f->replay_w, w, (size_t)steps * f->n);
  }
  if (v) {
    f->replay_v = dalloc((size_t)steps * m_v);
    dcopy(f->replay_v, v, (size_t)steps * m_v);
  }
}

/* ---- constructors ----------------------------------------------------------------------------- */

gko_filter* gko_new_vanilla(int n, int m, int c, const double* x0, const double* P0, const double* F,
                            const double* G, const double* H, const double* Q, const double* R,
                            int pure_predictor) { /* vanilla.go:21-62 */
  gko_filter* f = base_new(pure_predictor ? GKO_PREDICTOR : GKO_VANILLA, n, m, c, 0);
  if (!f) return NULL;
  set_lti_model(f, F, G, H, Q, R);
  f->prediction_only = pure_predictor;
  dcopy(f->x, x0, n);
  sym_from_upper(f->A, P0, n);
  dcopy(f->x_init, f->x, n);
  dcopy(f->A_init, f->A, (size_t)n * n);
  return f;
}

gko_filter* gko_new_information(int n, int m, int c, const double* i0, const double* I0, const double* F,
                                const double* G, const double* H, const double* Q, const double* R) {
  /* information.go:20-53 */
  gko_filter* f = base_new(GKO_INFORMATION, n, m, c, 0);
  if (!f) return NULL;
  set_lti_model(f, F, G, H, Q, R);
  dcopy(f->x, i0, n);
  sym_from_upper(f->A, I0, n);
  dcopy(f->x_init, f->x, n);
  dcopy(f->A_init, f->A, (size_t)n * n);
  /* errors are only printed by the reference (information.go:39-50) */
  gko_inverse(f->Finv, f->F, n, NULL);
  gko_inverse(f->Qinv, f->Q, n, NULL);
  gko_inverse(f->Rinv, f->R, m, NULL);
  f->rinv_dim = m;
  return f;
}

gko_filter* gko_new_information_from_state(int n, int m, int c, const double* x0, const double* P0,
                                           const double* F, const double* G, const double* H,
                                           const double* Q, const double* R) { /* information.go:65-81 */
  double* P = dalloc((size_t)n * n);
  double* I0 = dalloc((size_t)n * n);
  double* i0 = dalloc(n);
  sym_from_upper(P, P0, n);
  if (gko_inverse(I0, P, n, NULL) != 0) {
    memset(I0, 0, sizeof(double) * (size_t)n * n); /* information.go:69-72 */
  } else {
    gko_as_sym(I0, n); /* error ignored: information.go:74 */
  }
  gko_mulvec(i0, I0, x0, n, n);
  gko_filter* f = gko_new_information(n, m, c, i0, I0, F, G, H, Q, R);
  free(P);
  free(I0);
  free(i0);
  return f;
}

static void sqrt_refresh_noise(gko_filter* f) { /* squareroot.go:100-114 */
  gko_chol_lower(f->sqrtQ, f->Q, f->n);
  gko_chol_lower(f->sqrtR, f->R, f->m_r);
}

gko_filter* gko_new_sqrt(int n, int m, int c, const double* x0, const double* P0, const double* F,
                         const double* G, const double* H, const double* Q, const double* R) {
  /* squareroot.go:21-50 */
  gko_filter* f = base_new(GKO_SQRT, n, m, c, 0);
  if (!f) return NULL;
  set_lti_model(f, F, G, H, Q, R);
  dcopy(f->x, x0, n);
  double* P = dalloc((size_t)n * n);
  sym_from_upper(P, P0, n);
  gko_chol_lower(f->A, P, n); /* stddev = chol_lower(P0), ok flag ignored (squareroot.go:35-40) */
  free(P);
  dcopy(f->x_init, f->x, n);
  dcopy(f->A_init, f->A, (size_t)n * n);
  sqrt_refresh_noise(f);
  return f;
}

gko_filter* gko_new_hybrid(int n, int m, int q, const double* x0, const double* P0, const double* Q,
                           const double* R) { /* hybrid.go:23-34 */
  gko_filter* f = base_new(GKO_HYBRID, n, m, 0, q);
  if (!f) return NULL;
  dcopy(f->x, x0, n);
  sym_from_upper(f->A, P0, n);
  dcopy(f->x_init, f->x, n);
  dcopy(f->A_init, f->A, (size_t)n * n);
  if (Q && q > 0) sym_from_upper(f->Q, Q, q); /* Q is q x q (hybrid.go:118-121: Gamma*Q*Gamma^T) */
  sym_from_upper(f->R, R, m);
  f->locked = 1;
  return f;
}

gko_filter* gko_new_srif(int n, int m, const double* x0, const double* P0, const double* R, int non_tri_r) {
  /* srif.go:14-49 */
  gko_filter* f = base_new(GKO_SRIF, n, m, 0, 0);
  if (!f) return NULL;
  double* I0 = dalloc((size_t)n * n);
  for (int i = 0; i < n; ++i) I0[i * n + i] = 1.0 / P0[i * n + i]; /* srif.go:23-26: P0 assumed diagonal */
  gko_chol_lower(f->A, I0, n);                                      /* R0 (lower) */
  gko_mulvec(f->x, f->A, x0, n, n);                                 /* b0 = R0 x0 */
  free(I0);
  sym_from_upper(f->R, R, m);
  gko_chol_lower(f->sqrt_inv_noise, f->R, m); /* L; srif.go:43-48 inverts L then keeps L */
  {
    double* Linv = dalloc((size_t)m * m);
    int ierr = gko_inverse(Linv, f->sqrt_inv_noise, m, NULL);
    free(Linv);
    if (ierr != 0) { /* srif.go:43-45 returns the error */
      gko_free(f);
      return NULL;
    }
  }
  dcopy(f->x_init, f->x, n);
  dcopy(f->A_init, f->A, (size_t)n * n);
  f->non_tri_r = non_tri_r;
  f->locked = 1;
  return f;
}

/* ---- setters ---------------------------------------------------------------------------------- */

void gko_set_state_transition(gko_filter* f, const double* F) {
  dcopy(f->F, F, (size_t)f->n * f->n);
  if (f->kind == GKO_INFORMATION) gko_inverse(f->Finv, f->F, f->n, NULL); /* information.go:117-123 */
}

void gko_set_input_control(gko_filter* f, int c, const double* G) {
  /* vanilla.go:99-101 etc.: needCtrl is NOT re-evaluated by the setter */
  f->c = c;
  if (G && c > 0) dcopy(f->G, G, (size_t)f->n * c);
}

void gko_set_measurement_matrix(gko_filter* f, int m, const double* H) {
  f->m = m;
  dcopy(f->H, H, (size_t)m * f->n);
}

void gko_set_noise(gko_filter* f, const double* Q, int m_r, const double* R) {
  int qd = (f->kind == GKO_HYBRID) ? f->q : f->n;
  if (Q && qd > 0) sym_from_upper(f->Q, Q, qd);
  sym_from_upper(f->R, R, m_r);
  f->m_r = m_r;
  if (f->kind == GKO_SQRT) sqrt_refresh_noise(f);
  /* GKO_INFORMATION: Qinv/Rinv deliberately left stale (information.go:136-138) */
  free(f->replay_w);
  free(f->replay_v);
  free(f->replay_w2);
  f->replay_w = f->replay_v = f->replay_w2 = NULL;
}

void gko_reset(gko_filter* f) { /* vanilla.go:121-125, information.go:146-150, squareroot.go:122-126 */
  dcopy(f->x, f->x_init, f->n);
  dcopy(f->A, f->A_init, (size_t)f->n * f->n);
  f->step = 0;
}

/* ---- lazy estimate accessors ------------------------------------------------------------------ */

/* information.go:276-293: Covariance() = AsSymDense(inv(I)) or zeros when Inverse errors. */
static int info_covariance(const double* Imat, int n, double* P) {
  int ierr = gko_inverse(P, Imat, n, NULL);
  if (ierr != 0) {
    memset(P, 0, sizeof(double) * (size_t)n * n);
    return 0;
  }
  if (gko_as_sym(P, n) != 0) { /* 290: error dropped, the nil SymDense would crash; zeros here */
    memset(P, 0, sizeof(double) * (size_t)n * n);
    return 0;
  }
  return 1;
}

/* srif.go:223-235 State() = inv(R) b; returns GKO_ERR_SINGULAR_R where the reference panics. */
static int srif_state(const double* Rm, const double* b, int n, double* x, double* ws) {
  double* Rinv = ws;
  if (gko_inverse(Rinv, Rm, n, NULL) != 0) return GKO_ERR_SINGULAR_R;
  gko_mulvec(x, Rinv, b, n, n);
  return 0;
}

/* srif.go:253-265: Covariance() = AsSymDense(inv(R) inv(R)^T), nil (zeros here) if not invertible */
static int srif_covariance(const double* Rm, int n, double* P, double* ws) {
  double* Rinv = ws;
  if (gko_inverse(Rinv, Rm, n, NULL) != 0) {
    memset(P, 0, sizeof(double) * (size_t)n * n);
    return 0;
  }
  gko_mul_nt(P, Rinv, Rinv, n, n, n);
  gko_as_sym(P, n);
  return 1;
}

static void est_clear(gko_estimate* e, int n, int m) {
  e->n = n;
  e->m = m;
  e->innov_len = m;
  e->covar_ok = e->pred_covar_ok = 1;
  memset(e->state, 0, sizeof(double) * n);
  memset(e->meas, 0, sizeof(e->meas));
  memset(e->innov, 0, sizeof(e->innov));
  memset(e->obs_dev, 0, sizeof(e->obs_dev));
  memset(e->covar, 0, sizeof(double) * (size_t)n * n);
  memset(e->pred_covar, 0, sizeof(double) * (size_t)n * n);
  memset(e->gain, 0, sizeof(double) * (size_t)n * (m > 0 ? m : 1));
  memset(e->raw_vec, 0, sizeof(double) * n);
  memset(e->raw_mat, 0, sizeof(double) * (size_t)n * n);
  memset(e->raw_pred_mat, 0, sizeof(double) * (size_t)n * n);
}

void gko_initial_estimate(const gko_filter* f, gko_estimate* e) {
  int n = f->n;
  est_clear(e, n, f->m);
  switch (f->kind) {
    case GKO_VANILLA:
    case GKO_PREDICTOR:
    case GKO_HYBRID: /* est0 = {x0, 0, 0, P0, 0} (vanilla.go:34-37, hybrid.go:30-32) */
      dcopy(e->state, f->x_init, n);
      dcopy(e->covar, f->A_init, (size_t)n * n);
      break;
    case GKO_INFORMATION:
      dcopy(e->raw_vec, f->x_init, n);
      dcopy(e->raw_mat, f->A_init, (size_t)n * n);
      e->covar_ok = info_covariance(f->A_init, n, e->covar);
      gko_mulvec(e->state, e->covar, f->x_init, n, n);
      e->pred_covar_ok = 0; /* inverse of the zero predicted information matrix */
      dcopy(e->innov, f->x_init, n);
      e->innov_len = n;
      break;
    case GKO_SQRT:
      dcopy(e->state, f->x_init, n);
      dcopy(e->raw_mat, f->A_init, (size_t)n * n);
      gko_mul_nt(e->covar, f->A_init, f->A_init, n, n, n);
      gko_as_sym(e->covar, n);
      break;
    case GKO_SRIF:
      dcopy(e->raw_vec, f->x_init, n);
      dcopy(e->raw_mat, f->A_init, (size_t)n * n);
      dcopy(e->raw_pred_mat, f->A_init, (size_t)n * n);
      srif_state(f->A_init, f->x_init, n, e->state, f->ws);
      e->covar_ok = srif_covariance(f->A_init, n, e->covar, f->ws);
      e->pred_covar_ok = srif_covariance(f->A_init, n, e->pred_covar, f->ws);
      dcopy(e->innov, f->x_init, n);
      e->innov_len = n;
      break;
  }
}

/* ---- vanilla.go:128-220 ------------------------------------------------------------------------ */

static int vanilla_update(gko_filter* f, const double* y, const double* u, gko_estimate* e) {
  int n = f->n, m = f->m, c = f->c, ierr;
  if (m != f->m_r) return GKO_ERR_DIMS; /* mat64 would panic on HPHt + R */
  double* ws = f->ws;
  double* xm = ws; ws += n;          /* xKp1Minus */
  double* t1 = ws; ws += n;
  double* wk = ws; ws += n;
  double* vk = ws; ws += m;
  double* FP = ws; ws += n * n;
  double* Pm = ws; ws += n * n;      /* Pkp1Minus (dense, not yet symmetrised) */
  double* yhat = ws; ws += m;
  double* PHt = ws; ws += n * m;
  double* S = ws; ws += m * m;
  double* K = ws; ws += n * m;
  double* innov = ws; ws += m;
  double* xp = ws; ws += n;
  double* KH = ws; ws += n * n;
  double* T1 = ws; ws += n * n;
  double* Pp = ws; ws += n * n;
  double* KR = ws; ws += n * m;
  double* KRK = ws; ws += n * n;
  double* Ht = ws; ws += n * m;

  /* 138-146: x- = F x (+ G u) + Process(k) */
  gko_mulvec(xm, f->F, f->x, n, n);
  if (f->need_ctrl) {
    if (!u) return GKO_ERR_DIMS;
    gko_mulvec(t1, f->G, u, n, c);
    for (int i = 0; i < n; ++i) xm[i] = xm[i] + t1[i];
  }
  if ((ierr = noise_process(f, f->step, wk)) != 0) return ierr;
  for (int i = 0; i < n; ++i) xm[i] = xm[i] + wk[i];
  /* 149-152: P- = (F P) F^T + Q */
  gko_mul(FP, f->F, f->A, n, n, n);
  gko_mul_nt(Pm, FP, f->F, n, n, n);
  for (int i = 0; i < n * n; ++i) Pm[i] = Pm[i] + f->Q[i];
  /* 155-157: yhat = H x_prev + Measurement(k)   (previous posterior, not x-) */
  gko_mulvec(yhat, f->H, f->x, m, n);
  if ((ierr = noise_measurement(f, f->step, vk)) != 0) return ierr;
  for (int i = 0; i < m; ++i) yhat[i] = yhat[i] + vk[i];
  /* 160-168: K = P- H^T inv(H P- H^T + R) */
  gko_transpose(Ht, f->H, m, n);
  gko_mul(PHt, Pm, Ht, n, n, m);
  gko_mul(S, f->H, PHt, m, n, m);
  for (int i = 0; i < m * m; ++i) S[i] = S[i] + f->R[i];
  if (gko_inverse(S, S, m, NULL) != 0) return GKO_ERR_SINGULAR_S;
  gko_mul(K, PHt, S, n, m, m);

  est_clear(e, n, m);
  if (f->prediction_only) { /* 170-179 */
    gko_as_sym(Pm, n);      /* error ignored at 173 */
    dcopy(e->state, xm, n);
    dcopy(e->meas, yhat, m);
    dcopy(e->covar, Pm, (size_t)n * n);
    dcopy(e->pred_covar, Pm, (size_t)n * n);
    dcopy(e->gain, K, (size_t)n * m);
    dcopy(f->x, xm, n);
    dcopy(f->A, Pm, (size_t)n * n);
    f->step++;
    return 0;
  }
  /* 182-195: innovation and state update (the m == 1 branch is the same arithmetic) */
  gko_mulvec(t1, f->H, xm, m, n);
  for (int i = 0; i < m; ++i) innov[i] = y[i] - t1[i];
  if (m == 1) {
    for (int i = 0; i < n; ++i) xp[i] = innov[0] * K[i] + 0.0; /* Scale, then AddVec with zeros */
  } else {
    gko_mulvec(xp, K, innov, n, m);
  }
  for (int i = 0; i < n; ++i) xp[i] = xm[i] + xp[i];
  if ((ierr = noise_process_again(f, f->step, wk)) != 0) return ierr; /* second Process(k) call, 195 */
  for (int i = 0; i < n; ++i) xp[i] = xp[i] + wk[i];
  /* 197-205: Joseph form */
  gko_mul(KH, K, f->H, n, m, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) KH[i * n + j] = (i == j ? 1.0 : 0.0) - KH[i * n + j];
  gko_mul(T1, KH, Pm, n, n, n);
  gko_mul_nt(Pp, T1, KH, n, n, n);
  gko_mul(KR, K, f->R, n, m, m);
  gko_mul_nt(KRK, KR, K, n, m, n);
  for (int i = 0; i < n * n; ++i) Pp[i] = Pp[i] + KRK[i];
  /* 207-215 */
  if (gko_as_sym(Pm, n) != 0) return GKO_ERR_ASYMMETRIC;
  if (gko_as_sym(Pp, n) != 0) return GKO_ERR_ASYMMETRIC;
  dcopy(e->state, xp, n);
  dcopy(e->meas, yhat, m);
  dcopy(e->innov, innov, m);
  dcopy(e->covar, Pp, (size_t)n * n);
  dcopy(e->pred_covar, Pm, (size_t)n * n);
  dcopy(e->gain, K, (size_t)n * m);
  dcopy(f->x, xp, n);
  dcopy(f->A, Pp, (size_t)n * n);
  f->step++;
  return 0;
}

/* ---- information.go:153-227 --------------------------------------------------------------------- */

static int information_update(gko_filter* f, const double* y, const double* u, gko_estimate* e) {
  int n = f->n, m = f->m, c = f->c, ierr;
  double* ws = f->ws;
  double* IF = ws; ws += n * n;
  double* z = ws; ws += n * n;
  double* M = ws; ws += n * n;
  double* T = ws; ws += n * n;
  double* im = ws; ws += n;
  double* t1 = ws; ws += n;
  double* t2 = ws; ws += n;
  double* M1 = ws; ws += n * n;
  double* Im = ws; ws += n * n;
  double* Pprev = ws; ws += n * n;
  double* xprev = ws; ws += n;
  double* yhat = ws; ws += m;
  double* vk = ws; ws += m;
  double* HTR = ws; ws += n * m;
  double* ip = ws; ws += n;
  double* Ip = ws; ws += n * n;
  double* Ht = ws; ws += n * m;

  /* 161-165: z = Finv^T (I Finv) */
  gko_mul(IF, f->A, f->Finv, n, n, n);
  gko_mul_tn(z, f->Finv, IF, n, n, n);
  /* 169-174: M = -(z inv(z + Qinv)), Inverse error ignored */
  for (int i = 0; i < n * n; ++i) T[i] = z[i] + f->Qinv[i];
  gko_inverse(T, T, n, NULL);
  gko_mul(M, z, T, n, n, n);
  for (int i = 0; i < n * n; ++i) M[i] = -1.0 * M[i];
  /* 176-185: i- = (I + M)(Finv^T i + z G u) */
  gko_mulvec_t(im, f->Finv, f->x, n, n);
  if (f->need_ctrl) {
    if (!u) return GKO_ERR_DIMS;
    gko_mulvec(t1, f->G, u, n, c);
    gko_mulvec(t2, z, t1, n, n);
    for (int i = 0; i < n; ++i) im[i] = im[i] + t2[i];
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) M1[i * n + j] = (i == j ? 1.0 : 0.0) + M[i * n + j];
  gko_mulvec(t1, M1, im, n, n);
  dcopy(im, t1, n);
  /* 187-190: I- = z + M z^T */
  gko_mul_nt(Im, M, z, n, n, n);
  for (int i = 0; i < n * n; ++i) Im[i] = z[i] + Im[i];
  /* 192-194: yhat = H State(prev) + Measurement(k); State() = Covariance() i (248-255) */
  info_covariance(f->A, n, Pprev);
  gko_mulvec(xprev, Pprev, f->x, n, n);
  gko_mulvec(yhat, f->H, xprev, m, n);
  if ((ierr = noise_measurement(f, f->step, vk)) != 0) return ierr;
  for (int i = 0; i < m; ++i) yhat[i] = yhat[i] + vk[i];
  /* 196-203: HTR = Rinv[0,0] H^T when Rinv is 1x1 (even if H has several rows), else H^T Rinv */
  gko_transpose(Ht, f->H, m, n);
  if (f->rinv_dim == 1) {
    for (int i = 0; i < n * m; ++i) HTR[i] = f->Rinv[0] * Ht[i];
  } else {
    if (f->rinv_dim != m) return GKO_ERR_DIMS; /* mat64 panics */
    gko_mul(HTR, Ht, f->Rinv, n, m, m);
  }
  /* 205-212 */
  gko_mulvec(ip, HTR, y, n, m);
  for (int i = 0; i < n; ++i) ip[i] = ip[i] + im[i];
  gko_mul(Ip, HTR, f->H, n, m, n);
  for (int i = 0; i < n * n; ++i) Ip[i] = Im[i] + Ip[i];
  /* 214-222: the reference panics on asymmetry */
  if (gko_as_sym(Im, n) != 0) return GKO_ERR_ASYMMETRIC;
  if (gko_as_sym(Ip, n) != 0) return GKO_ERR_ASYMMETRIC;

  est_clear(e, n, m);
  dcopy(e->raw_vec, ip, n);
  dcopy(e->raw_mat, Ip, (size_t)n * n);
  dcopy(e->raw_pred_mat, Im, (size_t)n * n);
  dcopy(e->meas, yhat, m);
  dcopy(e->innov, ip, n); /* Innovation() returns the information state (information.go:272-274) */
  e->innov_len = n;
  e->covar_ok = info_covariance(Ip, n, e->covar);
  gko_mulvec(e->state, e->covar, ip, n, n);
  e->pred_covar_ok = info_covariance(Im, n, e->pred_covar);
  dcopy(f->x, ip, n);
  dcopy(f->A, Ip, (size_t)n * n);
  f->step++;
  return 0;
}

/* ---- squareroot.go:129-274 ---------------------------------------------------------------------- */

static int sqrt_update(gko_filter* f, const double* y, const double* u, gko_estimate* e) {
  int n = f->n, m = f->m, c = f->c, ierr;
  if (m != f->m_r) return GKO_ERR_DIMS;
  int d = n + m;
  double* ws = f->ws;
  double* xm = ws; ws += n;
  double* t1 = ws; ws += n;
  double* Cm = ws; ws += 2 * n * n;
  double* Uc = ws; ws += 2 * n * n;
  double* Sm = ws; ws += n * n;   /* SKp1Minus = top block of R (upper triangular) */
  double* SmT = ws; ws += n * n;
  double* B = ws; ws += n * m;    /* SKp1Minus^T H^T */
  double* Ht = ws; ws += n * m;
  double* Dl = ws; ws += d * d;
  double* Ud = ws; ws += d * d;
  double* Sp = ws; ws += n * n;
  double* Syy = ws; ws += m * m;
  double* W = ws; ws += n * m;
  double* SyyInv = ws; ws += m * m;
  double* K = ws; ws += n * m;
  double* yhat = ws; ws += m;
  double* vk = ws; ws += m;
  double* wk = ws; ws += n;
  double* innov = ws; ws += m;
  double* xp = ws; ws += n;

  /* 140-147: x- = F x (+ G u); no process noise here */
  gko_mulvec(xm, f->F, f->x, n, n);
  if (f->need_ctrl) {
    if (!u) return GKO_ERR_DIMS;
    gko_mulvec(t1, f->G, u, n, c);
    for (int i = 0; i < n; ++i) xm[i] = xm[i] + t1[i];
  }
  /* 155-175: C = [S^T F^T ; sqrtQ^T] */
  {
    double* St = Dl; /* scratch */
    double* Ft = Ud;
    gko_transpose(St, f->A, n, n);
    gko_transpose(Ft, f->F, n, n);
    gko_mul(Cm, St, Ft, n, n, n);
    gko_transpose(Cm + n * n, f->sqrtQ, n, n);
  }
  /* 176-185: S- := top n x n of R from QR(C)  (the UPPER factor is used as if it were S-) */
  gko_qr_r(Uc, Cm, 2 * n, n);
  dcopy(Sm, Uc, (size_t)n * n);
  gko_transpose(SmT, Sm, n, n);
  /* 190-216: Delta = [[sqrtR^T, 0],[S-^T H^T, S-^T]] */
  gko_transpose(Ht, f->H, m, n);
  gko_mul(B, SmT, Ht, n, n, m);
  for (int r = 0; r < d; ++r)
    for (int cc = 0; cc < d; ++cc) {
      double v;
      if (cc < m) {
        if (r < m) v = f->sqrtR[cc * m + r]; /* sqrtR^T */
        else v = B[(r - m) * m + cc];
      } else if (r < m) {
        v = 0.0;
      } else {
        v = SmT[(r - m) * n + (cc - m)];
      }
      Dl[r * d + cc] = v;
    }
  /* 218-234 */
  gko_qr_r(Ud, Dl, d, d);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Sp[i * n + j] = Ud[(m + j) * d + (m + i)]; /* (U[m:,m:])^T */
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) Syy[i * m + j] = Ud[j * d + i];             /* (U[:m,:m])^T */
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) W[i * m + j] = Ud[j * d + (m + i)];         /* (U[:m,m:])^T */
  /* 237-239: yhat = H x_prev + Measurement(k) */
  gko_mulvec(yhat, f->H, f->x, m, n);
  if ((ierr = noise_measurement(f, f->step, vk)) != 0) return ierr;
  for (int i = 0; i < m; ++i) yhat[i] = yhat[i] + vk[i];
  /* 242-252: K = W inv(Syy); the error check at 243 tests the wrong variable => never fires */
  gko_inverse(SyyInv, Syy, m, NULL);
  if (m == 1) {
    for (int i = 0; i < n; ++i) K[i] = SyyInv[0] * W[i];
  } else {
    gko_mul(K, W, SyyInv, n, m, m);
  }
  /* 255-268 */
  gko_mulvec(t1, f->H, xm, m, n);
  for (int i = 0; i < m; ++i) innov[i] = y[i] - t1[i];
  if (m == 1) {
    for (int i = 0; i < n; ++i) xp[i] = innov[0] * K[i] + 0.0;
  } else {
    gko_mulvec(xp, K, innov, n, m);
  }
  for (int i = 0; i < n; ++i) xp[i] = xm[i] + xp[i];
  if ((ierr = noise_process(f, f->step, wk)) != 0) return ierr;
  for (int i = 0; i < n; ++i) xp[i] = xp[i] + wk[i];

  est_clear(e, n, m);
  dcopy(e->state, xp, n);
  dcopy(e->meas, yhat, m);
  dcopy(e->innov, innov, m);
  dcopy(e->gain, K, (size_t)n * m);
  dcopy(e->raw_mat, Sp, (size_t)n * n);
  dcopy(e->raw_pred_mat, Sm, (size_t)n * n);
  /* 317-340: Covariance() = S S^T, PredCovariance() = S- S-^T (AsSymDense errors dropped) */
  gko_mul_nt(e->covar, Sp, Sp, n, n, n);
  gko_as_sym(e->covar, n);
  gko_mul_nt(e->pred_covar, Sm, Sm, n, n, n);
  gko_as_sym(e->pred_covar, n);
  dcopy(f->x, xp, n);
  dcopy(f->A, Sp, (size_t)n * n);
  f->step++;
  return 0;
}

int gko_update(gko_filter* f, const double* y, const double* u, gko_estimate* e) {
  switch (f->kind) {
    case GKO_VANILLA:
    case GKO_PREDICTOR: return vanilla_update(f, y, u, e);
    case GKO_INFORMATION: return information_update(f, y, u, e);
    case GKO_SQRT: return sqrt_update(f, y, u, e);
    default: return GKO_ERR_DIMS;
  }
}

/* ---- hybrid.go -------------------------------------------------------------------------------- */

void gko_prepare(gko_filter* f, const double* Phi, const double* Htilde) { /* hybrid.go:78-82, srif.go:82-86 */
  dcopy(f->Phi, Phi, (size_t)f->n * f->n);
  f->has_htilde = Htilde != NULL;
  if (Htilde) dcopy(f->Htilde, Htilde, (size_t)f->m * f->n);
  f->locked = 0;
}

void gko_prepare_pnt(gko_filter* f, const double* Gamma) { /* hybrid.go:86-89; srif.go:75 is a no-op */
  if (f->kind != GKO_HYBRID) return;
  dcopy(f->Gamma, Gamma, (size_t)f->n * f->q);
  f->snc = 1;
}

void gko_enable_ekf(gko_filter* f, int on) { /* hybrid.go:54-61; srif.go:66-72 no-ops */
  if (f->kind == GKO_HYBRID) f->ekf = on ? 1 : 0;
}

void gko_srif_set_non_tri_r(gko_filter* f, int v) { f->non_tri_r = v; }

static int hybrid_full_update(gko_filter* f, int pure_prediction, const double* real_obs,
                              const double* computed_obs, gko_estimate* e) { /* hybrid.go:104-204 */
  int n = f->n, m = f->m, q = f->q;
  if (f->locked) return GKO_ERR_LOCKED;
  double* ws = f->ws;
  double* PhiP = ws; ws += n * n;
  double* Pbar = ws; ws += n * n;
  double* GQ = ws; ws += n * (q > 0 ? q : 1);
  double* GQG = ws; ws += n * n;
  double* xbar = ws; ws += n;
  double* Ht = ws; ws += n * m;
  double* PHt = ws; ws += n * m;
  double* S = ws; ws += m * m;
  double* K = ws; ws += n * m;
  double* yv = ws; ws += m;
  double* innov = ws; ws += m;
  double* xhat = ws; ws += n;
  double* t1 = ws; ws += (n > m ? n : m);
  double* KH = ws; ws += n * n;
  double* T1 = ws; ws += n * n;
  double* P = ws; ws += n * n;
  double* KR = ws; ws += n * m;
  double* KRK = ws; ws += n * n;

  /* 114-123 */
  gko_mul(PhiP, f->Phi, f->A, n, n, n);
  gko_mul_nt(Pbar, PhiP, f->Phi, n, n, n);
  if (f->snc) {
    gko_mul(GQ, f->Gamma, f->Q, n, q, q);
    gko_mul_nt(GQG, GQ, f->Gamma, n, q, n);
    for (int i = 0; i < n * n; ++i) Pbar[i] = Pbar[i] + GQG[i];
  }
  if (pure_prediction) { /* 125-143 */
    if (f->ekf) {
      for (int i = 0; i < n; ++i) xbar[i] = 0.0; /* 128: literal zeros(6); generalised to n */
    } else {
      gko_mulvec(xbar, f->Phi, f->x, n, n);
    }
    if (gko_as_sym(Pbar, n) != 0) return GKO_ERR_ASYMMETRIC;
    est_clear(e, n, m);
    dcopy(e->state, xbar, n);
    dcopy(e->covar, Pbar, (size_t)n * n);
    dcopy(e->pred_covar, Pbar, (size_t)n * n);
    dcopy(f->x, xbar, n);
    dcopy(f->A, Pbar, (size_t)n * n);
    f->step++;
    f->snc = 0;
    f->locked = 1;
    return 0;
  }
  if (!f->has_htilde) return GKO_ERR_DIMS;
  /* 146-153 */
  gko_transpose(Ht, f->Htilde, m, n);
  gko_mul(PHt, Pbar, Ht, n, n, m);
  gko_mul(S, f->Htilde, PHt, m, n, m);
  for (int i = 0; i < m * m; ++i) S[i] = S[i] + f->R[i];
  if (gko_inverse(S, S, m, NULL) != 0) return GKO_ERR_SINGULAR_S;
  gko_mul(K, PHt, S, n, m, m);
  /* 156-157 */
  for (int i = 0; i < m; ++i) yv[i] = real_obs[i] - computed_obs[i];
  for (int i = 0; i < m; ++i) innov[i] = 0.0; /* EKF leaves innov an empty vector (159-161) */
  if (f->ekf) {
    gko_mulvec(xhat, K, yv, n, m);
  } else { /* 162-173 */
    gko_mulvec(xbar, f->Phi, f->x, n, n);
    gko_mulvec(t1, f->Htilde, xbar, m, n);
    for (int i = 0; i < m; ++i) innov[i] = yv[i] - t1[i];
    gko_mulvec(xhat, K, innov, n, m);
    for (int i = 0; i < n; ++i) xhat[i] = xbar[i] + xhat[i];
  }
  /* 174-182 */
  gko_mul(KH, K, f->Htilde, n, m, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) KH[i * n + j] = (i == j ? 1.0 : 0.0) - KH[i * n + j];
  gko_mul(T1, KH, Pbar, n, n, n);
  gko_mul_nt(P, T1, KH, n, n, n);
  gko_mul(KR, K, f->R, n, m, m);
  gko_mul_nt(KRK, KR, K, n, m, n);
  for (int i = 0; i < n * n; ++i) P[i] = P[i] + KRK[i];
  /* 184-192 */
  if (gko_as_sym(Pbar, n) != 0) return GKO_ERR_ASYMMETRIC;
  if (gko_as_sym(P, n) != 0) return GKO_ERR_ASYMMETRIC;
  est_clear(e, n, m);
  dcopy(e->state, xhat, n);
  dcopy(e->meas, real_obs, m);
  dcopy(e->innov, innov, m);
  e->innov_len = f->ekf ? 0 : m;
  dcopy(e->obs_dev, yv, m);
  dcopy(e->covar, P, (size_t)n * n);
  dcopy(e->pred_covar, Pbar, (size_t)n * n);
  dcopy(e->gain, K, (size_t)n * m);
  dcopy(f->x, xhat, n);
  dcopy(f->A, P, (size_t)n * n);
  f->step++;
  f->snc = 0;
  f->locked = 1;
  return 0;
}

/* ---- srif.go ---------------------------------------------------------------------------------- */

void gko_measurement_srif_update(int n, int m, const double* R, const double* H, const double* b,
                                 const double* y, double* Rk, double* bk, double* ek) { /* srif.go:298-340 */
  int rows = n + m, cols = n + 1;
  double* A = dalloc((size_t)rows * cols);
  for (int i = 0; i < rows; ++i) {
    for (int j = 0; j < n; ++j) A[i * cols + j] = (i < n) ? R[i * n + j] : H[(i - n) * n + j];
    A[i * cols + n] = (i < n) ? b[i] : y[i - n];
  }
  gko_householder_transf(A, n, m);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) Rk[i * n + j] = A[i * cols + j];
    bk[i] = A[i * cols + n];
  }
  for (int i = 0; i < m; ++i) ek[i] = A[(n + i) * cols + n];
  free(A);
}

static int srif_full_update(gko_filter* f, int pure_prediction, const double* real_obs,
                            const double* computed_obs, gko_estimate* e) { /* srif.go:101-160 */
  int n = f->n, m = f->m, ierr;
  if (f->locked) return GKO_ERR_LOCKED;
  double* ws = f->ws;
  double* scratch = ws; ws += n * n;
  double* invPhi = ws; ws += n * n;
  double* Rbar = ws; ws += n * n;
  double* xprev = ws; ws += n;
  double* xbar = ws; ws += n;
  double* bbar = ws; ws += n;
  double* yv = ws; ws += m;
  double* yw = ws; ws += m;
  double* Hw = ws; ws += m * n;
  double* Rk = ws; ws += n * n;
  double* bk = ws; ws += n;
  double* ek = ws; ws += m;

  /* 110-115 */
  if (gko_inverse(invPhi, f->Phi, n, NULL) != 0) return GKO_ERR_SINGULAR_PHI;
  gko_mul(Rbar, f->A, invPhi, n, n, n);
  /* 117-119: xbar = Phi State(prev), bbar = Rbar xbar */
  if ((ierr = srif_state(f->A, f->x, n, xprev, scratch)) != 0) return ierr;
  gko_mulvec(xbar, f->Phi, xprev, n, n);
  gko_mulvec(bbar, Rbar, xbar, n, n);
  /* 121-132: the "!nonTriR" branch augments [Rbar | bbar] and slices it back: values unchanged */
  (void)f->non_tri_r;

  if (pure_prediction) { /* 134-141 */
    est_clear(e, n, m);
    dcopy(e->raw_vec, bbar, n);
    dcopy(e->raw_mat, Rbar, (size_t)n * n);
    dcopy(e->raw_pred_mat, Rbar, (size_t)n * n);
    dcopy(e->innov, bbar, n);
    e->innov_len = n;
    dcopy(f->x, bbar, n);
    dcopy(f->A, Rbar, (size_t)n * n);
    f->step++;
    f->locked = 1;
    ierr = srif_state(Rbar, bbar, n, e->state, scratch);
    e->covar_ok = srif_covariance(Rbar, n, e->covar, scratch);
    e->pred_covar_ok = e->covar_ok;
    dcopy(e->pred_covar, e->covar, (size_t)n * n);
    return ierr;
  }
  if (!f->has_htilde) return GKO_ERR_DIMS;
  /* 143-148: whitening multiplies by L = chol(R), not by its inverse */
  for (int i = 0; i < m; ++i) yv[i] = real_obs[i] - computed_obs[i];
  gko_mul(Hw, f->sqrt_inv_noise, f->Htilde, m, m, n);
  gko_mulvec(yw, f->sqrt_inv_noise, yv, m, m);
  /* 150 */
  gko_measurement_srif_update(n, m, Rbar, Hw, bbar, yw, Rk, bk, ek);
  est_clear(e, n, m);
  dcopy(e->raw_vec, bk, n);
  dcopy(e->raw_mat, Rk, (size_t)n * n);
  dcopy(e->raw_pred_mat, Rbar, (size_t)n * n);
  dcopy(e->meas, real_obs, m);
  dcopy(e->obs_dev, yw, m); /* Delta-obs is the whitened deviation (srif.go:154) */
  dcopy(e->innov, bk, n);   /* Innovation() returns the sqrt-information state (238-240) */
  e->innov_len = n;
  dcopy(f->x, bk, n);
  dcopy(f->A, Rk, (size_t)n * n);
  f->step++;
  f->locked = 1;
  ierr = srif_state(Rk, bk, n, e->state, scratch);
  e->covar_ok = srif_covariance(Rk, n, e->covar, scratch);
  e->pred_covar_ok = srif_covariance(Rbar, n, e->pred_covar, scratch);
  return ierr;
}

int gko_nl_predict(gko_filter* f, gko_estimate* e) {
  if (f->kind == GKO_HYBRID) return hybrid_full_update(f, 1, NULL, NULL, e);
  if (f->kind == GKO_SRIF) return srif_full_update(f, 1, NULL, NULL, e);
  return GKO_ERR_DIMS;
}

int gko_nl_update(gko_filter* f, const double* real_obs, const double* computed_obs, gko_estimate* e) {
  if (f->kind == GKO_HYBRID) return hybrid_full_update(f, 0, real_obs, computed_obs, e);
  if (f->kind == GKO_SRIF) return srif_full_update(f, 0, real_obs, computed_obs, e);
  return GKO_ERR_DIMS;
}

int gko_smooth_all(int n, int steps, const double* Phi, double* x, double* P) {
  /* hybrid.go:209-238 (no-SNC branch) == srif.go:165-192 */
  double* S = dalloc((size_t)n * n);
  double* SP = dalloc((size_t)n * n);
  double* SPS = dalloc((size_t)n * n);
  double* xh = dalloc(n);
  int rc = 0;
  for (int k = steps - 2; k >= 0; --k) {
    const double* Phi1 = Phi + (size_t)(k + 1) * n * n;
    if (gko_inverse(S, Phi1, n, NULL) != 0) { rc = GKO_ERR_SINGULAR_PHI; break; }
    gko_mul(SP, S, P + (size_t)(k + 1) * n * n, n, n, n);
    gko_mul_nt(SPS, SP, S, n, n, n);
    gko_mulvec(xh, S, x + (size_t)(k + 1) * n, n, n);
    if (gko_as_sym(SPS, n) != 0) { rc = GKO_ERR_ASYMMETRIC; break; }
    dcopy(x + (size_t)k * n, xh, n);
    dcopy(P + (size_t)k * n * n, SPS, (size_t)n * n);
  }
  free(S); free(SP); free(SPS); free(xh);
  return rc;
}

/* ---- batch runners: one independent filter object per filter, OpenMP over filters ----------------
 * (the goroutine-per-filter structure a Go caller would use; bench.py's CPU baseline of the NLDKF and
 * large-state workloads, and a batch-sized parity checker).  Streams use the engine's SoA layout
 * [step][component][filter] for the NLDKF run and the filter-major layout for the LDKF run. */
int gko_run_nl_batch(int kind, int n, int m, int64_t nf, int steps, const uint8_t* flags, const double* x0,
                     const double* P0, const double* R, const double* Phi, const double* Htilde,
                     const double* real_obs, const double* computed_obs, int threads, double* out_state,
                     double* out_covar) {
  int rc_all = 0;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
  for (int64_t f = 0; f < nf; ++f) {
    gko_filter* kf = kind == GKO_SRIF ? gko_new_srif(n, m, x0, P0, R, 0) : gko_new_hybrid(n, m, 0, x0, P0, NULL, R);
    gko_estimate* e = (gko_estimate*)malloc(sizeof(gko_estimate));
    double Ph[GKO_MAXN * GKO_MAXN], Ht[GKO_MAXM * GKO_MAXN], ro[GKO_MAXM], co[GKO_MAXM];
    int rc = 0;
    for (int k = 0; k < steps && rc == 0; ++k) {
      for (int i = 0; i < n * n; ++i) Ph[i] = Phi[((size_t)k * n * n + i) * nf + f];
      for (int i = 0; i < m * n; ++i) Ht[i] = Htilde[((size_t)k * m * n + i) * nf + f];
      for (int a = 0; a < m; ++a) {
        ro[a] = real_obs[((size_t)k * m + a) * nf + f];
        co[a] = computed_obs[((size_t)k * m + a) * nf + f];
      }
      const unsigned fl = flags ? flags[k] : 1u;
      gko_prepare(kf, Ph, Ht);
      gko_enable_ekf(kf, (fl & 2u) != 0);
      rc = (fl & 1u) ? gko_nl_update(kf, ro, co, e) : gko_nl_predict(kf, e);
    }
    if (rc == 0) {
      for (int i = 0; i < n; ++i) out_state[(size_t)i * nf + f] = e->state[i];
      for (int i = 0; i < n * n; ++i) out_covar[(size_t)i * nf + f] = e->covar[i];
    } else {
#pragma omp critical
      rc_all = rc;
    }
    free(e);
    gko_free(kf);
  }
  return rc_all;
}

int gko_run_vanilla_batch(int n, int m, int64_t nf, int steps, const double* x0, const double* P0, const double* F,
                          const double* H, const double* Q, const double* R, const double* y, int threads,
                          double* out_state, double* out_covar) {
  /* y [steps][nf][m], out_state [nf][n], out_covar [nf][n*n] (filter-major, like the large-state handles) */
  int rc_all = 0;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
  for (int64_t f = 0; f < nf; ++f) {
    gko_filter* kf = gko_new_vanilla(n, m, 0, x0, P0, F, NULL, H, Q, R, 0);
    gko_estimate* e = (gko_estimate*)malloc(sizeof(gko_estimate));
    int rc = 0;
    for (int k = 0; k < steps && rc == 0; ++k) rc = gko_update(kf, y + ((size_t)k * nf + f) * m, NULL, e);
    if (rc == 0) {
      for (int i = 0; i < n; ++i) out_state[(size_t)f * n + i] = e->state[i];
      if (out_covar)
        for (int i = 0; i < n * n; ++i) out_covar[(size_t)f * n * n + i] = e->covar[i];
    } else {
#pragma omp critical
      rc_all = rc;
    }
    free(e);
    gko_free(kf);
  }
  return rc_all;
}

/* ---- BatchKF (batch.go:34-79) -------------------------------------------------------------------- */
int gko_batch_solve(int n, int m, int count, const double* R, const double* H, const double* real_obs,
                    const double* computed_obs, double* xhat0, double* P0) {
  /* SetNextMeasurement (batch.go:41-61), `count` times: Lambda += (H^T R) H, N += (H^T R) y with
   * y = real - computed.  The reference multiplies by R, not inv(R) (batch.go:50): kept.
   * Solve (batch.go:64-79): P0 = AsSymDense(inv(Lambda)), xHat0 = P0 N. */
  double* Lam = dalloc((size_t)n * n);
  double* Nv = dalloc(n);
  double* HtR = dalloc((size_t)n * m);
  double* HtRH = dalloc((size_t)n * n);
  double* y = dalloc(m);
  double* t = dalloc(n);
  int rc = 0;
  for (int k = 0; k < count; ++k) {
    const double* Hk = H + (size_t)k * m * n;
    gko_mul_tn(HtR, Hk, R, n, m, m);
    gko_mul(HtRH, HtR, Hk, n, m, n);
    for (int i = 0; i < n * n; ++i) Lam[i] = Lam[i] + HtRH[i];
    for (int a = 0; a < m; ++a) y[a] = real_obs[(size_t)k * m + a] - computed_obs[(size_t)k * m + a];
    gko_mulvec(t, HtR, y, n, m);
    for (int i = 0; i < n; ++i) Nv[i] = Nv[i] + t[i];
  }
  if (gko_inverse(P0, Lam, n, NULL) != 0) rc = GKO_ERR_SINGULAR_S;
  else if (gko_as_sym(P0, n) != 0) rc = GKO_ERR_ASYMMETRIC;
  else gko_mulvec(xhat0, P0, Nv, n, n);
  free(Lam); free(Nv); free(HtR); free(HtRH); free(y); free(t);
  return rc;
}

/* ---- Philox4x32-10 + inverse normal CDF: the oracle's own noise stream --------------------------------- */

void gko_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* The engine's Gaussian transform (include/gokalman_b200_icdf.inc, tools/gen_icdf_table.py): one 32-bit word k
 * -> u = (k + 0.5) 2^-32 -> z = Phi^-1(u) by a quintic in the low 27 bits of the normalised tail probability on one
 * of 512 segments (16 per binary octave).  Same table, same fused Horner evaluation as the kernels: bit-identical. */
static const double gko_icdf_table[512 * 6] = {
#include "../include/gokalman_b200_icdf.inc"
};

double gko_icdf_normal(uint32_t k) {
  const uint32_t upper = k >> 31;
  const uint32_t j = upper ? ~k : k;
  const uint32_t J = (j << 1) | 1u;
  const int lz = __builtin_clz(J);
  const uint32_t Jn = J << lz;
  const uint32_t seg = ((31u - (uint32_t)lz) << 4) | ((Jn >> 27) & 15u);
  const double v = (double)(Jn & 0x07ffffffu);
  const double* c = gko_icdf_table + seg * 6;
  double g = fma(c[5], v, c[4]);
  g = fma(g, v, c[3]);
  g = fma(g, v, c[2]);
  g = fma(g, v, c[1]);
  g = fma(g, v, c[0]);
  return upper ? g : -g;
}

void gko_philox_normals(uint64_t seed, uint64_t trial, uint32_t step, int count, double* z) {
  /* normal 4b + i is word i of block b: counter = (trial_lo, trial_hi, step, b), key = seed */
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int b = 0; 4 * b < count; ++b) {
    uint32_t ctr[4] = {(uint32_t)trial, (uint32_t)(trial >> 32), step, (uint32_t)b};
    uint32_t o[4];
    gko_philox4x32_10(ctr, key, o);
    for (int i = 0; i < 4 && 4 * b + i < count; ++i) z[4 * b + i] = gko_icdf_normal(o[i]);
  }
}

/* ---- montecarlo.go:92-119 + chisquare.go:16-95 ------------------------------------------------ */

int gko_mc_chisquare(const gko_mc_config* cfg, double* nis_means, double* nees_means,
                     double* mean_state, double* std_state, double* truth_x_out, double* truth_y_out) {
  int n = cfg->n, m = cfg->m, c = cfg->c, steps = cfg->steps, trials = cfg->trials;
  if (!cfg->with_nees && !cfg->with_nis) return GKO_ERR_DIMS; /* chisquare.go:17-19 */
  if (cfg->kind == GKO_INFORMATION && cfg->with_nis && n != m) return GKO_ERR_DIMS; /* mat64 panics */
  size_t ns = (size_t)steps * trials;
  double* nees_s = cfg->with_nees ? dalloc(ns) : NULL; /* NEESsamples[k][run] */
  double* nis_s = cfg->with_nis ? dalloc(ns) : NULL;
  double* xs = (mean_state || std_state) ? dalloc(ns * n) : NULL; /* [k][i][run] */
  double* zero_u = dalloc(c > 0 ? c : 1);
  double* LQ = dalloc((size_t)n * n);
  double* LR = dalloc((size_t)m * m);
  if (!cfg->w) { /* distmv.NewNormal: mu + L z with L = chol(Sigma) (noise.go:146-153) */
    gko_chol_lower(LQ, cfg->Q, n);
    gko_chol_lower(LR, cfg->R, m);
  }
  int rc = 0;
  int nthreads = cfg->threads > 1 ? cfg->threads : 1;
  /* the tested filter's own model (chisquare.go:16: any LDKF) */
  const double* tF = cfg->tF ? cfg->tF : cfg->F;
  const double* tG = cfg->tG ? cfg->tG : cfg->G;
  const double* tH = cfg->tH ? cfg->tH : cfg->H;
  const double* tQ = cfg->tQ ? cfg->tQ : cfg->Q;
  const double* tR = cfg->tR ? cfg->tR : cfg->R;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
  {
    /* one truth generator and one tested filter per thread; Reset() between samples */
    gko_filter* truth = gko_new_vanilla(n, m, c, cfg->x0_truth, cfg->P0, cfg->F, cfg->G, cfg->H, cfg->Q,
                                        cfg->R, 1);
    gko_filter* kf = NULL;
    switch (cfg->kind) {
      case GKO_VANILLA:
        kf = gko_new_vanilla(n, m, c, cfg->x0_filter, cfg->P0, tF, tG, tH, tQ, tR, 0);
        break;
      case GKO_INFORMATION:
        kf = gko_new_information_from_state(n, m, c, cfg->x0_filter, cfg->P0, tF, tG, tH, tQ, tR);
        break;
      case GKO_SQRT:
        kf = gko_new_sqrt(n, m, c, cfg->x0_filter, cfg->P0, tF, tG, tH, tQ, tR);
        break;
    }
    gko_estimate* te = (gko_estimate*)malloc(sizeof(gko_estimate));
    gko_estimate* fe = (gko_estimate*)malloc(sizeof(gko_estimate));
    double* tx = dalloc((size_t)steps * n);
    double* ty = dalloc((size_t)steps * m);
    double* wv = dalloc((size_t)steps * n);
    double* vv = dalloc((size_t)steps * m);
    double* zz = dalloc((size_t)n + m + 4);
    double* Pinv = dalloc((size_t)n * n);
    double* t1 = dalloc((size_t)n > (size_t)m ? n : m);
    double* t2 = dalloc((size_t)n > (size_t)m ? n : m);
    double* Pyy0 = dalloc((size_t)n * m);
    double* Pyy = dalloc((size_t)m * m);
    double* Ht = dalloc((size_t)n * m);
    double* zero_y = dalloc(m);
    int local_rc = (truth && kf) ? 0 : GKO_ERR_DIMS;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int s = 0; s < trials; ++s) {
      if (local_rc) continue;
      /* -- montecarlo.go:108-117: one truth run -- */
      if (cfg->w) {
        gko_set_replay(truth, steps, cfg->w + (size_t)s * steps * n, cfg->v + (size_t)s * steps * m, m);
      } else {
        for (int k = 0; k < steps; ++k) {
          gko_philox_normals(cfg->seed, (uint64_t)(cfg->trial_offset + s), (uint32_t)k, n + m, zz);
          gko_mulvec(wv + (size_t)k * n, LQ, zz, n, n);
          gko_mulvec(vv + (size_t)k * m, LR, zz + n, m, m);
        }
        gko_set_replay(truth, steps, wv, vv, m);
      }
      gko_reset(truth);
      for (int k = 0; k < steps; ++k) {
        const double* u = cfg->controls ? cfg->controls + (size_t)k * c : zero_u;
        int r = gko_update(truth, zero_y, u, te); /* errors ignored: montecarlo.go:111 */
        (void)r;
        dcopy(tx + (size_t)k * n, te->state, n);
        dcopy(ty + (size_t)k * m, te->meas, m);
        if (xs)
          for (int i = 0; i < n; ++i) xs[((size_t)k * n + i) * trials + s] = te->state[i];
      }
      if (truth_x_out) dcopy(truth_x_out + (size_t)s * steps * n, tx, (size_t)steps * n);
      if (truth_y_out) dcopy(truth_y_out + (size_t)s * steps * m, ty, (size_t)steps * m);
      /* -- chisquare.go:36-79 -- */
      gko_reset(kf);
      for (int k = 0; k < steps; ++k) {
        const double* u = cfg->controls ? cfg->controls + (size_t)k * c : zero_u;
        int r = gko_update(kf, ty + (size_t)k * m, u, fe);
        if (r != 0) { local_rc = r; break; } /* chisquare.go:40-42 panics */
        if (nees_s) { /* 45-59 */
          gko_inverse(Pinv, fe->covar, n, NULL); /* error ignored */
          for (int i = 0; i < n; ++i) t1[i] = tx[(size_t)k * n + i] - fe->state[i];
          gko_mulvec(t2, Pinv, t1, n, n);
          double v = 0.0;
          for (int i = 0; i < n; ++i) v += t1[i] * t2[i];
          nees_s[(size_t)k * trials + s] = v;
        }
        if (nis_s) { /* 61-77 */
          gko_transpose(Ht, tH, m, n); /* kf.GetMeasurementMatrix(), kf.GetNoise().MeasurementMatrix(): 64-66 */
          gko_mul(Pyy0, fe->pred_covar, Ht, n, n, m);
          gko_mul(Pyy, tH, Pyy0, m, n, m);
          for (int i = 0; i < m * m; ++i) Pyy[i] = Pyy[i] + tR[i];
          gko_inverse(Pyy, Pyy, m, NULL);
          gko_mulvec(t2, Pyy, fe->innov, m, m);
          double v = 0.0;
          for (int i = 0; i < m; ++i) v += fe->innov[i] * t2[i];
          nis_s[(size_t)k * trials + s] = v;
        }
      }
    }
    if (local_rc) {
#ifdef _OPENMP
#pragma omp critical
#endif
      rc = local_rc;
    }
    gko_free(truth); gko_free(kf);
    free(te); free(fe); free(tx); free(ty); free(wv); free(vv); free(zz); free(Pinv);
    free(t1); free(t2); free(Pyy0); free(Pyy); free(Ht); free(zero_y);
  }
  /* chisquare.go:85-92: stat.Mean = sequential sum / runs */
  for (int k = 0; k < steps; ++k) {
    if (nees_s) {
      double s = 0.0;
      for (int r = 0; r < trials; ++r) s += nees_s[(size_t)k * trials + r];
      nees_means[k] = s / trials;
    } else if (nees_means) nees_means[k] = 0.0;
    if (nis_s) {
      double s = 0.0;
      for (int r = 0; r < trials; ++r) s += nis_s[(size_t)k * trials + r];
      nis_means[k] = s / trials;
    } else if (nis_means) nis_means[k] = 0.0;
  }
  /* montecarlo.go:18-59: stat.Mean and the unbiased stat.StdDev (two-pass, as gonum's stat) */
  if (xs) {
    for (int k = 0; k < steps; ++k)
      for (int i = 0; i < n; ++i) {
        const double* col = xs + ((size_t)k * n + i) * trials;
        double s = 0.0;
        for (int r = 0; r < trials; ++r) s += col[r];
        double mu = s / trials;
        if (mean_state) mean_state[(size_t)k * n + i] = mu;
        if (std_state) {
          double ss = 0.0, comp = 0.0;
          for (int r = 0; r < trials; ++r) {
            double d = col[r] - mu;
            ss += d * d;
            comp += d;
          }
          double var = (ss - comp * comp / trials) / (trials - 1);
          std_state[(size_t)k * n + i] = sqrt(var);
        }
      }
  }
  free(nees_s); free(nis_s); free(xs); free(zero_u); free(LQ); free(LR);
  return rc;
}
