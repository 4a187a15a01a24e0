/* gko_linalg.c -- see gko_linalg.h.  CPU oracle, TEST INFRASTRUCTURE ONLY. */
#include "gko_linalg.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- products: gonum dgemm/dgemv, sequential accumulation over the inner index ------------- */

void gko_mul(double* C, const double* A, const double* B, int r, int k, int c) {
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) {
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += A[i * k + l] * B[l * c + j];
      C[i * c + j] = s;
    }
}

void gko_mul_nt(double* C, const double* A, const double* B, int r, int k, int c) {
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) {
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += A[i * k + l] * B[j * k + l];
      C[i * c + j] = s;
    }
}

void gko_mul_tn(double* C, const double* A, const double* B, int r, int k, int c) {
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) {
      double s = 0.0;
      for (int l = 0; l < k; ++l) s += A[l * r + i] * B[l * c + j];
      C[i * c + j] = s;
    }
}

void gko_mulvec(double* y, const double* A, const double* x, int r, int c) {
  for (int i = 0; i < r; ++i) {
    double s = 0.0;
    for (int j = 0; j < c; ++j) s += A[i * c + j] * x[j];
    y[i] = s;
  }
}

void gko_mulvec_t(double* y, const double* A, const double* x, int r, int c) {
  for (int j = 0; j < c; ++j) {
    double s = 0.0;
    for (int i = 0; i < r; ++i) s += A[i * c + j] * x[i];
    y[j] = s;
  }
}

void gko_transpose(double* At, const double* A, int r, int c) {
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) At[j * r + i] = A[i * c + j];
}

/* ---- LU inverse: dgetf2 + dtrti2 + dgetri (unblocked), then the gonum condition test -------- */

int gko_inverse(double* Ainv, const double* A, int n, double* cond_out) {
  double* a = Ainv;
  int* ipiv = (int*)malloc(sizeof(int) * (size_t)n);
  double* work = (double*)malloc(sizeof(double) * (size_t)n);
  if (a != A) memcpy(a, A, sizeof(double) * (size_t)n * n);
  double anorm = 0.0; /* max row sum of the input (lapack.MaxRowSum) */
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += fabs(A[i * n + j]);
    if (s > anorm) anorm = s;
  }
  int singular = 0;
  /* dgetf2 */
  for (int j = 0; j < n; ++j) {
    int p = j;
    double pmax = fabs(a[j * n + j]);
    for (int i = j + 1; i < n; ++i) { /* idamax: first index of the largest |.| */
      double v = fabs(a[i * n + j]);
      if (v > pmax) { pmax = v; p = i; }
    }
    ipiv[j] = p;
    if (a[p * n + j] != 0.0) {
      if (p != j)
        for (int l = 0; l < n; ++l) {
          double t = a[j * n + l];
          a[j * n + l] = a[p * n + l];
          a[p * n + l] = t;
        }
      if (fabs(a[j * n + j]) >= DBL_MIN) {
        double rinv = 1.0 / a[j * n + j]; /* dscal by the reciprocal, as dgetf2 does */
        for (int i = j + 1; i < n; ++i) a[i * n + j] *= rinv;
      } else {
        for (int i = j + 1; i < n; ++i) a[i * n + j] /= a[j * n + j];
      }
    } else {
      singular = 1;
    }
    for (int i = j + 1; i < n; ++i) { /* dger trailing update */
      double lij = a[i * n + j];
      for (int l = j + 1; l < n; ++l) a[i * n + l] -= lij * a[j * n + l];
    }
  }
  if (singular) {
    if (cond_out) *cond_out = INFINITY;
    free(ipiv);
    free(work);
    return 1;
  }
  /* dtrti2, upper, non-unit: inv(U) in place */
  for (int j = 0; j < n; ++j) {
    a[j * n + j] = 1.0 / a[j * n + j];
    double ajj = -a[j * n + j];
    /* x := inv(U)[0:j,0:j] * U[0:j,j]  (dtrmv upper, no-trans, non-unit) */
    for (int i = 0; i < j; ++i) {
      double t = a[i * n + i] * a[i * n + j];
      for (int l = i + 1; l < j; ++l) t += a[i * n + l] * a[l * n + j];
      a[i * n + j] = t;
    }
    for (int i = 0; i < j; ++i) a[i * n + j] *= ajj;
  }
  /* dgetri unblocked: solve inv(A) * L = inv(U) */
  for (int j = n - 1; j >= 0; --j) {
    for (int i = j + 1; i < n; ++i) {
      work[i] = a[i * n + j];
      a[i * n + j] = 0.0;
    }
    if (j < n - 1)
      for (int i = 0; i < n; ++i) {
        double t = 0.0;
        for (int l = j + 1; l < n; ++l) t += a[i * n + l] * work[l];
        a[i * n + j] += -1.0 * t;
      }
  }
  for (int j = n - 2; j >= 0; --j) {
    int jp = ipiv[j];
    if (jp != j)
      for (int i = 0; i < n; ++i) {
        double t = a[i * n + j];
        a[i * n + j] = a[i * n + jp];
        a[i * n + jp] = t;
      }
  }
  double inorm = 0.0;
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += fabs(a[i * n + j]);
    if (s > inorm) inorm = s;
  }
  double cond = anorm * inorm;
  if (cond_out) *cond_out = cond;
  free(ipiv);
  free(work);
  if (!(cond <= 1e16)) return 2; /* mat64.ConditionTolerance; NaN counts as ill-conditioned */
  return 0;
}

/* ---- Cholesky: dpotf2 on the upper triangle, returned as L = U^T ----------------------------- */

int gko_chol_lower(double* L, const double* A, int n) {
  memset(L, 0, sizeof(double) * (size_t)n * n);
  for (int j = 0; j < n; ++j) {
    double ajj = A[j * n + j];
    for (int l = 0; l < j; ++l) ajj -= L[j * n + l] * L[j * n + l];
    if (!(ajj > 0.0)) return 0;
    ajj = sqrt(ajj);
    L[j * n + j] = ajj;
    for (int i = j + 1; i < n; ++i) {
      double s = 0.0;
      for (int l = 0; l < j; ++l) s += L[j * n + l] * L[i * n + l];
      L[i * n + j] = (A[j * n + i] - s) / ajj; /* upper-triangle entry A[j][i] */
    }
  }
  return 1;
}

/* ---- Householder QR: dgeqr2 / dlarfg / dlarf -------------------------------------------------- */

static double gko_dnrm2(const double* x, int n, int stride) {
  if (n < 1) return 0.0;
  if (n == 1) return fabs(x[0]);
  double scale = 0.0, ssq = 1.0;
  for (int i = 0; i < n; ++i) {
    double v = x[i * stride];
    if (v != 0.0) {
      double a = fabs(v);
      if (scale < a) {
        ssq = 1.0 + ssq * (scale / a) * (scale / a);
        scale = a;
      } else {
        ssq += (a / scale) * (a / scale);
      }
    }
  }
  return scale * sqrt(ssq);
}

static double gko_dlapy2(double x, double y) {
  double xa = fabs(x), ya = fabs(y);
  double w = xa > ya ? xa : ya, z = xa > ya ? ya : xa;
  if (z == 0.0) return w;
  return w * sqrt(1.0 + (z / w) * (z / w));
}

void gko_qr_r(double* R, const double* A, int rows, int cols) {
  double* a = (double*)malloc(sizeof(double) * (size_t)rows * cols);
  double* w = (double*)malloc(sizeof(double) * (size_t)cols);
  memcpy(a, A, sizeof(double) * (size_t)rows * cols);
  int kmax = rows < cols ? rows : cols;
  for (int i = 0; i < kmax; ++i) {
    /* dlarfg(rows - i, a[i][i], a[i+1:, i]) */
    double alpha = a[i * cols + i];
    double tau = 0.0, beta = alpha;
    int nn = rows - i;
    if (nn > 1) {
      double xnorm = gko_dnrm2(&a[(i + 1) * cols + i], nn - 1, cols);
      if (xnorm != 0.0) {
        beta = -copysign(gko_dlapy2(alpha, xnorm), alpha);
        /* (dlarfg's safmin rescaling loop is unreachable for the magnitudes on this path) */
        tau = (beta - alpha) / beta;
        double sc = 1.0 / (alpha - beta);
        for (int r = i + 1; r < rows; ++r) a[r * cols + i] *= sc;
      }
    }
    a[i * cols + i] = beta;
    if (i < cols - 1 && tau != 0.0) {
      /* dlarf(Left): v = [1; a[i+1:, i]], C = a[i:, i+1:]; w = C^T v; C -= tau v w^T */
      for (int j = i + 1; j < cols; ++j) w[j] = 0.0;
      for (int r = i; r < rows; ++r) {
        double vr = (r == i) ? 1.0 : a[r * cols + i];
        for (int j = i + 1; j < cols; ++j) w[j] += vr * a[r * cols + j];
      }
      for (int r = i; r < rows; ++r) {
        double vr = (r == i) ? 1.0 : a[r * cols + i];
        double t = -tau * vr;
        for (int j = i + 1; j < cols; ++j) a[r * cols + j] += t * w[j];
      }
    }
  }
  for (int r = 0; r < rows; ++r)
    for (int j = 0; j < cols; ++j) R[r * cols + j] = (j >= r) ? a[r * cols + j] : 0.0;
  free(a);
  free(w);
}

/* ---- helper.go restatements ------------------------------------------------------------------- */

double gko_sign(double v) { /* helper.go:133-138 */
  if (fabs(v) <= 1e-12) return 1.0;
  return v / fabs(v);
}

void gko_householder_transf(double* A, int n, int m) { /* helper.go:142-172 */
  int rows = n + m, cols = n + 1;
  double* u = (double*)malloc(sizeof(double) * (size_t)rows);
  for (int k = 0; k < n; ++k) {
    double sigma = 0.0;
    for (int i = k; i < rows; ++i) sigma += A[i * cols + k] * A[i * cols + k];
    sigma = sqrt(sigma) * gko_sign(A[k * cols + k]);
    for (int i = 0; i < rows; ++i) u[i] = 0.0;
    u[k] = A[k * cols + k] + sigma;
    A[k * cols + k] = -sigma;
    for (int i = k + 1; i < rows; ++i) u[i] = A[i * cols + k];
    double beta = 1.0 / (sigma * u[k]);
    for (int j = k + 1; j < n + 1; ++j) {
      double gamma = 0.0;
      for (int i = k; i < rows; ++i) gamma += u[i] * A[i * cols + j];
      gamma *= beta;
      for (int i = k; i < rows; ++i) A[i * cols + j] = A[i * cols + j] - gamma * u[i];
      for (int i = k + 1; i < rows; ++i) A[i * cols + k] = 0.0;
    }
  }
  free(u);
}

static int gko_eq_abs_or_rel(double a, double b, double abs_tol, double rel_tol) {
  /* gonum floats.EqualWithinAbsOrRel */
  if (a == b) return 1;
  double d = fabs(a - b);
  if (d <= abs_tol) return 1;
  if (d <= DBL_MIN) return d <= rel_tol * DBL_MIN;
  double mx = fabs(a) > fabs(b) ? fabs(a) : fabs(b);
  return d / mx <= rel_tol;
}

int gko_as_sym(double* A, int n) { /* helper.go:65-84 */
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (i != j && !gko_eq_abs_or_rel(A[j * n + i], A[i * n + j], 1e-6, 1e-2)) return 1;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) A[i * n + j] = A[j * n + i];
  return 0;
}
