/* gko_c2d.c -- CPU restatement of c2d.go:13-75 VanLoan.  TEST INFRASTRUCTURE ONLY.
 *
 * c2d.go builds M = [[-A dt, (Gamma W) Gamma^T dt], [0, A^T dt]] (31-54), takes expM.Exp(M) (57-58), reads
 * F = (E_22)^T and F1Q = E_12 (62-70) and returns F, AsSymDense(F F1Q) (71-74).  The exponential lives in gonum
 * (github.com/gonum/matrix/mat64 Dense.Exp, un-vendored, un-pinned): it is restated here from its published
 * algorithm -- N. J. Higham, "The scaling and squaring method for the matrix exponential revisited" (SIAM J. Matrix
 * Anal. Appl. 26(4), 2005), Algorithm 2.3: Pade approximants of degree 3, 5, 7, 9 or 13 chosen by the 1-norm against
 * theta_m, scaling by 2^-s for degree 13, LU solve of (V - U) X = V + U, s squarings.  Pinned by the reference's own
 * known-answer test (c2d_test.go:9-33, tests/test_host_c2d.py) and by scipy.linalg.expm (tests). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gko.h"
#include "gko_linalg.h"

static const double pade3[] = {120., 60., 12., 1.};
static const double pade5[] = {30240., 15120., 3360., 420., 30., 1.};
static const double pade7[] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
static const double pade9[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
static const double pade13[] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                                129060195264000., 10559470521600., 670442572800., 33522128640., 1323241920.,
                                40840800., 960960., 16380., 182., 1.};
static const double theta[] = {1.495585217958292e-2, 2.539398330063230e-1, 9.504178996162932e-1, 2.097847961257068e0,
                               5.371920351148152e0};

/* X = inv(P) Q by LU with partial pivoting (dgesv); P and Q are overwritten. Returns 0 or 1 (singular). */
static int lu_solve(double* P, double* Q, int d) {
  for (int j = 0; j < d; ++j) {
    int p = j;
    double pm = fabs(P[j * d + j]);
    for (int i = j + 1; i < d; ++i)
      if (fabs(P[i * d + j]) > pm) { pm = fabs(P[i * d + j]); p = i; }
    if (pm == 0.0) return 1;
    if (p != j)
      for (int l = 0; l < d; ++l) {
        double t = P[j * d + l]; P[j * d + l] = P[p * d + l]; P[p * d + l] = t;
        t = Q[j * d + l]; Q[j * d + l] = Q[p * d + l]; Q[p * d + l] = t;
      }
    for (int i = j + 1; i < d; ++i) {
      double l_ij = P[i * d + j] / P[j * d + j];
      for (int l = j + 1; l < d; ++l) P[i * d + l] -= l_ij * P[j * d + l];
      for (int l = 0; l < d; ++l) Q[i * d + l] -= l_ij * Q[j * d + l];
    }
  }
  for (int i = d - 1; i >= 0; --i)
    for (int l = 0; l < d; ++l) {
      double s = Q[i * d + l];
      for (int k = i + 1; k < d; ++k) s -= P[i * d + k] * Q[k * d + l];
      Q[i * d + l] = s / P[i * d + i];
    }
  return 0;
}

int gko_expm(double* E, const double* Ain, int d) {
  size_t sz = (size_t)d * d;
  double* A = (double*)malloc(sizeof(double) * sz * 8);
  double *A2 = A + sz, *A4 = A2 + sz, *A6 = A4 + sz, *U = A6 + sz, *V = U + sz, *T = V + sz, *W = T + sz;
  memcpy(A, Ain, sizeof(double) * sz);
  double nrm = 0.0;
  for (int j = 0; j < d; ++j) {
    double s = 0.0;
    for (int i = 0; i < d; ++i) s += fabs(A[i * d + j]);
    if (s > nrm) nrm = s;
  }
  int rc = 0, done = 0;
  const int degs[4] = {3, 5, 7, 9};
  const double* tabs[4] = {pade3, pade5, pade7, pade9};
  for (int t = 0; t < 4 && !done; ++t) {
    if (nrm <= theta[t]) {
      int m = degs[t];
      const double* b = tabs[t];
      gko_mul(A2, A, A, d, d, d);
      for (size_t i = 0; i < sz; ++i) { T[i] = ((int)(i / d) == (int)(i % d)) ? 1.0 : 0.0; W[i] = 0.0; V[i] = 0.0; } /* T = A^(2k) */
      for (int k = 0; 2 * k + 1 <= m; ++k) {
        for (size_t i = 0; i < sz; ++i) { W[i] += b[2 * k + 1] * T[i]; V[i] += b[2 * k] * T[i]; }
        if (2 * k + 3 <= m) { gko_mul(A4, T, A2, d, d, d); memcpy(T, A4, sizeof(double) * sz); }
      }
      gko_mul(U, A, W, d, d, d);
      for (size_t i = 0; i < sz; ++i) { T[i] = V[i] - U[i]; E[i] = V[i] + U[i]; }
      rc = lu_solve(T, E, d);
      done = 1;
    }
  }
  if (!done) {
    int s = 0;
    if (nrm > theta[4]) {
      s = (int)ceil(log2(nrm / theta[4]));
      if (s < 0) s = 0;
      for (size_t i = 0; i < sz; ++i) A[i] = ldexp(A[i], -s);
    }
    const double* b = pade13;
    gko_mul(A2, A, A, d, d, d);
    gko_mul(A4, A2, A2, d, d, d);
    gko_mul(A6, A4, A2, d, d, d);
    for (size_t i = 0; i < sz; ++i) T[i] = b[13] * A6[i] + b[11] * A4[i] + b[9] * A2[i];
    gko_mul(W, A6, T, d, d, d);
    for (size_t i = 0; i < sz; ++i)
      W[i] += b[7] * A6[i] + b[5] * A4[i] + b[3] * A2[i] + (((int)(i / d) == (int)(i % d)) ? b[1] : 0.0);
    gko_mul(U, A, W, d, d, d);
    for (size_t i = 0; i < sz; ++i) T[i] = b[12] * A6[i] + b[10] * A4[i] + b[8] * A2[i];
    gko_mul(V, A6, T, d, d, d);
    for (size_t i = 0; i < sz; ++i)
      V[i] += b[6] * A6[i] + b[4] * A4[i] + b[2] * A2[i] + (((int)(i / d) == (int)(i % d)) ? b[0] : 0.0);
    for (size_t i = 0; i < sz; ++i) { T[i] = V[i] - U[i]; E[i] = V[i] + U[i]; }
    rc = lu_solve(T, E, d);
    for (int k = 0; k < s && rc == 0; ++k) {
      gko_mul(T, E, E, d, d, d);
      memcpy(E, T, sizeof(double) * sz);
    }
  }
  free(A);
  return rc;
}

/* c2d.go:13-75 without the eigenvalue (Nyquist) warning, which does not change the outputs. A[n*n], Gamma[n*q],
 * W[q*q] -> F[n*n], Q[n*n]. */
int gko_van_loan(int n, int q, const double* A, const double* Gamma, const double* W, double dt, double* F, double* Q) {
  int d = 2 * n;
  double* GW = (double*)calloc((size_t)n * q + (size_t)n * n * 3 + (size_t)d * d * 2, sizeof(double));
  double *GWG = GW + (size_t)n * q, *F1Q = GWG + (size_t)n * n, *Qd = F1Q + (size_t)n * n, *M = Qd + (size_t)n * n, *E = M + (size_t)d * d;
  gko_mul(GW, Gamma, W, n, q, q);
  gko_mul_nt(GWG, GW, Gamma, n, q, n);
  for (int i = 0; i < n * n; ++i) GWG[i] = dt * GWG[i];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      M[i * d + j] = -(dt * A[i * n + j]);
      M[(i + n) * d + (j + n)] = dt * A[j * n + i];
      M[i * d + (j + n)] = GWG[i * n + j];
    }
  int rc = gko_expm(E, M, d);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      F1Q[i * n + j] = E[i * d + (n + j)];
      F[i * n + j] = E[(n + j) * d + (n + i)]; /* F = (E_22)^T */
    }
  gko_mul(Qd, F, F1Q, n, n, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Q[i * n + j] = (j >= i) ? Qd[i * n + j] : Qd[j * n + i]; /* AsSymDense keeps the upper triangle */
  free(GW);
  return rc;
}
