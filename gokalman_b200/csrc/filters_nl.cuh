// filters_nl.cuh -- per-thread non-linear (NLDKF) filter steps: HybridKF (CKF/EKF) and SRIF.
#pragma once
#include "filters.cuh"

namespace gkb {

template <int N, int M>
struct NlOut {
  double innov[M];
  double obsdev[M];
  double Ppred[N * (N + 1) / 2];  // hybrid: P-bar (packed)
  double K[N * M];
  double Rbar[N * N];             // SRIF: predicted sqrt-information matrix
};

// hybrid.go:104-204 fullUpdate.  x, P: previous estimate in, new one out.
//   has_meas = false -> Predict() (125-143);  ekf -> EKF branch (159-161);  snc -> PreparePNT(Gamma) was
//   called for this epoch (114-123).  Phi / Htilde / Gamma are this filter's matrices for the epoch.
template <int N, int M>
GKB_DEV int hybrid_step(const NlModel<N, M>& md, double (&x)[N], double (&P)[N * (N + 1) / 2],
                        const double (&Phi)[N * N], const double (&Ht)[M * N], const double (&real_obs)[M],
                        const double (&computed_obs)[M], const double* __restrict__ Gamma, bool has_meas,
                        bool ekf, bool snc, NlOut<N, M>& o) {
  constexpr int SN = N * (N + 1) / 2;
  // x-bar = Phi x (CKF); the EKF prediction is the zero vector (hybrid.go:127-131)
  double xbar[N];
  mulvec<N, N>(xbar, Phi, x);
  // 114-117: P-bar = (Phi P) Phi^T, row by row, upper triangle
  double Pb[SN];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double row[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = Phi[i * N] * P[sym_idx<N>(0, j)];
#pragma unroll
      for (int l = 1; l < N; ++l) s = fma(Phi[i * N + l], P[sym_idx<N>(l, j)], s);
      row[j] = s;
    }
#pragma unroll
    for (int j = i; j < N; ++j) {
      double s = row[0] * Phi[j * N];
#pragma unroll
      for (int l = 1; l < N; ++l) s = fma(row[l], Phi[j * N + l], s);
      Pb[sym_idx<N>(i, j)] = s;
    }
  }
  if (snc && Gamma != nullptr) {  // 118-123: + (Gamma Q) Gamma^T
    const int q = md.q;
    double Gm[N * GKB_MAX_Q];  // Gamma is shared by all filters: uniform loads, q <= GKB_MAX_Q
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int a = 0; a < GKB_MAX_Q; ++a) Gm[i * GKB_MAX_Q + a] = (a < q) ? __ldg(Gamma + i * q + a) : 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double gq[GKB_MAX_Q];
#pragma unroll
      for (int a = 0; a < GKB_MAX_Q; ++a) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < GKB_MAX_Q; ++b)
          if (a < q && b < q) s = fma(Gm[i * GKB_MAX_Q + b], md.Q[b * q + a], s);
        gq[a] = s;
      }
#pragma unroll
      for (int j = i; j < N; ++j) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < GKB_MAX_Q; ++a) s = fma(gq[a], Gm[j * GKB_MAX_Q + a], s);
        Pb[sym_idx<N>(i, j)] += s;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < SN; ++i) o.Ppred[i] = Pb[i];
  if (!has_meas) {  // Predict(): 125-143
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = ekf ? 0.0 : xbar[i];
#pragma unroll
    for (int i = 0; i < SN; ++i) P[i] = Pb[i];
#pragma unroll
    for (int a = 0; a < M; ++a) { o.innov[a] = 0.0; o.obsdev[a] = 0.0; }
#pragma unroll
    for (int i = 0; i < N * M; ++i) o.K[i] = 0.0;
    return 0;
  }
  // 146-153: K = P-bar Ht^T inv(Ht P-bar Ht^T + R)
  double PHt[N * M];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = Pb[sym_idx<N>(i, 0)] * Ht[a * N];
#pragma unroll
      for (int j = 1; j < N; ++j) s = fma(Pb[sym_idx<N>(i, j)], Ht[a * N + j], s);
      PHt[i * M + a] = s;
    }
  double S[M * M];
#pragma unroll
  for (int a = 0; a < M; ++a)
#pragma unroll
    for (int b = 0; b < M; ++b) {
      double s = md.R[a * M + b];
#pragma unroll
      for (int i = 0; i < N; ++i) s = fma(Ht[a * N + i], PHt[i * M + b], s);
      S[a * M + b] = s;
    }
  if (inverse_lu<M>(S) != 0) return GKB_ERR_SINGULAR_S;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = PHt[i * M] * S[a];
#pragma unroll
      for (int b = 1; b < M; ++b) s = fma(PHt[i * M + b], S[b * M + a], s);
      o.K[i * M + a] = s;
    }
  // 156-173
  double y[M];
#pragma unroll
  for (int a = 0; a < M; ++a) {
    y[a] = real_obs[a] - computed_obs[a];
    o.obsdev[a] = y[a];
  }
  double xhat[N];
  if (ekf) {
#pragma unroll
    for (int a = 0; a < M; ++a) o.innov[a] = 0.0;  // left as an empty vector by the reference (159-161)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = o.K[i * M] * y[0];
#pragma unroll
      for (int a = 1; a < M; ++a) s = fma(o.K[i * M + a], y[a], s);
      xhat[i] = s;
    }
  } else {
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = Ht[a * N] * xbar[0];
#pragma unroll
      for (int j = 1; j < N; ++j) s = fma(Ht[a * N + j], xbar[j], s);
      o.innov[a] = y[a] - s;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = o.K[i * M] * o.innov[0];
#pragma unroll
      for (int a = 1; a < M; ++a) s = fma(o.K[i * M + a], o.innov[a], s);
      xhat[i] = xbar[i] + s;
    }
  }
  // 174-182: Joseph form, restructured as in vanilla_step
  double Pp[SN];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double T[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = Pb[sym_idx<N>(i, j)];
#pragma unroll
      for (int a = 0; a < M; ++a) s = fma(-o.K[i * M + a], PHt[j * M + a], s);
      T[j] = s;
    }
    double V[M];
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = T[0] * Ht[a * N];
#pragma unroll
      for (int j = 1; j < N; ++j) s = fma(T[j], Ht[a * N + j], s);
#pragma unroll
      for (int b = 0; b < M; ++b) s = fma(-o.K[i * M + b], md.R[b * M + a], s);
      V[a] = s;
    }
#pragma unroll
    for (int j = i; j < N; ++j) {
      double s = T[j];
#pragma unroll
      for (int a = 0; a < M; ++a) s = fma(-V[a], o.K[j * M + a], s);
      Pp[sym_idx<N>(i, j)] = s;
    }
  }
  bool finite = true;
#pragma unroll
  for (int i = 0; i < N; ++i) finite = finite && isfinite(xhat[i]) && isfinite(Pp[sym_idx<N>(i, i)]);
  if (!finite) return GKB_ERR_NONFINITE;
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = xhat[i];
#pragma unroll
  for (int i = 0; i < SN; ++i) P[i] = Pp[i];
  return 0;
}

// Inverse of an upper-triangular matrix held packed (row-major upper triangle): the same arithmetic as the
// triangular shortcut of inverse_lu (dtrti2 + the ||A|| ||inv(A)|| <= 1e16 test), on 21 instead of 36 registers
// at n = 6.  Returns 0 ok, 1 singular (a zero on the diagonal), 2 ill-conditioned.
// STRAIGHT: no early return on a zero diagonal (the arithmetic then runs on inf / NaN and only the return code counts).
// NOCOND: the cond <= 1e16 test is dropped (same inverse, bit for bit; see inverse_lu_nopivot).
template <int N, bool STRAIGHT = false, bool NOCOND = false>
GKB_DEV int inverse_upper_packed(double (&u)[N * (N + 1) / 2]) {
  double anorm = 0.0;
  bool singular = false;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if constexpr (!NOCOND) {
      double s = 0.0;
#pragma unroll
      for (int j = i; j < N; ++j) s += fabs(u[sym_idx<N>(i, j)]);
      anorm = fmax(anorm, s);
    }
    singular = singular || (u[sym_idx<N>(i, i)] == 0.0);
  }
  if constexpr (!STRAIGHT) {
    if (singular) return 1;
  }
#pragma unroll
  for (int j = 0; j < N; ++j) {
    u[sym_idx<N>(j, j)] = rcp_nr(u[sym_idx<N>(j, j)]);
    const double ajj = -u[sym_idx<N>(j, j)];
#pragma unroll
    for (int i = 0; i < j; ++i) {
      double t = u[sym_idx<N>(i, i)] * u[sym_idx<N>(i, j)];
#pragma unroll
      for (int l = i + 1; l < j; ++l) t = fma(u[sym_idx<N>(i, l)], u[sym_idx<N>(l, j)], t);
      u[sym_idx<N>(i, j)] = t;
    }
#pragma unroll
    for (int i = 0; i < j; ++i) u[sym_idx<N>(i, j)] *= ajj;
  }
  if (STRAIGHT && singular) return 1;
  if constexpr (NOCOND) {
    return 0;
  } else {
    double inorm = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = i; j < N; ++j) s += fabs(u[sym_idx<N>(i, j)]);
      inorm = fmax(inorm, s);
    }
    return (anorm * inorm <= 1e16) ? 0 : 2;
  }
}

// srif.go:223-235 State() = inv(R) b.  Returns false where the reference panics (singular R).
// After a measurement update R is upper triangular (HouseholderTransf zeroes the sub-diagonal exactly): when
// that holds in every lane (warp vote) the inverse is taken on the packed triangle -- bit-identical to the
// general LU path, which would find nothing to eliminate.
template <int N>
GKB_DEV bool srif_state(double (&xs)[N], const double (&R)[N * N], const double (&b)[N]) {
  if constexpr (N >= 3) {
    bool lower_zero = true;
#pragma unroll
    for (int i = 1; i < N; ++i)
#pragma unroll
      for (int j = 0; j < i; ++j) lower_zero = lower_zero && (R[i * N + j] == 0.0);
    if (__all_sync(__activemask(), lower_zero)) {
      double U[N * (N + 1) / 2];
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i; j < N; ++j) U[sym_idx<N>(i, j)] = R[i * N + j];
      if (inverse_upper_packed<N>(U) != 0) return false;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double s = U[sym_idx<N>(i, i)] * b[i];
#pragma unroll
        for (int j = i + 1; j < N; ++j) s = fma(U[sym_idx<N>(i, j)], b[j], s);
        xs[i] = s;
      }
      return true;
    }
  }
  double T[N * N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) T[i] = R[i];
  if (inverse_lu<N>(T) != 0) return false;
  mulvec<N, N>(xs, T, b);
  return true;
}

// srif.go:253-265 Covariance() = AsSymDense(inv(R) inv(R)^T); zeros when R is not invertible.
template <int N>
GKB_DEV bool srif_covariance(double (&Pc)[N * N], const double (&R)[N * N]) {
  double T[N * N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) T[i] = R[i];
  if (inverse_lu<N>(T) != 0) {
#pragma unroll
    for (int i = 0; i < N * N; ++i) Pc[i] = 0.0;
    return false;
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) {
      double s = T[i * N] * T[j * N];
#pragma unroll
      for (int l = 1; l < N; ++l) s = fma(T[i * N + l], T[j * N + l], s);
      Pc[i * N + j] = s;
      Pc[j * N + i] = s;
    }
  return true;
}

// srif.go:101-160 fullUpdate.  b, R: previous sqrt-information state / matrix in, new ones out.
template <int N, int M>
GKB_DEV int srif_step(const NlModel<N, M>& md, double (&b)[N], double (&R)[N * N], double (&Phi)[N * N],
                      const double (&Ht)[M * N], const double (&real_obs)[M], const double (&computed_obs)[M],
                      bool has_meas, NlOut<N, M>& o) {
  // 117-118: x-bar = Phi State(prev)
  double xbar[N];
  {
    double xs[N];
    if (!srif_state<N>(xs, R, b)) return GKB_ERR_SINGULAR_R;
    mulvec<N, N>(xbar, Phi, xs);
  }
  // 110-115: R-bar = R inv(Phi)   (Phi is overwritten by its inverse)
  if (inverse_lu<N>(Phi) != 0) return GKB_ERR_SINGULAR_PHI;
  double bbar[N];
  // When R is upper triangular in every lane (the usual case: srif_state's vote), the products with its zero
  // sub-diagonal are skipped: adding +-0 never changes a finite sum, so the result is the same.
  bool lower_zero = true;
#pragma unroll
  for (int i = 1; i < N; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) lower_zero = lower_zero && (R[i * N + j] == 0.0);
  if (N >= 3 && __all_sync(__activemask(), lower_zero)) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = R[i * N + i] * Phi[i * N + j];
#pragma unroll
        for (int l = i + 1; l < N; ++l) s = fma(R[i * N + l], Phi[l * N + j], s);
        o.Rbar[i * N + j] = s;
      }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double row[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = R[i * N] * Phi[j];
#pragma unroll
        for (int l = 1; l < N; ++l) s = fma(R[i * N + l], Phi[l * N + j], s);
        row[j] = s;
      }
#pragma unroll
      for (int j = 0; j < N; ++j) o.Rbar[i * N + j] = row[j];
    }
  }
  // 119: b-bar = R-bar x-bar.  121-132: the "triangularise" branch copies values unchanged.
  mulvec<N, N>(bbar, o.Rbar, xbar);
  if (!has_meas) {  // Predict(): 134-141
#pragma unroll
    for (int i = 0; i < N; ++i) b[i] = bbar[i];
#pragma unroll
    for (int i = 0; i < N * N; ++i) R[i] = o.Rbar[i];
#pragma unroll
    for (int a = 0; a < M; ++a) { o.innov[a] = 0.0; o.obsdev[a] = 0.0; }
    return 0;
  }
  // 143-148: y = real - computed, whitened with L = chol(R_meas) (not its inverse: reference quirk)
  double y[M], yw[M];
#pragma unroll
  for (int a = 0; a < M; ++a) y[a] = real_obs[a] - computed_obs[a];
#pragma unroll
  for (int a = 0; a < M; ++a) {
    double s = md.L[a * M] * y[0];
#pragma unroll
    for (int c2 = 1; c2 < M; ++c2) s = fma(md.L[a * M + c2], y[c2], s);
    yw[a] = s;
    o.obsdev[a] = s;  // Delta-obs is the whitened deviation (srif.go:154)
    o.innov[a] = 0.0;
  }
  // 150, 298-340: A = [[R-bar, b-bar],[L Ht, yw]] -> HouseholderTransf
  constexpr int COLS = N + 1;
  double A[(N + M) * COLS];
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < N; ++j) A[i * COLS + j] = o.Rbar[i * N + j];
    A[i * COLS + N] = bbar[i];
  }
#pragma unroll
  for (int a = 0; a < M; ++a) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = md.L[a * M] * Ht[j];
#pragma unroll
      for (int c2 = 1; c2 < M; ++c2) s = fma(md.L[a * M + c2], Ht[c2 * N + j], s);
      A[(N + a) * COLS + j] = s;
    }
    A[(N + a) * COLS + N] = yw[a];
  }
  householder_transf<N, M>(A);
  bool finite = true;  // checked before R / b are committed: a failed Update leaves the previous estimate in place
#pragma unroll
  for (int i = 0; i < N; ++i) finite = finite && isfinite(A[i * COLS + N]);
  if (!finite) return GKB_ERR_NONFINITE;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < N; ++j) R[i * N + j] = A[i * COLS + j];
    b[i] = A[i * COLS + N];
  }
  return 0;
}

// ---- the SRIF epoch for the usual case, as straight-line code ----------------------------------------------
// Preconditions the caller has voted on: a measurement epoch, and R upper triangular in every lane of the warp
// (HouseholderTransf leaves exact zeros below the diagonal), held packed in U.  The epoch runs the same arithmetic
// as srif_step in the same order -- the triangular State(), the dgetf2 / dtrti2 / dgetri inverse of Phi, R-bar,
// b-bar, the whitened Householder update -- but SPECULATES that no lane needs a row interchange in the LU of Phi
// and that no lane hits an error; every branch and vote of the general routine collapses into one flag and one
// vote at the end.  When the vote fails nothing has been committed: the caller runs srif_step on the same inputs,
// which are still in the shared-memory stage.  `col` is this lane's column of the stage (Phi, Htilde, real,
// computed rows, STRIDE doubles apart); Htilde and the observations are read only when they are needed, which
// keeps them out of the registers during the two inverses.
template <int N, int M, int STRIDE>
GKB_DEV bool srif_step_tri(const NlModel<N, M>& md, double (&b)[N], double (&U)[N * (N + 1) / 2], const double* col,
                           bool pad_lane) {
  constexpr int SN = N * (N + 1) / 2, ROWS_PHI = N * N, ROWS_H = M * N;
  bool ok = true;
  double a[N * N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) a[i] = col[i * STRIDE];
  if (pad_lane) {  // the TMA zero-fills the columns past the last filter: keep their Phi invertible
#pragma unroll
    for (int i = 0; i < N; ++i) a[i * N + i] = 1.0;
  }
  // 117-119: the reference forms x-bar = Phi (inv(R) b) and then b-bar = R-bar x-bar = (R inv(Phi)) (Phi inv(R) b), which is
  // b itself up to rounding (cond(R) eps).  The production epoch takes b-bar = b -- 143 of the epoch's ~1010 FP64
  // instructions (the triangular inverse of R, two matrix-vector products with Phi and R-bar) for a result that is
  // closer to the exact one; the literal sequence stays in srif_step (general kernel, GKB_NL_PATH=plain), and the
  // difference is measured over all filters of the full-size run (tests/test_gpu_fullsize.py: <= 1e-10).  State() must
  // still exist (srif.go:228-230 panics on a singular R): an exactly zero diagonal entry sends the warp to srif_step.
#pragma unroll
  for (int i = 0; i < N; ++i) ok = ok && (U[sym_idx<N>(i, i)] != 0.0);
  // 110-115: inv(Phi) as inverse_lu computes it when no interchange is needed
  const bool inv_ok = inverse_lu_nopivot<N, true>(a);
  ok = ok && inv_ok;
  // R-bar = R inv(Phi) (R triangular), b-bar = R-bar x-bar, straight into the Householder work matrix
  constexpr int COLS = N + 1;
  double A[(N + M) * COLS];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = U[sym_idx<N>(i, i)] * a[i * N + j];
#pragma unroll
      for (int l = i + 1; l < N; ++l) s = fma(U[sym_idx<N>(i, l)], a[l * N + j], s);
      A[i * COLS + j] = s;
    }
#pragma unroll
  for (int i = 0; i < N; ++i) A[i * COLS + N] = b[i];
  // 143-150: whitened observation rows (L = chol(R_meas), not its inverse: reference quirk)
  {
    double y[M];
#pragma unroll
    for (int c2 = 0; c2 < M; ++c2)
      y[c2] = col[(ROWS_PHI + ROWS_H + c2) * STRIDE] - col[(ROWS_PHI + ROWS_H + M + c2) * STRIDE];
#pragma unroll
    for (int r = 0; r < M; ++r) {
      double s = md.L[r * M] * y[0];
#pragma unroll
      for (int c2 = 1; c2 < M; ++c2) s = fma(md.L[r * M + c2], y[c2], s);
      A[(N + r) * COLS + N] = s;
    }
    double Ht[M * N];
#pragma unroll
    for (int i = 0; i < M * N; ++i) Ht[i] = col[(ROWS_PHI + i) * STRIDE];
#pragma unroll
    for (int r = 0; r < M; ++r)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = md.L[r * M] * Ht[j];
#pragma unroll
        for (int c2 = 1; c2 < M; ++c2) s = fma(md.L[r * M + c2], Ht[c2 * N + j], s);
        A[(N + r) * COLS + j] = s;
      }
  }
  householder_transf<N, M>(A);
#pragma unroll
  for (int i = 0; i < N; ++i) ok = ok && isfinite(A[i * COLS + N]) && isfinite(A[i * COLS + i]);
  if (!__all_sync(0xffffffffu, ok)) return false;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = i; j < N; ++j) U[sym_idx<N>(i, j)] = A[i * COLS + j];
    b[i] = A[i * COLS + N];
  }
  return true;
}

}  // namespace gkb
