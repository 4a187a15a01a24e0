// smallmat.cuh -- register-resident FP64 dense algebra for one-filter-per-thread kernels (sm_100a).
//
// Every routine is a template over compile-time dimensions and is fully unrolled, so that all
// array indices are static and the matrices live in registers (no local memory).  Matrices are
// row-major `double a[R*C]`.  The factorisations follow the same unblocked LAPACK recipes the
// reference reaches through gonum: LU with partial pivoting + explicit inverse for
// mat64.Dense.Inverse, dpotf2 for Cholesky, dgeqr2/dlarfg signs for QR, and the reference's own
// HouseholderTransf (helper.go:142-172).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "fastmath.cuh"

#define GKB_DEV __device__ __forceinline__

namespace gkb {

// ---- packed symmetric (upper triangle, row-major) ----------------------------------------------
template <int N>
__host__ __device__ constexpr int sym_idx(int i, int j) {
  return i <= j ? (i * N - (i * (i - 1)) / 2 + (j - i)) : (j * N - (j * (j - 1)) / 2 + (i - j));
}
template <int N>
struct SymSize { static constexpr int value = N * (N + 1) / 2; };

// C[RxC] = A[RxK] * B[KxC]
template <int R, int K, int C>
GKB_DEV void mul(double (&out)[R * C], const double (&A)[R * K], const double (&B)[K * C]) {
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double s = A[i * K] * B[j];
#pragma unroll
      for (int l = 1; l < K; ++l) s = fma(A[i * K + l], B[l * C + j], s);
      out[i * C + j] = s;
    }
}
// C[RxC] = A[RxK] * B^T, B is [CxK]
template <int R, int K, int C>
GKB_DEV void mul_nt(double (&out)[R * C], const double (&A)[R * K], const double (&B)[C * K]) {
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double s = A[i * K] * B[j * K];
#pragma unroll
      for (int l = 1; l < K; ++l) s = fma(A[i * K + l], B[j * K + l], s);
      out[i * C + j] = s;
    }
}
// C[RxC] = A^T * B, A is [KxR], B is [KxC]
template <int R, int K, int C>
GKB_DEV void mul_tn(double (&out)[R * C], const double (&A)[K * R], const double (&B)[K * C]) {
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double s = A[i] * B[j];
#pragma unroll
      for (int l = 1; l < K; ++l) s = fma(A[l * R + i], B[l * C + j], s);
      out[i * C + j] = s;
    }
}
// y[R] = A[RxC] x[C]
template <int R, int C>
GKB_DEV void mulvec(double (&y)[R], const double (&A)[R * C], const double (&x)[C]) {
#pragma unroll
  for (int i = 0; i < R; ++i) {
    double s = A[i * C] * x[0];
#pragma unroll
    for (int j = 1; j < C; ++j) s = fma(A[i * C + j], x[j], s);
    y[i] = s;
  }
}
// y[C] = A^T x, A is [RxC]
template <int R, int C>
GKB_DEV void mulvec_t(double (&y)[C], const double (&A)[R * C], const double (&x)[R]) {
#pragma unroll
  for (int j = 0; j < C; ++j) {
    double s = A[j] * x[0];
#pragma unroll
    for (int i = 1; i < R; ++i) s = fma(A[i * C + j], x[i], s);
    y[j] = s;
  }
}
// y[N] = P x with P packed symmetric
template <int N>
GKB_DEV void symvec(double (&y)[N], const double (&P)[N * (N + 1) / 2], const double (&x)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = P[sym_idx<N>(i, 0)] * x[0];
#pragma unroll
    for (int j = 1; j < N; ++j) s = fma(P[sym_idx<N>(i, j)], x[j], s);
    y[i] = s;
  }
}
template <int N>
GKB_DEV void sym_expand(double (&full)[N * N], const double (&P)[N * (N + 1) / 2]) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) full[i * N + j] = P[sym_idx<N>(i, j)];
}
template <int N>
GKB_DEV void sym_pack_upper(double (&P)[N * (N + 1) / 2], const double (&full)[N * N]) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) P[sym_idx<N>(i, j)] = full[i * N + j];
}

// ---- explicit inverse by LU with partial pivoting (dgetf2 + dtrti2 + dgetri) --------------------
// Returns 0 ok, 1 exactly singular, 2 cond_inf > 1e16 (output still the computed inverse), like
// mat64.Dense.Inverse (the CPU oracle restates the same recipe).  Row swaps are predicated register
// swaps (static indexing); column swaps at the end likewise.
template <int N>
GKB_DEV int inverse_lu(double (&a)[N * N]) {
  if constexpr (N == 1) {
    if (a[0] == 0.0) return 1;
    a[0] = rcp_nr(a[0]);  // the condition number of a 1 x 1 matrix is 1: only an exact zero is an error
    return 0;
  } else {
    double anorm = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) s += fabs(a[i * N + j]);
      anorm = fmax(anorm, s);
    }
    // Two warp-uniform shortcuts for N >= 3 (the votes are over the lanes that are converged here):
    //  * the input is upper triangular in every lane (the SRIF's R after a measurement update, srif.go:223-235):
    //    dgetf2 would find no pivot to swap and only zero multipliers, so the factorisation and the
    //    dgetri back-multiplication are skipped -- the result is the same dtrti2 inverse, bit for bit;
    //  * no lane needs a row / column interchange: the predicated register swaps (N^2 selects per column)
    //    are skipped.
    constexpr bool kVote = N >= 3;
    bool tri = false;
    if constexpr (kVote) {
      bool lower_zero = true;
#pragma unroll
      for (int i = 1; i < N; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) lower_zero = lower_zero && (a[i * N + j] == 0.0);
      tri = __all_sync(__activemask(), lower_zero);
    }
    int piv[N];
    bool singular = false;
    bool any_swap = false;
    if (tri) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        piv[j] = j;
        singular = singular || (a[j * N + j] == 0.0);
      }
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        int p = j;
        double pmax = fabs(a[j * N + j]);
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
          double v = fabs(a[i * N + j]);
          if (v > pmax) { pmax = v; p = i; }
        }
        piv[j] = p;
        if (pmax != 0.0) {
          bool do_swaps = true;
          if constexpr (kVote) do_swaps = __any_sync(__activemask(), p != j);
          any_swap = any_swap || (p != j);
          if (do_swaps) {
#pragma unroll
            for (int i = j + 1; i < N; ++i) {
              bool sw = (p == i);
#pragma unroll
              for (int l = 0; l < N; ++l) {
                double t0 = a[j * N + l], t1 = a[i * N + l];
                a[j * N + l] = sw ? t1 : t0;
                a[i * N + l] = sw ? t0 : t1;
              }
            }
          }
          double rinv = rcp_nr(a[j * N + j]);
#pragma unroll
          for (int i = j + 1; i < N; ++i) a[i * N + j] *= rinv;
        } else {
          singular = true;
        }
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
          double lij = a[i * N + j];
#pragma unroll
          for (int l = j + 1; l < N; ++l) a[i * N + l] = fma(-lij, a[j * N + l], a[i * N + l]);
        }
      }
    }
    if (singular) return 1;
    // inv(U) in place (dtrti2 upper, non-unit)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      a[j * N + j] = rcp_nr(a[j * N + j]);
      double ajj = -a[j * N + j];
#pragma unroll
      for (int i = 0; i < j; ++i) {
        double t = a[i * N + i] * a[i * N + j];
#pragma unroll
        for (int l = i + 1; l < j; ++l) t = fma(a[i * N + l], a[l * N + j], t);
        a[i * N + j] = t;
      }
#pragma unroll
      for (int i = 0; i < j; ++i) a[i * N + j] *= ajj;
    }
    if (!tri) {
      // inv(A) L = inv(U)  (dgetri, unblocked)
#pragma unroll
      for (int j = N - 2; j >= 0; --j) {
        double work[N];
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
          work[i] = a[i * N + j];
          a[i * N + j] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double t = a[i * N + j + 1] * work[j + 1];
#pragma unroll
          for (int l = j + 2; l < N; ++l) t = fma(a[i * N + l], work[l], t);
          a[i * N + j] -= t;
        }
      }
      bool do_swaps = true;
      if constexpr (kVote) do_swaps = __any_sync(__activemask(), any_swap);
      if (do_swaps) {
#pragma unroll
        for (int j = N - 2; j >= 0; --j) {
#pragma unroll
          for (int jp = j + 1; jp < N; ++jp) {
            bool sw = (piv[j] == jp);
#pragma unroll
            for (int i = 0; i < N; ++i) {
              double t0 = a[i * N + j], t1 = a[i * N + jp];
              a[i * N + j] = sw ? t1 : t0;
              a[i * N + jp] = sw ? t0 : t1;
            }
          }
        }
      }
    }
    double inorm = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) s += fabs(a[i * N + j]);
      inorm = fmax(inorm, s);
    }
    double cond = anorm * inorm;
    return (cond <= 1e16) ? 0 : 2;
  }
}

// ---- the same inverse, speculating that dgetf2 needs no row interchange ---------------------------------------
// Straight-line code (no votes, no predicated swaps): the arithmetic of inverse_lu when every pivot is already on
// the diagonal.  Returns false -- and a meaningless matrix -- when that does not hold for this thread (a larger
// entry below a pivot, a zero pivot, cond_inf > 1e16): the caller then reruns inverse_lu on the original matrix.
// NOCOND = true drops the two infinity norms of the cond <= 1e16 test (72 + 2 N FP64 instructions at N = 6): the inverse
// itself is unchanged bit for bit; only a matrix that is invertible in floating point but has cond > 1e16 is no longer
// reported by THIS routine (the SRIF's speculative epoch uses it on state-transition matrices and checks the epoch's
// results for non-finite values instead; the general epoch and every read-out keep the full test).
template <int N, bool NOCOND = false>
GKB_DEV bool inverse_lu_nopivot(double (&a)[N * N]) {
  bool ok = true;
  double anorm = 0.0;
  if constexpr (!NOCOND) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) s += fabs(a[i * N + j]);
      anorm = fmax(anorm, s);
    }
  }
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const double pmax = fabs(a[j * N + j]);
    ok = ok && (pmax != 0.0);
#pragma unroll
    for (int i = j + 1; i < N; ++i) ok = ok && !(fabs(a[i * N + j]) > pmax);
    const double rinv = rcp_nr(a[j * N + j]);
#pragma unroll
    for (int i = j + 1; i < N; ++i) a[i * N + j] *= rinv;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      const double lij = a[i * N + j];
#pragma unroll
      for (int l = j + 1; l < N; ++l) a[i * N + l] = fma(-lij, a[j * N + l], a[i * N + l]);
    }
  }
#pragma unroll
  for (int j = 0; j < N; ++j) {  // dtrti2
    a[j * N + j] = rcp_nr(a[j * N + j]);
    const double ajj = -a[j * N + j];
#pragma unroll
    for (int i = 0; i < j; ++i) {
      double t = a[i * N + i] * a[i * N + j];
#pragma unroll
      for (int l = i + 1; l < j; ++l) t = fma(a[i * N + l], a[l * N + j], t);
      a[i * N + j] = t;
    }
#pragma unroll
    for (int i = 0; i < j; ++i) a[i * N + j] *= ajj;
  }
#pragma unroll
  for (int j = N - 2; j >= 0; --j) {  // dgetri
    double work[N];
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      work[i] = a[i * N + j];
      a[i * N + j] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double t = a[i * N + j + 1] * work[j + 1];
#pragma unroll
      for (int l = j + 2; l < N; ++l) t = fma(a[i * N + l], work[l], t);
      a[i * N + j] -= t;
    }
  }
  if constexpr (NOCOND) {
    return ok;
  } else {
    double inorm = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) s += fabs(a[i * N + j]);
      inorm = fmax(inorm, s);
    }
    return ok && (anorm * inorm <= 1e16);
  }
}

// inverse_lu with the speculation in front: the straight-line no-interchange inverse on a copy, one vote, and the
// general routine only when some lane of the warp needs an interchange or reports an error.  Same bits, same return
// codes.  For call sites whose matrices are (near) symmetric positive definite -- information matrices, covariances --
// where partial pivoting practically never swaps.
template <int N>
GKB_DEV int inverse_lu_fast(double (&a)[N * N]) {
  if constexpr (N == 1) {
    return inverse_lu<N>(a);
  } else {
    double t[N * N];
#pragma unroll
    for (int i = 0; i < N * N; ++i) t[i] = a[i];
    const bool ok = inverse_lu_nopivot<N>(t);
    if (__all_sync(__activemask(), ok)) {
#pragma unroll
      for (int i = 0; i < N * N; ++i) a[i] = t[i];
      return 0;
    }
    return inverse_lu<N>(a);
  }
}

// ---- lower Cholesky factor from the upper triangle (dpotf2); returns false if not PD -------------
template <int N>
GKB_DEV bool chol_lower(double (&L)[N * N], const double (&A)[N * N]) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < N * N; ++i) L[i] = 0.0;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double ajj = A[j * N + j];
#pragma unroll
    for (int l = 0; l < j; ++l) ajj = fma(-L[j * N + l], L[j * N + l], ajj);
    if (!(ajj > 0.0)) ok = false;
    ajj = sqrt(ajj);  // constructor-time only: keep the IEEE sqrt (NaN for a non-PD input, like dpotf2)
    L[j * N + j] = ajj;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double s = A[j * N + i];
#pragma unroll
      for (int l = 0; l < j; ++l) s = fma(-L[j * N + l], L[i * N + l], s);
      L[i * N + j] = s / ajj;
    }
  }
  return ok;
}

// ---- e^T P^-1 e for packed symmetric positive definite P, by LDL^T (no pivoting) ----------------
// Used for NEES / NIS (chisquare.go:51-58,72-76 use an LU inverse; P and S are SPD with
// cond <= ~1e4 on this path, so the two agree to ~1e-13: SURVEY.md section 7).
template <int N>
GKB_DEV double spd_quadform(const double (&P)[N * (N + 1) / 2], const double (&e)[N], bool& ok) {
  double L[N * N];  // unit lower (strict part used), d on the diagonal slots
  double dinv[N];
  double z[N];
  double q = 0.0;
  ok = true;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double d = P[sym_idx<N>(j, j)];
#pragma unroll
    for (int l = 0; l < j; ++l) d = fma(-L[j * N + l] * L[j * N + l], L[l * N + l], d);
    L[j * N + j] = d;
    if (!(d > 0.0)) ok = false;
    dinv[j] = rcp_nr(d);
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double s = P[sym_idx<N>(j, i)];
#pragma unroll
      for (int l = 0; l < j; ++l) s = fma(-L[i * N + l] * L[l * N + l], L[j * N + l], s);
      L[i * N + j] = s * dinv[j];
    }
    // forward substitution row j: z_j = e_j - sum_{l<j} L[j][l] z_l
    double zj = e[j];
#pragma unroll
    for (int l = 0; l < j; ++l) zj = fma(-L[j * N + l], z[l], zj);
    z[j] = zj;
    q = fma(zj * zj, dinv[j], q);
  }
  return q;
}

// ---- R factor of the Householder QR (dgeqr2 / dlarfg / dlarf), in place --------------------------
// On return the upper triangle of a[ROWS x COLS] holds R; entries below the diagonal are garbage
// (the Householder vectors) and must be ignored by the caller.
template <int ROWS, int COLS>
GKB_DEV void qr_r_inplace(double (&a)[ROWS * COLS]) {
  constexpr int KMAX = ROWS < COLS ? ROWS : COLS;
#pragma unroll
  for (int i = 0; i < KMAX; ++i) {
    if (ROWS - i > 1) {
      double alpha = a[i * COLS + i];
      // dnrm2 of the sub-column (plain sum of squares: magnitudes on this path are far from the
      // overflow/underflow range the scaled LAPACK loop protects against)
      double ss = 0.0;
#pragma unroll
      for (int r = i + 1; r < ROWS; ++r) ss = fma(a[r * COLS + i], a[r * COLS + i], ss);
      if (ss != 0.0) {
        double nrm = sqrt_nr(fma(alpha, alpha, ss));
        double beta = -copysign(nrm, alpha);
        double tau = (beta - alpha) * rcp_nr(beta);
        double sc = rcp_nr(alpha - beta);
#pragma unroll
        for (int r = i + 1; r < ROWS; ++r) a[r * COLS + i] *= sc;
        a[i * COLS + i] = beta;
#pragma unroll
        for (int j = i + 1; j < COLS; ++j) {
          double w = a[i * COLS + j];
#pragma unroll
          for (int r = i + 1; r < ROWS; ++r) w = fma(a[r * COLS + i], a[r * COLS + j], w);
          double t = -tau * w;
          a[i * COLS + j] += t;
#pragma unroll
          for (int r = i + 1; r < ROWS; ++r) a[r * COLS + j] = fma(t, a[r * COLS + i], a[r * COLS + j]);
        }
      }
    }
  }
}

// ---- helper.go:133-172 ---------------------------------------------------------------------------
GKB_DEV double ref_sign(double v) { return (fabs(v) <= 1e-12) ? 1.0 : copysign(1.0, v); }

// HouseholderTransf on A[(N+M) x (N+1)], in place.
template <int N, int M>
GKB_DEV void householder_transf(double (&A)[(N + M) * (N + 1)]) {
  constexpr int ROWS = N + M, COLS = N + 1;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double sigma = 0.0;
#pragma unroll
    for (int i = k; i < ROWS; ++i) sigma = fma(A[i * COLS + k], A[i * COLS + k], sigma);
    sigma = sqrt_nr(sigma) * ref_sign(A[k * COLS + k]);
    double uk = A[k * COLS + k] + sigma;
    A[k * COLS + k] = -sigma;
    double beta = rcp_nr(sigma * uk);
#pragma unroll
    for (int j = k + 1; j < COLS; ++j) {
      double gamma = uk * A[k * COLS + j];
#pragma unroll
      for (int i = k + 1; i < ROWS; ++i) gamma = fma(A[i * COLS + k], A[i * COLS + j], gamma);
      gamma *= beta;
      A[k * COLS + j] = fma(-gamma, uk, A[k * COLS + j]);
#pragma unroll
      for (int i = k + 1; i < ROWS; ++i) A[i * COLS + j] = fma(-gamma, A[i * COLS + k], A[i * COLS + j]);
    }
#pragma unroll
    for (int i = k + 1; i < ROWS; ++i) A[i * COLS + k] = 0.0;
  }
}

}  // namespace gkb
