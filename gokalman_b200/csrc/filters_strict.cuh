// filters_strict.cuh -- the REFERENCE-ORDER ("strict") evaluation of the filter steps.
//
// The production steps (filters.cuh, filters_nl.cuh) keep covariances packed, contract a*b+c into FMAs and
// evaluate the Joseph update with the products by the identity removed.  On well-conditioned problems that agrees
// with the reference to ~1e-14; on the statOD configuration (R = 1e-6 against P0 = 10) forming (I - K H) P-bar
// cancels to eps |P-bar| absolute and any change of rounding moves the result by eps |P-bar| / |P+|.  This file is
// the same step written exactly as the reference executes it -- every product a full dense product in the written
// order (hybrid.go:114-182), every a*b+c rounded twice (__dmul_rn / __dadd_rn: Go on amd64 never fuses), IEEE
// divisions, the explicit LU inverse of mat64.Dense.Inverse, the dense Joseph form, AsSymDense at the end -- so that
// the GPU can be held to the north star's 1e-10 on EVERY configuration (it lands at ~1e-15: the only differences
// left are none), and so that the fast production kernels can be measured against it on all 10^5 filters of a run
// instead of on the handful the CPU oracle can replay.  Selected per handle with gkb_set_strict().
#pragma once
#include "filters_nl.cuh"

namespace gkb {
namespace strict {

GKB_DEV double mul2(double a, double b) { return __dmul_rn(a, b); }
GKB_DEV double add2(double a, double b) { return __dadd_rn(a, b); }
GKB_DEV double div2(double a, double b) { return __ddiv_rn(a, b); }

// C[R x C] = A[R x K] B[K x C]: gonum dgemm, sequential sum over the inner index starting from 0
template <int R, int K, int C>
GKB_DEV void mul(double (&out)[R * C], const double (&A)[R * K], const double (&B)[K * C]) {
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < K; ++l) s = add2(s, mul2(A[i * K + l], B[l * C + j]));
      out[i * C + j] = s;
    }
}
// C[R x C] = A[R x K] B^T, B is [C x K]
template <int R, int K, int C>
GKB_DEV void mul_nt(double (&out)[R * C], const double (&A)[R * K], const double (&B)[C * K]) {
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < K; ++l) s = add2(s, mul2(A[i * K + l], B[j * K + l]));
      out[i * C + j] = s;
    }
}
template <int R, int C>
GKB_DEV void mulvec(double (&y)[R], const double (&A)[R * C], const double (&x)[C]) {
#pragma unroll
  for (int i = 0; i < R; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < C; ++j) s = add2(s, mul2(A[i * C + j], x[j]));
    y[i] = s;
  }
}

// mat64.Dense.Inverse: dgetf2 (partial pivoting, first index of the largest |.|) + dtrti2 + dgetri, then the
// ||A||inf ||inv(A)||inf <= 1e16 test.  Returns 0 ok, 1 exactly singular, 2 ill-conditioned (output = the inverse).
template <int N>
GKB_DEV int inverse(double (&a)[N * N]) {
  double anorm = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) s = add2(s, fabs(a[i * N + j]));
    if (s > anorm) anorm = s;
  }
  int piv[N];
  bool singular = false;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    int p = j;
    double pmax = fabs(a[j * N + j]);
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      const double v = fabs(a[i * N + j]);
      if (v > pmax) { pmax = v; p = i; }
    }
    piv[j] = p;
    if (pmax != 0.0) {
#pragma unroll
      for (int i = j + 1; i < N; ++i) {
        const bool sw = (p == i);
#pragma unroll
        for (int l = 0; l < N; ++l) {
          const double t0 = a[j * N + l], t1 = a[i * N + l];
          a[j * N + l] = sw ? t1 : t0;
          a[i * N + l] = sw ? t0 : t1;
        }
      }
      if (fabs(a[j * N + j]) >= 2.2250738585072014e-308) {
        const double rinv = div2(1.0, a[j * N + j]);
#pragma unroll
        for (int i = j + 1; i < N; ++i) a[i * N + j] = mul2(a[i * N + j], rinv);
      } else {
#pragma unroll
        for (int i = j + 1; i < N; ++i) a[i * N + j] = div2(a[i * N + j], a[j * N + j]);
      }
    } else {
      singular = true;
    }
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      const double lij = a[i * N + j];
#pragma unroll
      for (int l = j + 1; l < N; ++l) a[i * N + l] = add2(a[i * N + l], -mul2(lij, a[j * N + l]));
    }
  }
  if (singular) return 1;
#pragma unroll
  for (int j = 0; j < N; ++j) {  // dtrti2, upper, non-unit
    a[j * N + j] = div2(1.0, a[j * N + j]);
    const double ajj = -a[j * N + j];
#pragma unroll
    for (int i = 0; i < j; ++i) {
      double t = mul2(a[i * N + i], a[i * N + j]);
#pragma unroll
      for (int l = i + 1; l < j; ++l) t = add2(t, mul2(a[i * N + l], a[l * N + j]));
      a[i * N + j] = t;
    }
#pragma unroll
    for (int i = 0; i < j; ++i) a[i * N + j] = mul2(a[i * N + j], ajj);
  }
#pragma unroll
  for (int j = N - 2; j >= 0; --j) {  // dgetri, unblocked (j = N-1 only clears nothing)
    double work[N];
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      work[i] = a[i * N + j];
      a[i * N + j] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double t = 0.0;
#pragma unroll
      for (int l = j + 1; l < N; ++l) t = add2(t, mul2(a[i * N + l], work[l]));
      a[i * N + j] = add2(a[i * N + j], mul2(-1.0, t));
    }
  }
#pragma unroll
  for (int j = N - 2; j >= 0; --j) {
#pragma unroll
    for (int jp = j + 1; jp < N; ++jp) {
      const bool sw = (piv[j] == jp);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double t0 = a[i * N + j], t1 = a[i * N + jp];
        a[i * N + j] = sw ? t1 : t0;
        a[i * N + jp] = sw ? t0 : t1;
      }
    }
  }
  double inorm = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) s = add2(s, fabs(a[i * N + j]));
    if (s > inorm) inorm = s;
  }
  return (mul2(anorm, inorm) <= 1e16) ? 0 : 2;
}

// helper.go:65-84 AsSymDense: error when an off-diagonal pair differs by more than 1e-6 absolute AND more than 1e-2
// relative (floats.EqualWithinAbsOrRel); on success the upper triangle is mirrored (mat64.SymDense reads only it).
template <int N>
GKB_DEV bool as_sym(double (&A)[N * N]) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) {
      const double a = A[j * N + i], b = A[i * N + j];
      const double d = fabs(a - b);
      const bool eq = (a == b) || (d <= 1e-6) || (d / fmax(fabs(a), fabs(b)) <= 1e-2);
      ok = ok && eq;
    }
  if (!ok) return false;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) A[i * N + j] = A[j * N + i];
  return true;
}

// hybrid.go:104-204 fullUpdate, statement by statement.  x, P (dense, mirrored upper triangle): previous estimate
// in, new one out (untouched when an error is returned).  Ppred / K / innov / obsdev: the Estimate fields.
template <int N, int M>
GKB_DEV int hybrid_step(const NlModel<N, M>& md, double (&x)[N], double (&P)[N * N], const double (&Phi)[N * N],
                        const double (&Ht)[M * N], const double (&real_obs)[M], const double (&computed_obs)[M],
                        const double* __restrict__ Gamma, bool has_meas, bool ekf, bool snc, double (&Ppred)[N * N],
                        double (&K)[N * M], double (&innov)[M], double (&obsdev)[M]) {
  // 114-117: P-bar = (Phi P) Phi^T.  Row i of Phi P is formed and consumed at once (the entries, and the order of every
  // sum, are those of the two full products: only the live range of the intermediate shrinks from N*N to N values)
  double Pbar[N * N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double row[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < N; ++l) s = add2(s, mul2(Phi[i * N + l], P[l * N + j]));
      row[j] = s;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < N; ++l) s = add2(s, mul2(row[l], Phi[j * N + l]));
      Pbar[i * N + j] = s;
    }
  }
  if (snc && Gamma != nullptr) {  // 118-123: P-bar + (Gamma Q) Gamma^T
    const int q = md.q;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double gq[GKB_MAX_Q];
#pragma unroll
      for (int a = 0; a < GKB_MAX_Q; ++a) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < GKB_MAX_Q; ++b)
          if (a < q && b < q) s = add2(s, mul2(__ldg(Gamma + i * q + b), md.Q[b * q + a]));
        gq[a] = s;
      }
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < GKB_MAX_Q; ++a)
          if (a < q) s = add2(s, mul2(gq[a], __ldg(Gamma + j * q + a)));
        Pbar[i * N + j] = add2(Pbar[i * N + j], s);
      }
    }
  }
  if (!has_meas) {  // Predict(): 125-143
    double xbar[N];
    if (ekf) {
#pragma unroll
      for (int i = 0; i < N; ++i) xbar[i] = 0.0;
    } else {
      mulvec<N, N>(xbar, Phi, x);
    }
    if (!as_sym<N>(Pbar)) return GKB_ERR_ASYMMETRIC;
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = xbar[i];
#pragma unroll
    for (int i = 0; i < N * N; ++i) { P[i] = Pbar[i]; Ppred[i] = Pbar[i]; }
#pragma unroll
    for (int a = 0; a < M; ++a) { innov[a] = 0.0; obsdev[a] = 0.0; }
#pragma unroll
    for (int i = 0; i < N * M; ++i) K[i] = 0.0;
    return 0;
  }
  // 146-153
  double PHt[N * M];
  mul_nt<N, N, M>(PHt, Pbar, Ht);  // P-bar Ht^T (the transposed copy of the reference holds the same numbers)
  double S[M * M];
  mul<M, N, M>(S, Ht, PHt);
#pragma unroll
  for (int i = 0; i < M * M; ++i) S[i] = add2(S[i], md.R[i]);
  if (inverse<M>(S) != 0) return GKB_ERR_SINGULAR_S;
  double Kn[N * M];
  mul<N, M, M>(Kn, PHt, S);
  // 156-173
  double y[M], inn[M], xhat[N];
#pragma unroll
  for (int a = 0; a < M; ++a) { y[a] = add2(real_obs[a], -computed_obs[a]); inn[a] = 0.0; }
  if (ekf) {
    mulvec<N, M>(xhat, Kn, y);
  } else {
    double xbar[N], t1[M];
    mulvec<N, N>(xbar, Phi, x);
    mulvec<M, N>(t1, Ht, xbar);
#pragma unroll
    for (int a = 0; a < M; ++a) inn[a] = add2(y[a], -t1[a]);
    mulvec<N, M>(xhat, Kn, inn);
#pragma unroll
    for (int i = 0; i < N; ++i) xhat[i] = add2(xbar[i], xhat[i]);
  }
  // 174-182: dense Joseph form ((I - K H) P-bar) (I - K H)^T + (K R) K^T
  double Pn[N * N];
  {
    double KH[N * N];
    mul<N, M, N>(KH, Kn, Ht);
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) KH[i * N + j] = add2(i == j ? 1.0 : 0.0, -KH[i * N + j]);
    // ((I - K H) P-bar) (I - K H)^T + (K R) K^T, row by row (same entries and summation orders as the full products)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double t1[N], kr[M];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < N; ++l) s = add2(s, mul2(KH[i * N + l], Pbar[l * N + j]));
        t1[j] = s;
      }
#pragma unroll
      for (int a = 0; a < M; ++a) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < M; ++b) s = add2(s, mul2(Kn[i * M + b], md.R[b * M + a]));
        kr[a] = s;
      }
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < N; ++l) s = add2(s, mul2(t1[l], KH[j * N + l]));
        double r = 0.0;
#pragma unroll
        for (int a = 0; a < M; ++a) r = add2(r, mul2(kr[a], Kn[j * M + a]));
        Pn[i * N + j] = add2(s, r);
      }
    }
  }
  // 184-192
  if (!as_sym<N>(Pbar)) return GKB_ERR_ASYMMETRIC;
  if (!as_sym<N>(Pn)) return GKB_ERR_ASYMMETRIC;
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = xhat[i];
#pragma unroll
  for (int i = 0; i < N * N; ++i) { P[i] = Pn[i]; Ppred[i] = Pbar[i]; }
#pragma unroll
  for (int i = 0; i < N * M; ++i) K[i] = Kn[i];
#pragma unroll
  for (int a = 0; a < M; ++a) { innov[a] = inn[a]; obsdev[a] = y[a]; }
  return 0;
}


// ---------------------------------------------------------------------------------------------------------------
// The same step with the N x N matrices in lane-private SHARED-MEMORY columns and ROLLED loops (round 2).
//
// The fully unrolled register version above is ~56 KB of SASS per epoch at n = 6 (it streams through the
// instruction cache: `no_instruction` 0.47 stalls per issue) and spills (255 registers + 732 B).  Here a thread's
// matrices live at sm[e * kStrictThreads + threadIdx.x] (entry-major over the CTA: consecutive lanes, consecutive banks, no
// conflicts) and each of the four dense products is a loop of N iterations -- one column (or row) of the stored
// operand per iteration against a matrix held in registers with static indices:
//     T = Phi P        by columns of P   (Ps -> Ws)
//     P-bar = T Phi^T  by rows of T      (Ws, in place: row i of T is dead once row i of P-bar is formed)
//     X = A P-bar      by columns        (Ws, in place),  A = I - K H in the registers Phi occupied
//     P+ = X A^T + (K R) K^T  by rows    (Ws, in place)
// Every entry is the same sequence of individually rounded multiplications and additions, in the same order, as in
// the full products of the reference -- only WHERE the operands wait changes.  One deliberate difference: the
// reference's (and gonum's) sums start from 0; `0 + a*b` is elided here.  That is exact for every value except
// that (-0) would become +0: only the sign of an exactly-zero entry can differ, which no later operation of the step
// turns into a different number (zeros are only multiplied, added and compared; pivots are tested `!= 0` first).
//
// Ps holds the previous covariance as the PACKED upper triangle (what AsSymDense keeps: the reference reads P[i][j], i > j,
// from the upper half) and is NOT modified by the step: on success the new covariance is in Ws and the caller copies its
// upper triangle into Ps (hybrid_sm_commit); on an error nothing of the previous estimate has been touched (hybrid.go
// returns (nil, err)).  Per thread: N(N+1)/2 + N*N + (N*M + 2M) doubles + N*N for the kernel's Phi stage: 872 B at n = 6,
// m = 2, i.e. 109 KB per CTA of 128 threads, two CTAs per SM.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStrictThreads = 128;  // threads per CTA of the shared-memory strict kernel (measured: 2 CTAs of 128 at ~250
constexpr int kStrictMinBlocks = 2;  // registers beat 3 x 96 and 2 x 160 at 168 registers, which spill)
#define GKB_SM(p, e) (p)[(e) * kStrictThreads]

// dst = Mr * src, column by column (Mr: registers, row-major); src == dst allowed
template <int N>
GKB_DEV void sm_left_mul(const double (&Mr)[N * N], const double* src, double* dst) {
#pragma unroll 1
  for (int j = 0; j < N; ++j) {
    double p[N], o[N];
#pragma unroll
    for (int l = 0; l < N; ++l) p[l] = GKB_SM(src, l * N + j);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = mul2(Mr[i * N], p[0]);
#pragma unroll
      for (int l = 1; l < N; ++l) s = add2(s, mul2(Mr[i * N + l], p[l]));
      o[i] = s;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) GKB_SM(dst, i * N + j) = o[i];
  }
}

// dst = Mr * S, S symmetric, stored as its packed upper triangle (static addresses: fully unrolled)
template <int N>
GKB_DEV void sm_left_mul_sym(const double (&Mr)[N * N], const double* src, double* dst) {
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double p[N];
#pragma unroll
    for (int l = 0; l < N; ++l) p[l] = GKB_SM(src, sym_idx<N>(l, j));
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = mul2(Mr[i * N], p[0]);
#pragma unroll
      for (int l = 1; l < N; ++l) s = add2(s, mul2(Mr[i * N + l], p[l]));
      GKB_SM(dst, i * N + j) = s;
    }
  }
}

// P := AsSymDense(W): the upper triangle of the (already tested) work matrix becomes the packed covariance
template <int N>
GKB_DEV void hybrid_sm_commit(double* Ps, const double* Ws) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) GKB_SM(Ps, sym_idx<N>(i, j)) = GKB_SM(Ws, i * N + j);
}

// AsSymDense on a shared-memory matrix (helper.go:65-84): false when an off-diagonal pair differs by more than 1e-6
// absolute AND more than 1e-2 relative; MIRROR copies the upper triangle down (mat64.SymDense reads only the upper one).
// One straight-line pass decides the common case (every pair within 1e-6) and mirrors the pairs it has accepted; only
// when a pair fails that test does the exact predicate run, on the pairs as they still stand.
template <int N, bool MIRROR>
GKB_DEV bool sm_as_sym(double* A) {
  bool fast = true;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) {
      const double a = GKB_SM(A, j * N + i), b = GKB_SM(A, i * N + j);
      const bool pass = fabs(a - b) <= 1e-6;
      fast = fast && pass;
      if (MIRROR && pass) GKB_SM(A, i * N + j) = a;
    }
  if (fast) return true;
  bool ok = true;
#pragma unroll 1
  for (int i = 0; i < N; ++i)
#pragma unroll 1
    for (int j = 0; j < i; ++j) {
      const double a = GKB_SM(A, j * N + i), b = GKB_SM(A, i * N + j);
      const double d = fabs(a - b);
      ok = ok && ((a == b) || (d <= 1e-6) || (d / fmax(fabs(a), fabs(b)) <= 1e-2));
    }
  if (!ok) return false;
  if (MIRROR) {
#pragma unroll 1
    for (int i = 0; i < N; ++i)
#pragma unroll 1
      for (int j = 0; j < i; ++j) GKB_SM(A, i * N + j) = GKB_SM(A, j * N + i);
  }
  return true;
}

// Phase A (hybrid.go:114-123 and the x-bar of 126 / 163): Ws = P-bar, xbar = Phi x (CKF).  Phi is dead afterwards.
template <int N, int M>
GKB_DEV void hybrid_sm_predict(const NlModel<N, M>& md, const double (&x)[N], const double* Ps, double* Ws,
                               const double (&Phi)[N * N], const double* __restrict__ Gamma, bool snc, bool ekf,
                               double (&xbar)[N]) {
  sm_left_mul_sym<N>(Phi, Ps, Ws);  // T = Phi P
#pragma unroll 1
  for (int i = 0; i < N; ++i) {  // P-bar = T Phi^T, row i in place
    double t[N], o[N];
#pragma unroll
    for (int l = 0; l < N; ++l) t[l] = GKB_SM(Ws, i * N + l);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = mul2(t[0], Phi[j * N]);
#pragma unroll
      for (int l = 1; l < N; ++l) s = add2(s, mul2(t[l], Phi[j * N + l]));
      o[j] = s;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) GKB_SM(Ws, i * N + j) = o[j];
  }
  if (snc && Gamma != nullptr) {  // 118-123: P-bar + (Gamma Q) Gamma^T  (rare: sums from 0 as written)
    const int q = md.q;
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
      double gq[GKB_MAX_Q];
#pragma unroll
      for (int a = 0; a < GKB_MAX_Q; ++a) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < GKB_MAX_Q; ++b)
          if (a < q && b < q) s = add2(s, mul2(__ldg(Gamma + i * q + b), md.Q[b * q + a]));
        gq[a] = s;
      }
#pragma unroll 1
      for (int j = 0; j < N; ++j) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < GKB_MAX_Q; ++a)
          if (a < q) s = add2(s, mul2(gq[a], __ldg(Gamma + j * q + a)));
        GKB_SM(Ws, i * N + j) = add2(GKB_SM(Ws, i * N + j), s);
      }
    }
  }
  // x-bar = Phi x is formed by the reference only where it is used (hybrid.go:127-131, 162-164): never in EKF mode
  // (`ekf` is shared by the batch: a uniform branch)
#pragma unroll
  for (int i = 0; i < N; ++i) xbar[i] = 0.0;
  if (!ekf) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = mul2(Phi[i * N], x[0]);
#pragma unroll
      for (int j = 1; j < N; ++j) s = add2(s, mul2(Phi[i * N + j], x[j]));
      xbar[i] = s;
    }
  }
}

// Phase B (hybrid.go:125-204).  Ws: P-bar in, the new covariance out.  pred_out (global, this filter's column, or
// nullptr) receives AsSymDense(P-bar) -- the caller overwrites it with NaN if the step fails after that point.
template <int N, int M>
GKB_DEV int hybrid_sm_update(const NlModel<N, M>& md, double (&x)[N], double* Ws, double* KRs, const double (&xbar)[N],
                             const double (&Ht)[M * N], const double (&real_obs)[M], const double (&computed_obs)[M],
                             bool has_meas, bool ekf, double* __restrict__ pred_out, int64_t nf, double (&K)[N * M],
                             double (&innov)[M], double (&obsdev)[M]) {
  auto write_pred = [&]() {
    if (pred_out == nullptr) return;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) __stcs(pred_out + (int64_t)(i * N + j) * nf, GKB_SM(Ws, (i <= j) ? (i * N + j) : (j * N + i)));
  };
  if (!has_meas) {  // Predict(): 125-143
    if (!sm_as_sym<N, false>(Ws)) return GKB_ERR_ASYMMETRIC;
    write_pred();
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = ekf ? 0.0 : xbar[i];
#pragma unroll
    for (int a = 0; a < M; ++a) { innov[a] = 0.0; obsdev[a] = 0.0; }
#pragma unroll
    for (int i = 0; i < N * M; ++i) K[i] = 0.0;
    return 0;
  }
  // 146-153
  double PHt[N * M];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double row[N];
#pragma unroll
    for (int l = 0; l < N; ++l) row[l] = GKB_SM(Ws, i * N + l);
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = mul2(row[0], Ht[a * N]);
#pragma unroll
      for (int l = 1; l < N; ++l) s = add2(s, mul2(row[l], Ht[a * N + l]));
      PHt[i * M + a] = s;
    }
  }
  double S[M * M];
#pragma unroll
  for (int a = 0; a < M; ++a)
#pragma unroll
    for (int b = 0; b < M; ++b) {
      double s = mul2(Ht[a * N], PHt[b]);
#pragma unroll
      for (int l = 1; l < N; ++l) s = add2(s, mul2(Ht[a * N + l], PHt[l * M + b]));
      S[a * M + b] = add2(s, md.R[a * M + b]);
    }
  if (inverse<M>(S) != 0) return GKB_ERR_SINGULAR_S;
  double Kn[N * M];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = mul2(PHt[i * M], S[a]);
#pragma unroll
      for (int b = 1; b < M; ++b) s = add2(s, mul2(PHt[i * M + b], S[b * M + a]));
      Kn[i * M + a] = s;
    }
  // 156-173
  double y[M], inn[M], xhat[N];
#pragma unroll
  for (int a = 0; a < M; ++a) { y[a] = add2(real_obs[a], -computed_obs[a]); inn[a] = 0.0; }
  if (ekf) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = mul2(Kn[i * M], y[0]);
#pragma unroll
      for (int a = 1; a < M; ++a) s = add2(s, mul2(Kn[i * M + a], y[a]));
      xhat[i] = s;
    }
  } else {
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = mul2(Ht[a * N], xbar[0]);
#pragma unroll
      for (int l = 1; l < N; ++l) s = add2(s, mul2(Ht[a * N + l], xbar[l]));
      inn[a] = add2(y[a], -s);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = mul2(Kn[i * M], inn[0]);
#pragma unroll
      for (int a = 1; a < M; ++a) s = add2(s, mul2(Kn[i * M + a], inn[a]));
      xhat[i] = add2(xbar[i], s);
    }
  }
  // 184-187 (moved up: P-bar is not modified between 117 and 184, and X below overwrites it)
  if (!sm_as_sym<N, false>(Ws)) return GKB_ERR_ASYMMETRIC;
  write_pred();
  // 174-182: dense Joseph form ((I - K H) P-bar) (I - K H)^T + (K R) K^T
  {
    double A[N * N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = mul2(Kn[i * M], Ht[j]);
#pragma unroll
        for (int a = 1; a < M; ++a) s = add2(s, mul2(Kn[i * M + a], Ht[a * N + j]));
        A[i * N + j] = (i == j) ? add2(1.0, -s) : -s;  // (0 + (-s) elided off the diagonal: exact up to the sign of a zero)
      }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int a = 0; a < M; ++a) {
        double s = mul2(Kn[i * M], md.R[a]);
#pragma unroll
        for (int b = 1; b < M; ++b) s = add2(s, mul2(Kn[i * M + b], md.R[b * M + a]));
        GKB_SM(KRs, i * M + a) = s;
      }
    sm_left_mul<N>(A, Ws, Ws);  // X = (I - K H) P-bar
#pragma unroll 1
    for (int i = 0; i < N; ++i) {  // P+ row i = X row i (I - K H)^T + (K R) row i K^T
      double t[N], kr[M], o[N];
#pragma unroll
      for (int l = 0; l < N; ++l) t[l] = GKB_SM(Ws, i * N + l);
#pragma unroll
      for (int a = 0; a < M; ++a) kr[a] = GKB_SM(KRs, i * M + a);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = mul2(t[0], A[j * N]);
#pragma unroll
        for (int l = 1; l < N; ++l) s = add2(s, mul2(t[l], A[j * N + l]));
        double r = mul2(kr[0], Kn[j * M]);
#pragma unroll
        for (int a = 1; a < M; ++a) r = add2(r, mul2(kr[a], Kn[j * M + a]));
        o[j] = add2(s, r);
      }
#pragma unroll
      for (int j = 0; j < N; ++j) GKB_SM(Ws, i * N + j) = o[j];
    }
  }
  // 188-192 (the caller keeps the upper triangle: hybrid_sm_commit)
  if (!sm_as_sym<N, false>(Ws)) return GKB_ERR_ASYMMETRIC;
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = xhat[i];
#pragma unroll
  for (int i = 0; i < N * M; ++i) K[i] = Kn[i];
#pragma unroll
  for (int a = 0; a < M; ++a) { innov[a] = inn[a]; obsdev[a] = y[a]; }
  return 0;
}

}  // namespace strict
}  // namespace gkb
