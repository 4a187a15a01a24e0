// filters.cuh -- per-thread filter steps (one filter per thread, state in FP64 registers).
//
// Each *_step function advances ONE filter by ONE Update() of the reference and is shared by the
// batched-update kernels (kernels_lti.cu, kernels_nl.cu) and the fused Monte Carlo / chi-square
// kernel (kernels_mc.cu).  Models are passed to the kernels by value, so their entries are read
// from the constant bank directly as DFMA operands and cost no registers.
//
// Arithmetic notes (DESIGN.md "numerics"):
//  * covariances are kept in packed symmetric form (the reference's AsSymDense keeps the upper
//    triangle, helper.go:65-84), so F P F^T and the Joseph update only form the upper triangle;
//  * the Joseph form (I-KH) P- (I-KH)^T + K R K^T (vanilla.go:197-205, hybrid.go:174-182) is
//    evaluated as T = P- - K (P- H^T)^T, P+ = T - (T H^T - K R) K^T: the same quantity with the
//    identity multiplications removed, O(n^2 m) instead of O(n^3);
//  * inverses of S = H P- H^T + R use the same LU-with-partial-pivoting explicit inverse and the
//    same error conditions as the reference.
#pragma once
#include "smallmat.cuh"
#include "../../include/gokalman_b200.h"

namespace gkb {

template <int N, int M>
struct VanillaModel {  // vanilla.go:65-74
  double F[N * N];
  double G[N * GKB_MAX_C];
  double H[M * N];
  double Q[N * N];
  double R[M * M];
  int c;
  int need_ctrl;
};

template <int N, int M>
struct InfoModel {  // information.go:84-95
  double Finv[N * N];
  double G[N * GKB_MAX_C];
  double H[M * N];
  double Qinv[N * N];
  double Rinv[M * M];
  double R[M * M];  // the filter's CURRENT measurement noise (GetNoise().MeasurementMatrix()): chi-square NIS only
  int rinv_dim;  // 1 => scalar broadcast path (information.go:197-199)
  int c;
  int need_ctrl;
};

template <int N, int M>
struct SqrtModel {  // squareroot.go:53-63
  double F[N * N];
  double G[N * GKB_MAX_C];
  double H[M * N];
  double sqrtQ[N * N];  // lower Cholesky factors
  double sqrtR[M * M];
  int c;
  int need_ctrl;
};

template <int N, int M>
struct NlModel {  // hybrid.go:37-46 / srif.go:52-60 (shared part)
  double Q[GKB_MAX_Q * GKB_MAX_Q];
  double R[M * M];
  double L[M * M];  // SRIF: chol_lower(R) kept as "sqrtInvNoise" (srif.go:48)
  int q;
};

// What one Update() produces besides the new state (the Estimate fields, kalman.go:64-72).
template <int N, int M>
struct StepOut {
  double yhat[M];
  double innov[M];
  double Ppred[N * (N + 1) / 2];
  double K[N * M];
  double Sinv[M * M];
};

// ---- vanilla.go:128-220 ---------------------------------------------------------------------------
// x, P: previous posterior in, new posterior out (unchanged when an error is returned).
// gu = G u (formed once per step for all filters by gu_kernel).  NOISY = false compiles out the
// "+ Process(k)" / "+ Measurement(k)" additions for a Noiseless filter (they would add +0.0);
// CHECK = false drops the engine's own non-finite guard (the reference has none).
// w is what the first Process(k) call returns (vanilla.go:146), w2 what the second one returns (vanilla.go:195):
// the same vector for BatchNoise-style replay (noise.go:73-78), two different draws for AWGN (noise.go:127-131).
template <int N, int M, bool PREDICTOR, bool NOISY = true, bool CHECK = true>
GKB_DEV int vanilla_step(const VanillaModel<N, M>& md, double (&x)[N], double (&P)[N * (N + 1) / 2],
                         const double (&y)[M], const double (&gu)[N], const double (&w)[N],
                         const double (&v)[M], StepOut<N, M>& o, const double (&w2)[N]) {
  constexpr int SN = N * (N + 1) / 2;
  // 138-146: x- = F x + G u + Process(k)
  double xm[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = md.F[i * N] * x[0];
#pragma unroll
    for (int j = 1; j < N; ++j) s = fma(md.F[i * N + j], x[j], s);
    if (md.need_ctrl) s += gu[i];
    xm[i] = NOISY ? (s + w[i]) : s;
  }
  // 155-157: yhat = H x_prev + Measurement(k)
#pragma unroll
  for (int a = 0; a < M; ++a) {
    double s = md.H[a * N] * x[0];
#pragma unroll
    for (int j = 1; j < N; ++j) s = fma(md.H[a * N + j], x[j], s);
    o.yhat[a] = NOISY ? (s + v[a]) : s;
  }
  // 149-152: P- = (F P) F^T + Q, upper triangle only
  double Pm[SN];
  {
    double FP[N * N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = md.F[i * N] * P[sym_idx<N>(0, j)];
#pragma unroll
        for (int l = 1; l < N; ++l) s = fma(md.F[i * N + l], P[sym_idx<N>(l, j)], s);
        FP[i * N + j] = s;
      }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) {
        double s = md.Q[i * N + j];
#pragma unroll
        for (int l = 0; l < N; ++l) s = fma(FP[i * N + l], md.F[j * N + l], s);
        Pm[sym_idx<N>(i, j)] = s;
      }
  }
  // 160-168: K = P- H^T inv(H P- H^T + R)
  double PHt[N * M];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = Pm[sym_idx<N>(i, 0)] * md.H[a * N];
#pragma unroll
      for (int j = 1; j < N; ++j) s = fma(Pm[sym_idx<N>(i, j)], md.H[a * N + j], s);
      PHt[i * M + a] = s;
    }
  double S[M * M];
#pragma unroll
  for (int a = 0; a < M; ++a)
#pragma unroll
    for (int b = 0; b < M; ++b) {
      double s = md.R[a * M + b];
#pragma unroll
      for (int i = 0; i < N; ++i) s = fma(md.H[a * N + i], PHt[i * M + b], s);
      S[a * M + b] = s;
    }
  if (inverse_lu<M>(S) != 0) return GKB_ERR_SINGULAR_S;
#pragma unroll
  for (int i = 0; i < M * M; ++i) o.Sinv[i] = S[i];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = PHt[i * M] * S[a];
#pragma unroll
      for (int b = 1; b < M; ++b) s = fma(PHt[i * M + b], S[b * M + a], s);
      o.K[i * M + a] = s;
    }
#pragma unroll
  for (int i = 0; i < SN; ++i) o.Ppred[i] = Pm[i];
  if constexpr (PREDICTOR) {  // 170-179
#pragma unroll
    for (int a = 0; a < M; ++a) o.innov[a] = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = xm[i];
#pragma unroll
    for (int i = 0; i < SN; ++i) P[i] = Pm[i];
    return 0;
  } else {
    // 182-195: innovation, x+ = x- + K nu + Process(k)
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = md.H[a * N] * xm[0];
#pragma unroll
      for (int j = 1; j < N; ++j) s = fma(md.H[a * N + j], xm[j], s);
      o.innov[a] = y[a] - s;
    }
    double xp[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = o.K[i * M] * o.innov[0];
#pragma unroll
      for (int a = 1; a < M; ++a) s = fma(o.K[i * M + a], o.innov[a], s);
      xp[i] = NOISY ? ((xm[i] + s) + w2[i]) : (xm[i] + s);
    }
    // 197-205: Joseph form, restructured (see header)
    double Pp[SN];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double T[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = Pm[sym_idx<N>(i, j)];
#pragma unroll
        for (int a = 0; a < M; ++a) s = fma(-o.K[i * M + a], PHt[j * M + a], s);
        T[j] = s;
      }
      double V[M];
#pragma unroll
      for (int a = 0; a < M; ++a) {
        double s = T[0] * md.H[a * N];
#pragma unroll
        for (int j = 1; j < N; ++j) s = fma(T[j], md.H[a * N + j], s);
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
    if constexpr (CHECK) {
      bool finite = true;
#pragma unroll
      for (int i = 0; i < N; ++i) finite = finite && isfinite(xp[i]) && isfinite(Pp[sym_idx<N>(i, i)]);
      if (!finite) return GKB_ERR_NONFINITE;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = xp[i];
#pragma unroll
    for (int i = 0; i < SN; ++i) P[i] = Pp[i];
    return 0;
  }
}

// BatchNoise / Noiseless form: both Process(k) calls see the same vector.
template <int N, int M, bool PREDICTOR, bool NOISY = true, bool CHECK = true>
GKB_DEV int vanilla_step(const VanillaModel<N, M>& md, double (&x)[N], double (&P)[N * (N + 1) / 2],
                         const double (&y)[M], const double (&gu)[N], const double (&w)[N],
                         const double (&v)[M], StepOut<N, M>& o) {
  return vanilla_step<N, M, PREDICTOR, NOISY, CHECK>(md, x, P, y, gu, w, v, o, w);
}

}  // namespace gkb
