// filters_info_sqrt.cuh -- per-thread Information and Square-Root filter steps.
#pragma once
#include "filters.cuh"

namespace gkb {

template <int N, int M>
struct InfoOut {
  double yhat[M];
  double Ipred[N * (N + 1) / 2];
};

// information.go:276-293 Covariance(): AsSymDense(inv(I)), or zeros when Inverse reports an error
// (exactly singular, or condition number above 1e16).  Returns true when invertible.
template <int N>
GKB_DEV bool info_covariance(double (&Pc)[N * N], const double (&I)[N * (N + 1) / 2]) {
  sym_expand<N>(Pc, I);
  int ierr = inverse_lu_fast<N>(Pc);
  if (ierr != 0) {
#pragma unroll
    for (int i = 0; i < N * N; ++i) Pc[i] = 0.0;
    return false;
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) Pc[i * N + j] = Pc[j * N + i];
  return true;
}

// information.go:153-227.  iv, I: previous information state / matrix in, new ones out.
// `want_yhat` (uniform) skips the State(prev) inverse when Estimate.Measurement() is not wanted.
template <int N, int M>
GKB_DEV int info_step(const InfoModel<N, M>& md, double (&iv)[N], double (&I)[N * (N + 1) / 2],
                      const double (&y)[M], const double (&gu)[N], const double (&v)[M], InfoOut<N, M>& o,
                      bool want_yhat = true) {
  constexpr int SN = N * (N + 1) / 2;
  // 192-194 (hoisted: it only reads the previous estimate): yhat = H State(prev) + Measurement(k)
  if (want_yhat) {
    double Pc[N * N], xs[N];
    info_covariance<N>(Pc, I);
    mulvec<N, N>(xs, Pc, iv);
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = md.H[a * N] * xs[0];
#pragma unroll
      for (int j = 1; j < N; ++j) s = fma(md.H[a * N + j], xs[j], s);
      o.yhat[a] = s + v[a];
    }
  } else {
#pragma unroll
    for (int a = 0; a < M; ++a) o.yhat[a] = 0.0;
  }
  // 161-165: z = Finv^T (I Finv)
  double z[N * N];
  {
    double IF[N * N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = I[sym_idx<N>(i, 0)] * md.Finv[j];
#pragma unroll
        for (int l = 1; l < N; ++l) s = fma(I[sym_idx<N>(i, l)], md.Finv[l * N + j], s);
        IF[i * N + j] = s;
      }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = md.Finv[i] * IF[j];
#pragma unroll
        for (int l = 1; l < N; ++l) s = fma(md.Finv[l * N + i], IF[l * N + j], s);
        z[i * N + j] = s;
      }
  }
  // 169-174: M = -(z inv(z + Qinv)); the Inverse error is ignored by the reference
  double Mm[N * N];
  {
    double T[N * N];
#pragma unroll
    for (int i = 0; i < N * N; ++i) T[i] = z[i] + md.Qinv[i];
    (void)inverse_lu_fast<N>(T);
    mul<N, N, N>(Mm, z, T);
#pragma unroll
    for (int i = 0; i < N * N; ++i) Mm[i] = -Mm[i];
  }
  // 176-185: i- = (I + M)(Finv^T i + z G u)
  double im[N];
  {
    double t[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = md.Finv[i] * iv[0];
#pragma unroll
      for (int l = 1; l < N; ++l) s = fma(md.Finv[l * N + i], iv[l], s);
      t[i] = s;
    }
    if (md.need_ctrl) {
      double zg[N];
      mulvec<N, N>(zg, z, gu);
#pragma unroll
      for (int i = 0; i < N; ++i) t[i] += zg[i];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) s = fma((i == j ? 1.0 : 0.0) + Mm[i * N + j], t[j], s);
      im[i] = s;
    }
  }
  // 187-190: I- = z + M z^T (upper triangle)
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) {
      double s = Mm[i * N] * z[j * N];
#pragma unroll
      for (int l = 1; l < N; ++l) s = fma(Mm[i * N + l], z[j * N + l], s);
      o.Ipred[sym_idx<N>(i, j)] = z[i * N + j] + s;
    }
  // 196-203: HTR = Rinv[0,0] H^T if Rinv is 1x1 (even for a 2-row H), else H^T Rinv
  double HTR[N * M];
  if (md.rinv_dim == 1) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int a = 0; a < M; ++a) HTR[i * M + a] = md.Rinv[0] * md.H[a * N + i];
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int a = 0; a < M; ++a) {
        double s = md.H[i] * md.Rinv[a];
#pragma unroll
        for (int b = 1; b < M; ++b) s = fma(md.H[b * N + i], md.Rinv[b * M + a], s);
        HTR[i * M + a] = s;
      }
  }
  // 205-212 (nothing is committed before the non-finite check: a failed Update leaves the previous estimate in place)
  bool finite = true;
  double ivn[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = HTR[i * M] * y[0];
#pragma unroll
    for (int a = 1; a < M; ++a) s = fma(HTR[i * M + a], y[a], s);
    ivn[i] = s + im[i];
    finite = finite && isfinite(ivn[i]);
  }
  if (!finite) return GKB_ERR_NONFINITE;
#pragma unroll
  for (int i = 0; i < N; ++i) iv[i] = ivn[i];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) {
      double s = HTR[i * M] * md.H[j];
#pragma unroll
      for (int a = 1; a < M; ++a) s = fma(HTR[i * M + a], md.H[a * N + j], s);
      I[sym_idx<N>(i, j)] = o.Ipred[sym_idx<N>(i, j)] + s;
    }
  (void)SN;
  return 0;
}

template <int N, int M>
struct SqrtOut {
  double yhat[M];
  double innov[M];
  double K[N * M];
  double Spred[N * N];  // the UPPER QR factor, used by the reference as if it were S- (squareroot.go:181-185)
};

// squareroot.go:129-274.  x, S (P = S S^T): previous estimate in, new one out.
template <int N, int M>
GKB_DEV int sqrt_step(const SqrtModel<N, M>& md, double (&x)[N], double (&S)[N * N], const double (&y)[M],
                      const double (&gu)[N], const double (&w)[N], const double (&v)[M], SqrtOut<N, M>& o) {
  // 140-147: x- = F x + G u (no process noise here)
  double xm[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = md.F[i * N] * x[0];
#pragma unroll
    for (int j = 1; j < N; ++j) s = fma(md.F[i * N + j], x[j], s);
    if (md.need_ctrl) s += gu[i];
    xm[i] = s;
  }
  // 237-239: yhat = H x_prev + Measurement(k)
#pragma unroll
  for (int a = 0; a < M; ++a) {
    double s = md.H[a * N] * x[0];
#pragma unroll
    for (int j = 1; j < N; ++j) s = fma(md.H[a * N + j], x[j], s);
    o.yhat[a] = s + v[a];
  }
  // 155-185: C = [S^T F^T ; sqrtQ^T], S- := top block of qr(C).R
  {
    double Cm[2 * N * N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = S[i] * md.F[j * N];
#pragma unroll
        for (int l = 1; l < N; ++l) s = fma(S[l * N + i], md.F[j * N + l], s);
        Cm[i * N + j] = s;
        Cm[(N + i) * N + j] = md.sqrtQ[j * N + i];
      }
    qr_r_inplace<2 * N, N>(Cm);
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) o.Spred[i * N + j] = (j >= i) ? Cm[i * N + j] : 0.0;
  }
  // 190-222: Delta = [[sqrtR^T, 0],[S-^T H^T, S-^T]] -> qr
  constexpr int D = N + M;
  double Dl[D * D];
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int cc = 0; cc < D; ++cc) {
      double val;
      if (cc < M) {
        if (r < M) {
          val = md.sqrtR[cc * M + r];
        } else {
          // (S-^T H^T)[r-M][cc] = sum_k S-[k][r-M] H[cc][k], S- upper => k <= r-M
          double s = 0.0;
#pragma unroll
          for (int l = 0; l < N; ++l)
            if (l <= r - M) s = fma(o.Spred[l * N + (r - M)], md.H[cc * N + l], s);
          val = s;
        }
      } else if (r < M) {
        val = 0.0;
      } else {
        val = o.Spred[(cc - M) * N + (r - M)];  // S-^T
      }
      Dl[r * D + cc] = val;
    }
  qr_r_inplace<D, D>(Dl);
  // 225-234: S+ = (U[m:,m:])^T, Syy = (U[:m,:m])^T, W = (U[:m,m:])^T
  double Syy[M * M];
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) Syy[i * M + j] = (i >= j) ? Dl[j * D + i] : 0.0;
  // 242-252: K = W inv(Syy); the reference's error check is dead code (tests `err`, not `invErr`)
  (void)inverse_lu<M>(Syy);
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = Dl[0 * D + (M + i)] * Syy[a];
#pragma unroll
      for (int b = 1; b < M; ++b) s = fma(Dl[b * D + (M + i)], Syy[b * M + a], s);
      o.K[i * M + a] = s;
    }
  // 255-268
#pragma unroll
  for (int a = 0; a < M; ++a) {
    double s = md.H[a * N] * xm[0];
#pragma unroll
    for (int j = 1; j < N; ++j) s = fma(md.H[a * N + j], xm[j], s);
    o.innov[a] = y[a] - s;
  }
  bool finite = true;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = o.K[i * M] * o.innov[0];
#pragma unroll
    for (int a = 1; a < M; ++a) s = fma(o.K[i * M + a], o.innov[a], s);
    x[i] = (xm[i] + s) + w[i];
    finite = finite && isfinite(x[i]);
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) S[i * N + j] = (j <= i) ? Dl[(M + j) * D + (M + i)] : 0.0;
  return finite ? 0 : GKB_ERR_NONFINITE;
}

}  // namespace gkb
