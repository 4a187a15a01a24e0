// kernels_c2d.cu -- Van Loan's continuous-to-discrete conversion (c2d.go:13-75) for a BATCH of systems on the device.
//
// The reference computes, for one system (A, Gamma, W, dt):  M = [[-A dt, Gamma W Gamma^T dt], [0, A^T dt]],
// E = expm(M) (mat64.Dense.Exp),  F = (E_22)^T,  Q = AsSymDense(F E_12).  It is the step BEFORE the filters when every
// filter of a batch carries its own model (a parameter sweep over A or dt): SURVEY 8(f) rank 4.  One thread per
// system; the 2n x 2n matrices (n <= 8) live in local memory, which is fine for a constructor-time operation.
// expm: Higham's scaling and squaring with Pade approximants of degree 3 / 5 / 7 / 9 / 13 chosen by the 1-norm
// (N. J. Higham, "The scaling and squaring method for the matrix exponential revisited", 2005: Algorithm 10.20 of
// "Functions of Matrices") -- the algorithm behind gonum's Dense.Exp and scipy.linalg.expm's backbone.
// The Nyquist warning of c2d.go:15-28 needs A's eigenvalues and stays on the host (api.VanLoan).
#include "engine_internal.h"

namespace gkb {

namespace {
constexpr int kC2dMax = 2 * GKB_MAX_N;  // 16

struct Mat {
  double a[kC2dMax * kC2dMax];
};

__device__ void mm(Mat& C, const Mat& A, const Mat& B, int d) {  // C = A B (C distinct from A, B)
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      double s = 0.0;
      for (int l = 0; l < d; ++l) s = fma(A.a[i * d + l], B.a[l * d + j], s);
      C.a[i * d + j] = s;
    }
}
__device__ double norm1(const Mat& A, int d) {
  double best = 0.0;
  for (int j = 0; j < d; ++j) {
    double s = 0.0;
    for (int i = 0; i < d; ++i) s += fabs(A.a[i * d + j]);
    best = fmax(best, s);
  }
  return best;
}
// X <- inv(P) Qm by LU with partial pivoting on P (destroyed); returns false when P is singular
__device__ bool solve(Mat& P, Mat& Qm, int d) {
  for (int j = 0; j < d; ++j) {
    int p = j;
    double pm = fabs(P.a[j * d + j]);
    for (int i = j + 1; i < d; ++i)
      if (fabs(P.a[i * d + j]) > pm) { pm = fabs(P.a[i * d + j]); p = i; }
    if (pm == 0.0) return false;
    if (p != j)
      for (int l = 0; l < d; ++l) {
        double t = P.a[j * d + l]; P.a[j * d + l] = P.a[p * d + l]; P.a[p * d + l] = t;
        t = Qm.a[j * d + l]; Qm.a[j * d + l] = Qm.a[p * d + l]; Qm.a[p * d + l] = t;
      }
    const double inv = 1.0 / P.a[j * d + j];
    for (int i = j + 1; i < d; ++i) {
      const double l_ij = P.a[i * d + j] * inv;
      if (l_ij == 0.0) continue;
      for (int l = j + 1; l < d; ++l) P.a[i * d + l] = fma(-l_ij, P.a[j * d + l], P.a[i * d + l]);
      for (int l = 0; l < d; ++l) Qm.a[i * d + l] = fma(-l_ij, Qm.a[j * d + l], Qm.a[i * d + l]);
    }
  }
  for (int i = d - 1; i >= 0; --i) {
    const double inv = 1.0 / P.a[i * d + i];
    for (int l = 0; l < d; ++l) {
      double s = Qm.a[i * d + l];
      for (int k = i + 1; k < d; ++k) s = fma(-P.a[i * d + k], Qm.a[k * d + l], s);
      Qm.a[i * d + l] = s * inv;
    }
  }
  return true;
}

__constant__ double kPade3[4] = {120., 60., 12., 1.};
__constant__ double kPade5[6] = {30240., 15120., 3360., 420., 30., 1.};
__constant__ double kPade7[8] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
__constant__ double kPade9[10] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
__constant__ double kPade13[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                                   129060195264000., 10559470521600., 670442572800., 33522128640., 1323241920.,
                                   40840800., 960960., 16380., 182., 1.};
__constant__ double kTheta[5] = {1.495585217958292e-2, 2.539398330063230e-1, 9.504178996162932e-1, 2.097847961257068e0,
                                 5.371920351148152e0};

// E <- expm(A) (A destroyed).  Returns false when the Pade denominator is singular (never for finite input).
__device__ bool expm(Mat& E, Mat& A, int d) {
  const double nrm = norm1(A, d);
  Mat A2, U, V, T;
  const int degs[4] = {3, 5, 7, 9};
  for (int t = 0; t < 4; ++t) {
    if (nrm <= kTheta[t]) {
      const int m = degs[t];
      const double* b = m == 3 ? kPade3 : m == 5 ? kPade5 : m == 7 ? kPade7 : kPade9;
      mm(A2, A, A, d);
      // U = A (b1 I + b3 A2 + b5 A4 + ...), V = b0 I + b2 A2 + b4 A4 + ...   (P = running even power of A)
      Mat P, Us;
      for (int i = 0; i < d * d; ++i) { P.a[i] = (i / d == i % d) ? 1.0 : 0.0; Us.a[i] = 0.0; V.a[i] = 0.0; }
      for (int k = 0; 2 * k + 1 <= m; ++k) {
        for (int i = 0; i < d * d; ++i) { Us.a[i] = fma(b[2 * k + 1], P.a[i], Us.a[i]); V.a[i] = fma(b[2 * k], P.a[i], V.a[i]); }
        if (2 * k + 3 <= m) { mm(T, P, A2, d); P = T; }
      }
      mm(U, A, Us, d);
      for (int i = 0; i < d * d; ++i) { T.a[i] = V.a[i] - U.a[i]; E.a[i] = V.a[i] + U.a[i]; }
      return solve(T, E, d);
    }
  }
  int s = 0;
  if (nrm > kTheta[4]) {
    s = (int)ceil(log2(nrm / kTheta[4]));
    if (s < 0) s = 0;
    const double sc = ldexp(1.0, -s);
    for (int i = 0; i < d * d; ++i) A.a[i] *= sc;
  }
  const double* b = kPade13;
  Mat A4, A6;
  mm(A2, A, A, d);
  mm(A4, A2, A2, d);
  mm(A6, A4, A2, d);
  for (int i = 0; i < d * d; ++i) T.a[i] = fma(b[13], A6.a[i], fma(b[11], A4.a[i], b[9] * A2.a[i]));
  mm(V, A6, T, d);  // V used as scratch: A6 (b13 A6 + b11 A4 + b9 A2)
  for (int i = 0; i < d * d; ++i)
    V.a[i] += fma(b[7], A6.a[i], fma(b[5], A4.a[i], fma(b[3], A2.a[i], (i / d == i % d) ? b[1] : 0.0)));
  mm(U, A, V, d);
  for (int i = 0; i < d * d; ++i) T.a[i] = fma(b[12], A6.a[i], fma(b[10], A4.a[i], b[8] * A2.a[i]));
  mm(V, A6, T, d);
  for (int i = 0; i < d * d; ++i)
    V.a[i] += fma(b[6], A6.a[i], fma(b[4], A4.a[i], fma(b[2], A2.a[i], (i / d == i % d) ? b[0] : 0.0)));
  for (int i = 0; i < d * d; ++i) { T.a[i] = V.a[i] - U.a[i]; E.a[i] = V.a[i] + U.a[i]; }
  if (!solve(T, E, d)) return false;
  for (int k = 0; k < s; ++k) {
    mm(T, E, E, d);
    E = T;
  }
  return true;
}
}  // namespace

// A [n*n][N] (or shared [n*n]), Gamma [n*q][N] (or shared), W [q*q] shared, dt [N] (or one value) -> F, Q [n*n][N]
__global__ void __launch_bounds__(64)
van_loan_kernel(int n, int q, int64_t count, const double* __restrict__ A, int a_shared, const double* __restrict__ Gamma,
                int g_shared, const double* __restrict__ W, const double* __restrict__ dt, int dt_shared,
                double* __restrict__ F, double* __restrict__ Q, int32_t* __restrict__ status) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= count) return;
  const int d = 2 * n;
  const double h = dt_shared ? dt[0] : dt[tid];
  double Am[GKB_MAX_N * GKB_MAX_N], Gm[GKB_MAX_N * GKB_MAX_N], GW[GKB_MAX_N * GKB_MAX_N], GWG[GKB_MAX_N * GKB_MAX_N];
  for (int i = 0; i < n * n; ++i) Am[i] = a_shared ? A[i] : A[(int64_t)i * count + tid];
  for (int i = 0; i < n * q; ++i) Gm[i] = g_shared ? Gamma[i] : Gamma[(int64_t)i * count + tid];
  for (int i = 0; i < n; ++i)  // c2d.go:31-33: (Gamma W) Gamma^T, scaled by dt
    for (int j = 0; j < q; ++j) {
      double s = 0.0;
      for (int l = 0; l < q; ++l) s = fma(Gm[i * q + l], W[l * q + j], s);
      GW[i * q + j] = s;
    }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int l = 0; l < q; ++l) s = fma(GW[i * q + l], Gm[j * q + l], s);
      GWG[i * n + j] = s * h;
    }
  Mat M, E;
  for (int i = 0; i < d * d; ++i) M.a[i] = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {  // c2d.go:43-54
      M.a[i * d + j] = -(Am[i * n + j] * h);
      M.a[(i + n) * d + (j + n)] = Am[j * n + i] * h;
      M.a[i * d + (j + n)] = GWG[i * n + j];
    }
  const bool ok = expm(E, M, d);
  // c2d.go:62-74: F = (E_22)^T, Q = AsSymDense(F E_12): the upper triangle is kept
  double Fm[GKB_MAX_N * GKB_MAX_N];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Fm[i * n + j] = E.a[(n + j) * d + (n + i)];
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      double s = 0.0;
      for (int l = 0; l < n; ++l) s = fma(Fm[i * n + l], E.a[l * d + (n + j)], s);
      Q[(int64_t)(i * n + j) * count + tid] = s;
      Q[(int64_t)(j * n + i) * count + tid] = s;
    }
  for (int i = 0; i < n * n; ++i) F[(int64_t)i * count + tid] = Fm[i];
  if (status) status[tid] = ok ? 0 : GKB_ERR_NONFINITE;
}

int launch_van_loan(int n, int q, int64_t count, const double* A, int a_shared, const double* Gamma, int g_shared,
                    const double* W, const double* dt, int dt_shared, double* F, double* Q, int32_t* status, cudaStream_t s) {
  if (n < 1 || n > GKB_MAX_N || q < 1 || q > GKB_MAX_N) return GKB_ERR_UNSUPPORTED;
  van_loan_kernel<<<(unsigned)((count + 63) / 64), 64, 0, s>>>(n, q, count, A, a_shared, Gamma, g_shared, W, dt, dt_shared, F, Q, status);
  return 0;
}

}  // namespace gkb
