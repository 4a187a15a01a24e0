// kernels_nl.cu -- batched NLDKF kernels: HybridKF (CKF / EKF, hybrid.go:104-204) and SRIF
// (srif.go:101-160), one filter per thread, time loop in the kernel.
//
// Per epoch each filter consumes its own Phi (n x n), Htilde (m x n), real and computed
// observation, streamed from HBM in SoA [step][component][filter]: a warp reads one coalesced
// 256-byte row per component.  Per-epoch flags (Predict vs Update, EKF on/off, SNC on/off) are
// shared by all filters, so the control flow is warp-uniform.
#include <cstring>

#include "kernels_nl.cuh"

namespace gkb {

// ---- SmoothAll (hybrid.go:209-238, srif.go:165-192) ---------------------------------------------------------
// Backward sweep over a stored history: for k = steps-2 .. 0:  S = inv(Phi_{k+1}),  x_k = S x_{k+1},
// P_k = (S P_{k+1}) S^T (upper triangle kept, AsSymDense).  x_{k+1} / P_{k+1} are the values the previous
// iteration wrote, so every estimate is the final one mapped back through the STMs -- exactly what the
// reference's in-place loop does.  One filter per thread; per (filter, step) the kernel reads one Phi and
// writes one state + covariance: HBM-bound streaming, no reuse.
// the rare path of smooth_all_kernel, kept out of line so that its predicated swaps do not cost the loop registers
template <int N>
__device__ __noinline__ int smooth_inverse_slow(double (&S)[N * N], const double* __restrict__ Phi, int phi_shared, int64_t k,
                                                int64_t nf, int64_t tid) {
  nl_load<N * N>(S, Phi, phi_shared, k, nf, tid);
  return inverse_lu<N>(S);
}

template <int N>
__global__ void __launch_bounds__(kThreads)
smooth_all_kernel(int64_t nf, int steps, const double* __restrict__ Phi, int phi_shared, double* __restrict__ xs,
                  double* __restrict__ Ps, int32_t* __restrict__ status) {
  constexpr int SN = N * (N + 1) / 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nf) return;
  double x[N], P[SN];
  {
    const double* xl = xs + (int64_t)(steps - 1) * N * nf + tid;
    const double* Pl = Ps + (int64_t)(steps - 1) * N * N * nf + tid;
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = xl[(int64_t)i * nf];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) P[sym_idx<N>(i, j)] = Pl[(int64_t)(i * N + j) * nf];
  }
  // Phi of the next iteration is requested before the arithmetic of this one (the loads are the only long-latency
  // operations of the loop), and the inverse first runs as straight-line code that speculates on "no row
  // interchange" (inverse_lu_nopivot); a warp with a lane for which that fails redoes the epoch's inverse with
  // the general routine on a re-read of Phi (an L2 hit).
  double Snext[N * N];
  if (steps >= 2) nl_load<N * N>(Snext, Phi, phi_shared, steps - 1, nf, tid);
  for (int k = steps - 2; k >= 0; --k) {
    double S[N * N];
#pragma unroll
    for (int i = 0; i < N * N; ++i) S[i] = Snext[i];
    if (k >= 1) nl_load<N * N>(Snext, Phi, phi_shared, k, nf, tid);
    bool ok = inverse_lu_nopivot<N>(S);
    if (!__all_sync(__activemask(), ok)) {
      double T[N * N];  // its address escapes to the out-of-line routine: a stack array, on this path only
      ok = smooth_inverse_slow<N>(T, Phi, phi_shared, k + 1, nf, tid) == 0;
#pragma unroll
      for (int i = 0; i < N * N; ++i) S[i] = T[i];
    }
    if (!ok) {  // "provided STM is not invertible": the reference stops here
      if (status != nullptr && status[tid] == 0) status[tid] = GKB_ERR_SINGULAR_PHI;
      return;
    }
    double xn[N];
    mulvec<N, N>(xn, S, x);
    double Pn[SN];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double row[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double acc = S[i * N] * P[sym_idx<N>(0, j)];
#pragma unroll
        for (int l = 1; l < N; ++l) acc = fma(S[i * N + l], P[sym_idx<N>(l, j)], acc);
        row[j] = acc;
      }
#pragma unroll
      for (int j = i; j < N; ++j) {
        double acc = row[0] * S[j * N];
#pragma unroll
        for (int l = 1; l < N; ++l) acc = fma(row[l], S[j * N + l], acc);
        Pn[sym_idx<N>(i, j)] = acc;
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = xn[i];
#pragma unroll
    for (int i = 0; i < SN; ++i) P[i] = Pn[i];
    double* xo = xs + (int64_t)k * N * nf + tid;
    double* Po = Ps + (int64_t)k * N * N * nf + tid;
#pragma unroll
    for (int i = 0; i < N; ++i) __stcs(xo + (int64_t)i * nf, x[i]);
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) __stcs(Po + (int64_t)(i * N + j) * nf, P[sym_idx<N>(i, j)]);
  }
}

int launch_smooth_all(int n, int64_t nf, int steps, const double* Phi, int phi_shared, double* xs, double* Ps,
                      int32_t* status, cudaStream_t s) {
  const unsigned grid = (unsigned)((nf + kThreads - 1) / kThreads);
  switch (n) {
#define GKB_SM(NN) case NN: smooth_all_kernel<NN><<<grid, kThreads, 0, s>>>(nf, steps, Phi, phi_shared, xs, Ps, status); return 0;
    GKB_SM(1) GKB_SM(2) GKB_SM(3) GKB_SM(4) GKB_SM(5) GKB_SM(6) GKB_SM(7) GKB_SM(8)
#undef GKB_SM
    default: return GKB_ERR_UNSUPPORTED;
  }
}

// ---- HouseholderTransf (helper.go:142-172), the reference's exported helper, on a batch of matrices --------
// A [(n+m)*(n+1)][count] (row-major components, matrix index fastest), transformed in place: one matrix per
// thread, the same register routine the SRIF measurement update (srif.go:298-340) uses.
template <int N, int M>
__global__ void __launch_bounds__(kThreads) householder_kernel(int64_t count, double* __restrict__ A) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= count) return;
  constexpr int LEN = (N + M) * (N + 1);
  double a[LEN];
#pragma unroll
  for (int i = 0; i < LEN; ++i) a[i] = A[(int64_t)i * count + tid];
  householder_transf<N, M>(a);
#pragma unroll
  for (int i = 0; i < LEN; ++i) A[(int64_t)i * count + tid] = a[i];
}

int launch_householder(int n, int m, int64_t count, double* A, cudaStream_t s) {
  const unsigned grid = (unsigned)((count + kThreads - 1) / kThreads);
#define GKB_CASE(NN, MM) \
  if (n == NN && m == MM) { householder_kernel<NN, MM><<<grid, kThreads, 0, s>>>(count, A); return 0; }
  GKB_FOR_EACH_LTI_SHAPE(GKB_CASE)
  GKB_CASE(2, 3)  // srif_test.go:31-56 (the reference's measurementSRIFUpdate known-answer test)
#undef GKB_CASE
  return GKB_ERR_UNSUPPORTED;
}

// ---- BatchKF (batch.go:34-79) ------------------------------------------------------------------------------
// One thread per batch filter: `steps` SetNextMeasurement accumulations Lambda += (H^T R) H, N += (H^T R) y
// (R, not its inverse: the reference's formula, batch.go:50) over the filter's measurement streams, then
// Solve(): P0 = inv(Lambda) by LU (upper triangle kept, AsSymDense), xHat0 = P0 N.  The full dense Lambda is
// accumulated in the reference's operation order.  Streams H [steps][m*n][N], observations [steps][m][N].
template <int N, int M>
// n = 7, 8 with m >= 2: Lambda alone is 49 / 64 doubles -- no register cap there (measured: x2.5 at (8,2), x1.5 at (8,3), x1.3 at
// (7,2); m = 1 keeps the cap: its thin stream wants the third CTA more than the registers, 0.86 without it)
__global__ void __launch_bounds__(kThreads, ((N <= 6 || M == 1) ? 3 : 2))
batch_solve_kernel(const __grid_constant__ NlModel<N, M> md, int64_t nf, int steps, const double* __restrict__ H,
                   int h_shared, const double* __restrict__ real_obs, const double* __restrict__ computed_obs,
                   double* __restrict__ xhat0, double* __restrict__ P0, int32_t* __restrict__ status) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nf) return;
  double Lam[N * N], Nv[N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) Lam[i] = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) Nv[i] = 0.0;
  // The loop is a pure stream (128 B per measurement at n = 6, m = 2): the next measurement is requested before the
  // arithmetic of the current one, and the register budget is capped (3 CTAs per SM) so that enough loads are in
  // flight -- the one LU inverse after the loop is what would otherwise claim 250 registers.
  double Hn[M * N], rn[M], cn[M];
  nl_load<M * N>(Hn, H, h_shared, 0, nf, tid);
  nl_load<M>(rn, real_obs, 0, 0, nf, tid);
  nl_load<M>(cn, computed_obs, 0, 0, nf, tid);
  for (int k = 0; k < steps; ++k) {
    double Hk[M * N], ro[M], co[M];
#pragma unroll
    for (int i = 0; i < M * N; ++i) Hk[i] = Hn[i];
#pragma unroll
    for (int a = 0; a < M; ++a) { ro[a] = rn[a]; co[a] = cn[a]; }
    if (k + 1 < steps) {
      nl_load<M * N>(Hn, H, h_shared, k + 1, nf, tid);
      nl_load<M>(rn, real_obs, 0, k + 1, nf, tid);
      nl_load<M>(cn, computed_obs, 0, k + 1, nf, tid);
    }
    double HtR[N * M];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int a = 0; a < M; ++a) {
        double s = Hk[i] * md.R[a];
#pragma unroll
        for (int b = 1; b < M; ++b) s = fma(Hk[b * N + i], md.R[b * M + a], s);
        HtR[i * M + a] = s;
      }
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = HtR[i * M] * Hk[j];
#pragma unroll
        for (int a = 1; a < M; ++a) s = fma(HtR[i * M + a], Hk[a * N + j], s);
        Lam[i * N + j] += s;
      }
      double s = HtR[i * M] * (ro[0] - co[0]);
#pragma unroll
      for (int a = 1; a < M; ++a) s = fma(HtR[i * M + a], ro[a] - co[a], s);
      Nv[i] += s;
    }
  }
  int st = 0;
  if (inverse_lu<N>(Lam) != 0) st = GKB_ERR_SINGULAR_S;  // batch.go:66-68 returns the error
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) Lam[i * N + j] = Lam[j * N + i];
  double x[N];
  mulvec<N, N>(x, Lam, Nv);
  if (st != 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = 0.0;
#pragma unroll
    for (int i = 0; i < N * N; ++i) Lam[i] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) xhat0[(int64_t)i * nf + tid] = x[i];
#pragma unroll
  for (int i = 0; i < N * N; ++i) P0[(int64_t)i * nf + tid] = Lam[i];
  if (status != nullptr) status[tid] = st;
}

template <int N, int M>
static int launch_batch_shape(const double* R_host, int64_t nf, int steps, const double* H, int h_shared,
                              const double* real_obs, const double* computed_obs, double* xhat0, double* P0,
                              int32_t* status, cudaStream_t s) {
  NlModel<N, M> md;
  memset(&md, 0, sizeof md);
  for (int i = 0; i < M * M; ++i) md.R[i] = R_host[i];
  const unsigned grid = (unsigned)((nf + kThreads - 1) / kThreads);
  batch_solve_kernel<N, M><<<grid, kThreads, 0, s>>>(md, nf, steps, H, h_shared, real_obs, computed_obs, xhat0, P0, status);
  return 0;
}

int launch_batch_solve(int n, int m, const double* R_host, int64_t nf, int steps, const double* H, int h_shared,
                       const double* real_obs, const double* computed_obs, double* xhat0, double* P0, int32_t* status,
                       cudaStream_t s) {
#define GKB_CASE(NN, MM) \
  if (n == NN && m == MM) return launch_batch_shape<NN, MM>(R_host, nf, steps, H, h_shared, real_obs, computed_obs, xhat0, P0, status, s);
  GKB_FOR_EACH_LTI_SHAPE(GKB_CASE)
#undef GKB_CASE
  return GKB_ERR_UNSUPPORTED;
}

template <int N, int M>
static int launch_nl_shape(const HostModel& hm, const NlIo& io, cudaStream_t s) {
  if (hm.kind != GKB_HYBRID && hm.kind != GKB_SRIF) return GKB_ERR_UNSUPPORTED;
  // production configuration (per-filter streams, final outputs only): the TMA kernels of kernels_nl_tma.cu
  // (gkb_set_strict: the hybrid's reference-order kernel / the SRIF's literal general epoch instead)
  if (!io.strict && launch_nl_tma(hm, io, s) == 0) return 0;
  return launch_nl_general<N, M>(hm, io, s);
}

int launch_nl_run(const HostModel& hm, const NlIo& io, cudaStream_t s) {
#define GKB_CASE(NN, MM) \
  if (hm.n == NN && hm.m == MM) return launch_nl_shape<NN, MM>(hm, io, s);
  GKB_FOR_EACH_SHAPE(GKB_CASE)
#undef GKB_CASE
  return launch_nl_run_big(hm, io, s);  // n = 7, 8 (kernels_nl_big.cu)
}

}  // namespace gkb
