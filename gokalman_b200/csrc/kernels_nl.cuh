// kernels_nl.cuh -- the general NLDKF kernel templates (HybridKF production / strict arithmetic, SRIF), shared by
// kernels_nl.cu (n <= 6) and kernels_nl_big.cu (n = 7, 8: the north star's "n <= 8", compiled in their own
// translation unit because the fully unrolled 8 x 8 code takes minutes to build).
#pragma once
#include "engine_internal.h"
#include "filters_nl.cuh"
#include "filters_strict.cuh"

namespace gkb {

template <int C>
GKB_DEV void nl_load(double (&dst)[C], const double* __restrict__ src, int shared, int64_t k, int64_t nf,
                     int64_t tid) {
  if (shared) {
#pragma unroll
    for (int i = 0; i < C; ++i) dst[i] = __ldg(src + k * C + i);
  } else {
    const double* p = src + k * C * nf + tid;
#pragma unroll
    for (int i = 0; i < C; ++i) dst[i] = __ldcs(p + (int64_t)i * nf);  // streamed once: evict-first
  }
}
template <int C>
GKB_DEV void nl_out(double* base, int k, int every_step, const double (&src)[C], int64_t nf, int64_t tid) {
  if (base == nullptr) return;
  double* dst = base + (every_step ? (int64_t)k * C * nf : 0) + tid;
#pragma unroll
  for (int i = 0; i < C; ++i) __stcs(dst + (int64_t)i * nf, src[i]);
}

// A failed epoch (the reference returns (nil, err)): NaN rows in the caller's output arrays, never stale data.
template <int N, int M, int INNOV>
GKB_DEV void nl_fail_outputs(const NlIo& io, int k, int64_t tid) {
  if (!(io.every_step || k == io.steps - 1)) return;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  auto fill = [&](double* base, int comps) {
    if (base == nullptr) return;
    double* dst = base + (io.every_step ? (int64_t)k * comps * io.nf : 0) + tid;
    for (int i = 0; i < comps; ++i) dst[(int64_t)i * io.nf] = qnan;
  };
  fill(io.o_state, N); fill(io.o_meas, M); fill(io.o_innov, INNOV); fill(io.o_obsdev, M); fill(io.o_gain, N * M);
  fill(io.o_covar, N * N); fill(io.o_pred, N * N);
}

template <int N, int M>
__global__ void __launch_bounds__(kThreads)
hybrid_run_kernel(const __grid_constant__ NlModel<N, M> md, const __grid_constant__ NlIo io) {
  constexpr int SN = N * (N + 1) / 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= io.nf) return;
  double x[N], P[SN];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = io.vec[(int64_t)i * io.nf + tid];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) P[sym_idx<N>(i, j)] = io.mat[(int64_t)(i * N + j) * io.nf + tid];
  int status = 0;
  for (int k = 0; k < io.steps; ++k) {
    const unsigned fl = io.flags ? io.flags[k] : (unsigned)GKB_F_MEAS;
    const bool has_meas = (fl & GKB_F_MEAS) != 0, ekf = (fl & GKB_F_EKF) != 0, snc = (fl & GKB_F_SNC) != 0;
    double Phi[N * N], Ht[M * N], ro[M], co[M];
    nl_load<N * N>(Phi, io.Phi, io.phi_shared, k, io.nf, tid);
    if (has_meas) {
      nl_load<M * N>(Ht, io.Htilde, io.h_shared, k, io.nf, tid);
      nl_load<M>(ro, io.real_obs, 0, k, io.nf, tid);
      nl_load<M>(co, io.computed_obs, 0, k, io.nf, tid);
    } else {
#pragma unroll
      for (int i = 0; i < M * N; ++i) Ht[i] = 0.0;
#pragma unroll
      for (int a = 0; a < M; ++a) { ro[a] = 0.0; co[a] = 0.0; }
    }
    NlOut<N, M> o;
    const double* Gk = (snc && io.Gamma) ? io.Gamma + (int64_t)k * N * md.q : nullptr;
    int err = hybrid_step<N, M>(md, x, P, Phi, Ht, ro, co, Gk, has_meas, ekf, snc, o);
    if (err != 0) {
      if (status == 0) status = err;
      nl_fail_outputs<N, M, M>(io, k, tid);
      continue;
    }
    if (io.every_step || k == io.steps - 1) {
      nl_out<N>(io.o_state, k, io.every_step, x, io.nf, tid);
      nl_out<M>(io.o_meas, k, io.every_step, ro, io.nf, tid);  // Measurement() = realObservation (hybrid.go:198)
      nl_out<M>(io.o_innov, k, io.every_step, o.innov, io.nf, tid);
      nl_out<M>(io.o_obsdev, k, io.every_step, o.obsdev, io.nf, tid);
      nl_out<N * M>(io.o_gain, k, io.every_step, o.K, io.nf, tid);
      if (io.o_covar != nullptr) {
        double full[N * N];
        sym_expand<N>(full, P);
        nl_out<N * N>(io.o_covar, k, io.every_step, full, io.nf, tid);
      }
      if (io.o_pred != nullptr) {
        double full[N * N];
        sym_expand<N>(full, o.Ppred);
        nl_out<N * N>(io.o_pred, k, io.every_step, full, io.nf, tid);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) io.vec[(int64_t)i * io.nf + tid] = x[i];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) io.mat[(int64_t)(i * N + j) * io.nf + tid] = P[sym_idx<N>(i, j)];
  if (io.status != nullptr && status != 0 && io.status[tid] == 0) io.status[tid] = status;
}

// The same run in REFERENCE-ORDER arithmetic (filters_strict.cuh: dense products in the written order, no FMA
// contraction, dense Joseph form, AsSymDense): the validation twin of hybrid_run_kernel, selected per handle with
// gkb_set_strict().  Every call shape of the general kernel is supported (every-step outputs, SNC, shared streams).
template <int N, int M>
__global__ void __launch_bounds__(kThreads)
hybrid_run_strict_kernel(const __grid_constant__ NlModel<N, M> md, const __grid_constant__ NlIo io) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= io.nf) return;
  double x[N], P[N * N];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = io.vec[(int64_t)i * io.nf + tid];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j)  // the stored matrix is the mirrored upper triangle (AsSymDense)
      P[i * N + j] = io.mat[(int64_t)((i <= j) ? (i * N + j) : (j * N + i)) * io.nf + tid];
  int status = 0;
  for (int k = 0; k < io.steps; ++k) {
    const unsigned fl = io.flags ? io.flags[k] : (unsigned)GKB_F_MEAS;
    const bool has_meas = (fl & GKB_F_MEAS) != 0, ekf = (fl & GKB_F_EKF) != 0, snc = (fl & GKB_F_SNC) != 0;
    double Phi[N * N], Ht[M * N], ro[M], co[M];
    nl_load<N * N>(Phi, io.Phi, io.phi_shared, k, io.nf, tid);
    if (has_meas) {
      nl_load<M * N>(Ht, io.Htilde, io.h_shared, k, io.nf, tid);
      nl_load<M>(ro, io.real_obs, 0, k, io.nf, tid);
      nl_load<M>(co, io.computed_obs, 0, k, io.nf, tid);
    } else {
#pragma unroll
      for (int i = 0; i < M * N; ++i) Ht[i] = 0.0;
#pragma unroll
      for (int a = 0; a < M; ++a) { ro[a] = 0.0; co[a] = 0.0; }
    }
    double Ppred[N * N], K[N * M], innov[M], obsdev[M];
    const double* Gk = (snc && io.Gamma) ? io.Gamma + (int64_t)k * N * md.q : nullptr;
    int err = strict::hybrid_step<N, M>(md, x, P, Phi, Ht, ro, co, Gk, has_meas, ekf, snc, Ppred, K, innov, obsdev);
    if (err != 0) {
      if (status == 0) status = err;
      nl_fail_outputs<N, M, M>(io, k, tid);
      continue;
    }
    if (io.every_step || k == io.steps - 1) {
      nl_out<N>(io.o_state, k, io.every_step, x, io.nf, tid);
      nl_out<M>(io.o_meas, k, io.every_step, ro, io.nf, tid);
      nl_out<M>(io.o_innov, k, io.every_step, innov, io.nf, tid);
      nl_out<M>(io.o_obsdev, k, io.every_step, obsdev, io.nf, tid);
      nl_out<N * M>(io.o_gain, k, io.every_step, K, io.nf, tid);
      nl_out<N * N>(io.o_covar, k, io.every_step, P, io.nf, tid);
      nl_out<N * N>(io.o_pred, k, io.every_step, Ppred, io.nf, tid);
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) io.vec[(int64_t)i * io.nf + tid] = x[i];
#pragma unroll
  for (int i = 0; i < N * N; ++i) io.mat[(int64_t)i * io.nf + tid] = P[i];
  if (io.status != nullptr && status != 0 && io.status[tid] == 0) io.status[tid] = status;
}

template <int N, int M>
__global__ void __launch_bounds__(kThreads)
srif_run_kernel(const __grid_constant__ NlModel<N, M> md, const __grid_constant__ NlIo io) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= io.nf) return;
  double b[N], R[N * N];
#pragma unroll
  for (int i = 0; i < N; ++i) b[i] = io.vec[(int64_t)i * io.nf + tid];
#pragma unroll
  for (int i = 0; i < N * N; ++i) R[i] = io.mat[(int64_t)i * io.nf + tid];
  int status = 0;
  for (int k = 0; k < io.steps; ++k) {
    const unsigned fl = io.flags ? io.flags[k] : (unsigned)GKB_F_MEAS;
    const bool has_meas = (fl & GKB_F_MEAS) != 0;
    double Phi[N * N], Ht[M * N], ro[M], co[M];
    nl_load<N * N>(Phi, io.Phi, io.phi_shared, k, io.nf, tid);
    if (has_meas) {
      nl_load<M * N>(Ht, io.Htilde, io.h_shared, k, io.nf, tid);
      nl_load<M>(ro, io.real_obs, 0, k, io.nf, tid);
      nl_load<M>(co, io.computed_obs, 0, k, io.nf, tid);
    } else {
#pragma unroll
      for (int i = 0; i < M * N; ++i) Ht[i] = 0.0;
#pragma unroll
      for (int a = 0; a < M; ++a) { ro[a] = 0.0; co[a] = 0.0; }
    }
    NlOut<N, M> o;
    int err = srif_step<N, M>(md, b, R, Phi, Ht, ro, co, has_meas, o);
    if (err != 0) {
      if (status == 0) status = err;
      nl_fail_outputs<N, M, N>(io, k, tid);
      continue;
    }
    if (io.every_step || k == io.steps - 1) {
      nl_out<M>(io.o_meas, k, io.every_step, ro, io.nf, tid);
      nl_out<N>(io.o_innov, k, io.every_step, b, io.nf, tid);  // Innovation() = b (srif.go:238-240)
      nl_out<M>(io.o_obsdev, k, io.every_step, o.obsdev, io.nf, tid);
      if (io.o_state != nullptr) {  // srif.go:223-235
        double xs[N];
        if (!srif_state<N>(xs, R, b)) {
          if (status == 0) status = GKB_ERR_SINGULAR_R;
#pragma unroll
          for (int i = 0; i < N; ++i) xs[i] = 0.0;
        }
        nl_out<N>(io.o_state, k, io.every_step, xs, io.nf, tid);
      }
      if (io.o_covar != nullptr) {  // srif.go:253-265
        double Pc[N * N];
        srif_covariance<N>(Pc, R);
        nl_out<N * N>(io.o_covar, k, io.every_step, Pc, io.nf, tid);
      }
      if (io.o_pred != nullptr) {  // srif.go:268-281
        double Pc[N * N];
        srif_covariance<N>(Pc, o.Rbar);
        nl_out<N * N>(io.o_pred, k, io.every_step, Pc, io.nf, tid);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) io.vec[(int64_t)i * io.nf + tid] = b[i];
#pragma unroll
  for (int i = 0; i < N * N; ++i) io.mat[(int64_t)i * io.nf + tid] = R[i];
  if (io.status != nullptr && status != 0 && io.status[tid] == 0) io.status[tid] = status;
}


// Launches the general kernels for one compiled shape (shared by the two dispatch tables).
template <int N, int M>
static int launch_nl_general(const HostModel& hm, const NlIo& io, cudaStream_t s) {
  const unsigned grid = (unsigned)((io.nf + kThreads - 1) / kThreads);
  NlModel<N, M> md;
  for (int i = 0; i < GKB_MAX_Q * GKB_MAX_Q; ++i) md.Q[i] = 0.0;
  for (int i = 0; i < hm.q * hm.q; ++i) md.Q[i] = hm.Q[i];
  for (int i = 0; i < M * M; ++i) { md.R[i] = hm.R[i]; md.L[i] = hm.L[i]; }
  md.q = hm.q;
  if (hm.kind == GKB_SRIF) {
    srif_run_kernel<N, M><<<grid, kThreads, 0, s>>>(md, io);
    return 0;
  }
  if (hm.kind != GKB_HYBRID) return GKB_ERR_UNSUPPORTED;
  if (io.strict) hybrid_run_strict_kernel<N, M><<<grid, kThreads, 0, s>>>(md, io);  // reference-order arithmetic (gkb_set_strict)
  else hybrid_run_kernel<N, M><<<grid, kThreads, 0, s>>>(md, io);
  return 0;
}

}  // namespace gkb
