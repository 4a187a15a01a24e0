// kernels_nl.cuh -- the general NLDKF kernel templates (HybridKF production / strict arithmetic, SRIF), shared by
// kernels_nl.cu (n <= 6) and kernels_nl_big.cu (n = 7, 8: the north star's "n <= 8", compiled in their own
// translation unit because the fully unrolled 8 x 8 code takes minutes to build).
#pragma once
#include <cstdlib>
#include <cstring>

#include "engine_internal.h"
#include "filters_nl.cuh"
#include "filters_strict.cuh"
#include "sched.cuh"

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
// L2 prefetch of the rows nl_load<C> will read for epoch k (per-filter streams only)
template <int C>
GKB_DEV void nl_prefetch_l2(const double* __restrict__ src, int shared, int64_t k, int64_t nf, int64_t tid) {
  if (shared || src == nullptr) return;
  const double* p = src + k * C * nf + tid;
#pragma unroll
  for (int i = 0; i < C; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (int64_t)i * nf));
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

// The run in REFERENCE-ORDER arithmetic (filters_strict.cuh: dense products in the written order, no FMA contraction,
// dense Joseph form, AsSymDense), selected per handle with gkb_set_strict(): the kernel that holds the north star's
// 1e-10 (at exactly 0 difference) on every configuration.  Covariance and work matrix live in lane-private
// shared-memory columns and the four dense products are rolled loops (strict::hybrid_sm_predict / hybrid_sm_update): a
// few KB of SASS instead of 56 KB, no spills.  Every call shape of the general kernel is supported (every-step outputs,
// SNC, shared streams).
//   * Phi of epoch k + 1 travels global -> shared with cp.async (8 bytes per entry into a lane-private stage) while epoch
//     k's update runs: it is requested once epoch k's Phi registers have been consumed (after P-bar) and collected at the
//     top of epoch k + 1 -- the registers cannot hold a second Phi next to I - K H (that version spilled and lost 25 %).
//   * Work distribution: like the production kernel's, a TASK is (chunk of the epochs, group of 32 filters); persistent
//     warps claim tasks from one atomic counter in chunk-major order and a group's state travels between chunks through
//     the handle's state arrays (release / acquire flag per group), so 3125 groups over 1184 resident warps no longer
//     leave a 12 % tail.  io.chunks == 0: one whole-run task per warp (grid = all groups).
#ifndef GKB_STRICT_TASKS_PER_WARP
#define GKB_STRICT_TASKS_PER_WARP 16.0
#endif
constexpr double kStrictTasksPerWarp = GKB_STRICT_TASKS_PER_WARP;  // scheduler granularity (measured: 16 beats 8 by 4 %)
using strict::kStrictMinBlocks;
using strict::kStrictThreads;
template <int N, int M>
constexpr size_t strict_smem_bytes() { return (size_t)(N * (N + 1) / 2 + 2 * N * N + N * M + 2 * M) * kStrictThreads * sizeof(double); }

GKB_DEV void nl_cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
GKB_DEV void nl_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
GKB_DEV void nl_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Epochs [k0, k1) of filter `tid` (one thread), state from / to the handle's arrays.
// Shared-memory columns of the thread: Ps (packed covariance), Ws (work matrix), Hs (M*N + 2M: the stage of H-tilde and the
// two observations between the end of one epoch and the update of the next, and K R during the Joseph products -- the two
// uses never overlap), Fs (the Phi stage).
template <int N, int M, bool SCHED>
GKB_DEV void strict_task(const NlModel<N, M>& md, const NlIo& io, double* Ps, double* Ws, double* Hs, double* Fs, int64_t tid,
                         int k0, int k1) {
  auto ld_state = [](const double* p) { return SCHED ? __ldcg(p) : *p; };  // SCHED: written by another SM -> read from L2
  const bool phi_staged = !io.phi_shared, h_staged = !io.h_shared;
  const int64_t nf = io.nf;
  auto flags_of = [&](int k) { return io.flags ? (unsigned)io.flags[k] : (unsigned)GKB_F_MEAS; };
  auto request_phi = [&](int k) {
    const double* p = io.Phi + (int64_t)k * N * N * nf + tid;
#pragma unroll
    for (int i = 0; i < N * N; ++i, p += nf) nl_cp_async8(&GKB_SM(Fs, i), p);
    nl_cp_async_commit();
  };
  auto request_meas = [&](int k) {  // H-tilde (unless shared by the batch) and the two observation vectors of epoch k
    if (h_staged) {
      const double* p = io.Htilde + (int64_t)k * M * N * nf + tid;
#pragma unroll
      for (int i = 0; i < M * N; ++i, p += nf) nl_cp_async8(&GKB_SM(Hs, i), p);
    }
    const double* pr = io.real_obs + (int64_t)k * M * nf + tid;
    const double* pc = io.computed_obs + (int64_t)k * M * nf + tid;
#pragma unroll
    for (int a = 0; a < M; ++a) {
      nl_cp_async8(&GKB_SM(Hs, M * N + a), pr + (int64_t)a * nf);
      nl_cp_async8(&GKB_SM(Hs, M * N + M + a), pc + (int64_t)a * nf);
    }
    nl_cp_async_commit();
  };
  unsigned fl = k0 < k1 ? flags_of(k0) : 0u;
  if (k0 < k1) {
    if (phi_staged) request_phi(k0);
    if ((fl & GKB_F_MEAS) != 0) request_meas(k0);
  }
  double x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = ld_state(io.vec + (int64_t)i * nf + tid);
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j)  // the stored matrix is the mirrored upper triangle (AsSymDense)
      GKB_SM(Ps, sym_idx<N>(i, j)) = ld_state(io.mat + (int64_t)(i * N + j) * nf + tid);
  int status = 0;
  for (int k = k0; k < k1; ++k) {
    const bool has_meas = (fl & GKB_F_MEAS) != 0, ekf = (fl & GKB_F_EKF) != 0, snc = (fl & GKB_F_SNC) != 0;
    const unsigned fl_next = (k + 1 < k1) ? flags_of(k + 1) : 0u;  // read an epoch ahead: its latency is never waited for
    double Phi[N * N];
    nl_cp_async_wait_all();  // this epoch's stages (requested an epoch ago)
    if (phi_staged) {
#pragma unroll
      for (int i = 0; i < N * N; ++i) Phi[i] = GKB_SM(Fs, i);
    } else {
      nl_load<N * N>(Phi, io.Phi, 1, k, nf, tid);
    }
    const double* Gk = (snc && io.Gamma) ? io.Gamma + (int64_t)k * N * md.q : nullptr;
    double xbar[N];
    strict::hybrid_sm_predict<N, M>(md, x, Ps, Ws, Phi, Gk, snc, ekf, xbar);
    if (phi_staged && k + 1 < k1) request_phi(k + 1);  // (the stage's values have all been used by now)
    double Ht[M * N], ro[M], co[M];
    if (has_meas) {
      if (h_staged) {
#pragma unroll
        for (int i = 0; i < M * N; ++i) Ht[i] = GKB_SM(Hs, i);
      } else {
        nl_load<M * N>(Ht, io.Htilde, 1, k, nf, tid);
      }
#pragma unroll
      for (int a = 0; a < M; ++a) { ro[a] = GKB_SM(Hs, M * N + a); co[a] = GKB_SM(Hs, M * N + M + a); }
    } else {
#pragma unroll
      for (int i = 0; i < M * N; ++i) Ht[i] = 0.0;
#pragma unroll
      for (int a = 0; a < M; ++a) { ro[a] = 0.0; co[a] = 0.0; }
    }
    const bool emit = io.every_step || k == io.steps - 1;
    double* pred_out = (emit && io.o_pred != nullptr) ? io.o_pred + (io.every_step ? (int64_t)k * N * N * nf : 0) + tid : nullptr;
    double K[N * M], innov[M], obsdev[M];
    const int err = strict::hybrid_sm_update<N, M>(md, x, Ws, Hs, xbar, Ht, ro, co, has_meas, ekf, pred_out, nf, K, innov, obsdev);
    if ((fl_next & GKB_F_MEAS) != 0) request_meas(k + 1);  // Hs is free again (K R has been consumed)
    fl = fl_next;
    if (err != 0) {  // the previous estimate (x, Ps) is untouched
      if (status == 0) status = err;
      nl_fail_outputs<N, M, M>(io, k, tid);
      continue;
    }
    strict::hybrid_sm_commit<N>(Ps, Ws);
    if (emit) {
      nl_out<N>(io.o_state, k, io.every_step, x, nf, tid);
      nl_out<M>(io.o_meas, k, io.every_step, ro, nf, tid);
      nl_out<M>(io.o_innov, k, io.every_step, innov, nf, tid);
      nl_out<M>(io.o_obsdev, k, io.every_step, obsdev, nf, tid);
      nl_out<N * M>(io.o_gain, k, io.every_step, K, nf, tid);
      if (io.o_covar != nullptr) {
        double* dst = io.o_covar + (io.every_step ? (int64_t)k * N * N * nf : 0) + tid;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
          for (int j = 0; j < N; ++j) __stcs(dst + (int64_t)(i * N + j) * nf, GKB_SM(Ps, sym_idx<N>(i, j)));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) io.vec[(int64_t)i * nf + tid] = x[i];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) io.mat[(int64_t)(i * N + j) * nf + tid] = GKB_SM(Ps, sym_idx<N>(i, j));
  if (io.status != nullptr && status != 0 && io.status[tid] == 0) io.status[tid] = status;
}

template <int N, int M>
__global__ void __launch_bounds__(kStrictThreads, (N <= 6 ? kStrictMinBlocks : 1))
hybrid_run_strict_kernel(const __grid_constant__ NlModel<N, M> md, const __grid_constant__ NlIo io) {
  extern __shared__ double gkb_strict_sm[];
  constexpr int SN = N * (N + 1) / 2;
  double* Ps = gkb_strict_sm + threadIdx.x;
  double* Ws = Ps + SN * kStrictThreads;
  double* Hs = Ws + N * N * kStrictThreads;
  double* Fs = Hs + (N * M + 2 * M) * kStrictThreads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = (int)((io.nf + 31) / 32);
  if (io.chunks <= 0) {  // one whole-run task per warp
    const int g = blockIdx.x * (kStrictThreads / 32) + warp;
    const int64_t tid = (int64_t)g * 32 + lane;
    if (tid < io.nf) strict_task<N, M, false>(md, io, Ps, Ws, Hs, Fs, tid, 0, io.steps);
    return;
  }
  const int n_tasks = groups * io.chunks;
  int c, g;
  while (sched_claim(io.sched, n_tasks, groups, lane, c, g)) {
    const int k0 = c * io.chunk_len, k1 = min(io.steps, k0 + io.chunk_len);
    if (c > 0) sched_acquire_group(io.sched + 1 + g, c, lane);
    const int64_t tid = (int64_t)g * 32 + lane;
    if (tid < io.nf) strict_task<N, M, true>(md, io, Ps, Ws, Hs, Fs, tid, k0, k1);
    if (c != io.chunks - 1) sched_release_group(io.sched + 1 + g, c + 1, lane);
  }
}

// (register version, kept for A/B: GKB_STRICT_PATH=regs)  The same run in REFERENCE-ORDER arithmetic (filters_strict.cuh: dense products in the written order, no FMA
// contraction, dense Joseph form, AsSymDense): the validation twin of hybrid_run_kernel, selected per handle with
// gkb_set_strict().  Every call shape of the general kernel is supported (every-step outputs, SNC, shared streams).
template <int N, int M>
__global__ void __launch_bounds__(kThreads)
hybrid_run_strict_regs_kernel(const __grid_constant__ NlModel<N, M> md, const __grid_constant__ NlIo io) {
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

// Out-of-line read-outs for n = 7, 8: State() / Covariance() each carry an unrolled 8 x 8 LU inverse; inlined into the
// epoch loop (they run only when an output row is due) they cost the loop its registers.  The arrays cross the call in
// local memory -- same arithmetic, same bits.
template <int N>
__device__ __noinline__ bool srif_state_ool(double* xs, const double* R, const double* b) {
  double x_[N], R_[N * N], b_[N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) R_[i] = R[i];
#pragma unroll
  for (int i = 0; i < N; ++i) b_[i] = b[i];
  const bool ok = srif_state<N>(x_, R_, b_);
#pragma unroll
  for (int i = 0; i < N; ++i) xs[i] = x_[i];
  return ok;
}
template <int N>
__device__ __noinline__ bool srif_covariance_ool(double* Pc, const double* R) {
  double P_[N * N], R_[N * N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) R_[i] = R[i];
  const bool ok = srif_covariance<N>(P_, R_);
#pragma unroll
  for (int i = 0; i < N * N; ++i) Pc[i] = P_[i];
  return ok;
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
        bool st_ok;
        if constexpr (N >= 7) {  // copies cross the call: R and b themselves stay in registers
          double Rc[N * N], bc[N];
#pragma unroll
          for (int i = 0; i < N * N; ++i) Rc[i] = R[i];
#pragma unroll
          for (int i = 0; i < N; ++i) bc[i] = b[i];
          st_ok = srif_state_ool<N>(xs, Rc, bc);
        } else {
          st_ok = srif_state<N>(xs, R, b);
        }
        if (!st_ok) {
          if (status == 0) status = GKB_ERR_SINGULAR_R;
#pragma unroll
          for (int i = 0; i < N; ++i) xs[i] = 0.0;
        }
        nl_out<N>(io.o_state, k, io.every_step, xs, io.nf, tid);
      }
      if (io.o_covar != nullptr) {  // srif.go:253-265
        double Pc[N * N];
        if constexpr (N >= 7) {
          double Rc[N * N];
#pragma unroll
          for (int i = 0; i < N * N; ++i) Rc[i] = R[i];
          srif_covariance_ool<N>(Pc, Rc);
        } else {
          srif_covariance<N>(Pc, R);
        }
        nl_out<N * N>(io.o_covar, k, io.every_step, Pc, io.nf, tid);
      }
      if (io.o_pred != nullptr) {  // srif.go:268-281
        double Pc[N * N];
        if constexpr (N >= 7) {
          double Rc[N * N];
#pragma unroll
          for (int i = 0; i < N * N; ++i) Rc[i] = o.Rbar[i];
          srif_covariance_ool<N>(Pc, Rc);
        } else {
          srif_covariance<N>(Pc, o.Rbar);
        }
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
  if (io.strict) {  // reference-order arithmetic (gkb_set_strict)
    const char* path = getenv("GKB_STRICT_PATH");  // =regs: the register version (A/B switch for tests, same bits)
    if (path != nullptr && !strcmp(path, "regs")) {
      hybrid_run_strict_regs_kernel<N, M><<<grid, kThreads, 0, s>>>(md, io);
      return 0;
    }
    constexpr size_t smem = strict_smem_bytes<N, M>();
    auto kern = hybrid_run_strict_kernel<N, M>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return GKB_ERR_CUDA;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int sms = 148, device = 0, per_sm = 1;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStrictThreads, smem);
    if (per_sm < 1) per_sm = 1;
    constexpr int kWarps = kStrictThreads / 32;
    const int64_t groups = (io.nf + 31) / 32;
    const int64_t slots = (int64_t)sms * per_sm * kWarps;
    NlIo io2 = io;
    int64_t ctas = (groups + kWarps - 1) / kWarps;
    bool forced = false;
    const int chunks = io.sched != nullptr ? sched_pick_chunks(groups, slots, io.steps, kStrictTasksPerWarp, &forced) : 1;
    const bool scheduled = io.sched != nullptr && (chunks > 1 || forced);
    sched_set_chunks(io2, chunks, scheduled);
    if (scheduled) {
      if (ctas > (int64_t)sms * per_sm) ctas = (int64_t)sms * per_sm;
      cudaMemsetAsync(io.sched, 0, sizeof(int) * (size_t)(groups + 1), s);
    }
    kern<<<(unsigned)ctas, kStrictThreads, smem, s>>>(md, io2);
  } else {
    hybrid_run_kernel<N, M><<<grid, kThreads, 0, s>>>(md, io);
  }
  return 0;
}

}  // namespace gkb
