// kernels_lti.cu -- batched LDKF.Update kernels: one filter per thread, time loop in the kernel.
//
// Replaces the per-call bodies of Vanilla.Update (vanilla.go:128-220), Information.Update
// (information.go:153-227) and SquareRoot.Update (squareroot.go:129-274).  State and covariance
// stay in FP64 registers across steps; measurements stream in SoA [step][component][filter] so a
// warp's loads are one coalesced 256-byte row per component; the model is a by-value kernel
// parameter (constant bank).
#include "engine_internal.h"
#include "filters.cuh"
#include "filters_info_sqrt.cuh"

namespace gkb {

template <int C>
GKB_DEV void load_soa(double (&dst)[C], const double* __restrict__ src, int64_t nf, int64_t tid) {
#pragma unroll
  for (int i = 0; i < C; ++i) dst[i] = src[(int64_t)i * nf + tid];
}
template <int C>
GKB_DEV void store_soa(double* __restrict__ dst, const double (&src)[C], int64_t nf, int64_t tid) {
#pragma unroll
  for (int i = 0; i < C; ++i) dst[(int64_t)i * nf + tid] = src[i];
}
template <int C>
GKB_DEV void write_out(double* base, int k, int every_step, const double (&src)[C], int64_t nf, int64_t tid) {
  if (base == nullptr) return;
  double* dst = base + (every_step ? (int64_t)k * C * nf : 0);
  store_soa<C>(dst, src, nf, tid);
}

// A failed Update returns (nil, err) in the reference: the rows of that (filter, step) in the caller's output arrays
// are filled with NaN (never left holding data of an earlier call), the filter keeps its previous estimate and the
// first error goes to the status array.
template <int C>
GKB_DEV void write_nan(double* base, int k, int every_step, int64_t nf, int64_t tid) {
  if (base == nullptr) return;
  double* dst = base + (every_step ? (int64_t)k * C * nf : 0);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
  for (int i = 0; i < C; ++i) dst[(int64_t)i * nf + tid] = qnan;
}
template <int N, int M, int INNOV>
GKB_DEV void fail_outputs(const LtiIo& io, int k, int64_t tid) {
  if (!(io.every_step || k == io.steps - 1)) return;
  write_nan<N>(io.o_state, k, io.every_step, io.nf, tid);
  write_nan<M>(io.o_meas, k, io.every_step, io.nf, tid);
  write_nan<INNOV>(io.o_innov, k, io.every_step, io.nf, tid);
  write_nan<N * M>(io.o_gain, k, io.every_step, io.nf, tid);
  write_nan<N * N>(io.o_covar, k, io.every_step, io.nf, tid);
  write_nan<N * N>(io.o_pred, k, io.every_step, io.nf, tid);
}

template <int N, int M>
GKB_DEV void load_inputs(const LtiIo& io, int k, int64_t tid, double (&y)[M], double (&w)[N], double (&v)[M],
                         int& err) {
  if (io.y_shared) {
#pragma unroll
    for (int a = 0; a < M; ++a) y[a] = __ldg(io.y + (int64_t)k * M + a);
  } else {
    load_soa<M>(y, io.y + (int64_t)k * M * io.nf, io.nf, tid);
  }
  const int ks = io.step0 + k;
  if (io.w != nullptr) {
    if (ks < io.replay_steps) load_soa<N>(w, io.w + (int64_t)ks * N * io.nf, io.nf, tid);
    else err = GKB_ERR_NOISE_RANGE;
  }
  if (io.v != nullptr) {
    if (ks < io.replay_steps) load_soa<M>(v, io.v + (int64_t)ks * M * io.nf, io.nf, tid);
    else err = GKB_ERR_NOISE_RANGE;
  }
}

// ---- Vanilla / pure predictor -------------------------------------------------------------------
template <int N, int M, bool PREDICTOR>
__global__ void __launch_bounds__(kThreads)
vanilla_update_kernel(const __grid_constant__ VanillaModel<N, M> md, const __grid_constant__ LtiIo io) {
  constexpr int SN = N * (N + 1) / 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= io.nf) return;
  double x[N], P[SN];
  load_soa<N>(x, io.vec, io.nf, tid);
  {
    double full[N * N];
    load_soa<N * N>(full, io.mat, io.nf, tid);
    sym_pack_upper<N>(P, full);
  }
  int status = 0;
  for (int k = 0; k < io.steps; ++k) {
    double y[M], w[N], v[M], gu[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { w[i] = 0.0; gu[i] = 0.0; }
#pragma unroll
    for (int a = 0; a < M; ++a) v[a] = 0.0;
    int err = 0;
    load_inputs<N, M>(io, k, tid, y, w, v, err);
    if (md.need_ctrl && io.gu != nullptr) {
#pragma unroll
      for (int i = 0; i < N; ++i) gu[i] = __ldg(io.gu + (int64_t)k * N + i);
    }
    StepOut<N, M> o;
    // AWGN: the second Process(k) call of the step draws afresh (noise.go:127-131); replay / noiseless: the same vector.
    // ONE call site: two inlined copies of the step made ptxas give up on the n = 8, m = 3 shape (32 registers, 38 KB spill).
    double w2[N];
    if (io.w2 != nullptr) {
      load_soa<N>(w2, io.w2 + (int64_t)(io.step0 + k) * N * io.nf, io.nf, tid);
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) w2[i] = w[i];
    }
    if (err == 0) err = vanilla_step<N, M, PREDICTOR>(md, x, P, y, gu, w, v, o, w2);
    if (err != 0) {
      if (status == 0) status = err;
      fail_outputs<N, M, M>(io, k, tid);
      continue;  // like the reference, a failed Update leaves the previous estimate in place
    }
    if (io.every_step || k == io.steps - 1) {
      write_out<N>(io.o_state, k, io.every_step, x, io.nf, tid);
      write_out<M>(io.o_meas, k, io.every_step, o.yhat, io.nf, tid);
      write_out<M>(io.o_innov, k, io.every_step, o.innov, io.nf, tid);
      write_out<N * M>(io.o_gain, k, io.every_step, o.K, io.nf, tid);
      if (io.o_covar != nullptr) {
        double full[N * N];
        sym_expand<N>(full, P);
        write_out<N * N>(io.o_covar, k, io.every_step, full, io.nf, tid);
      }
      if (io.o_pred != nullptr) {
        double full[N * N];
        sym_expand<N>(full, o.Ppred);
        write_out<N * N>(io.o_pred, k, io.every_step, full, io.nf, tid);
      }
    }
  }
  store_soa<N>(io.vec, x, io.nf, tid);
  {
    double full[N * N];
    sym_expand<N>(full, P);
    store_soa<N * N>(io.mat, full, io.nf, tid);
  }
  if (io.status != nullptr && status != 0 && io.status[tid] == 0) io.status[tid] = status;
}

// ---- Information ----------------------------------------------------------------------------------
template <int N, int M>
__global__ void __launch_bounds__(kThreads)
info_update_kernel(const __grid_constant__ InfoModel<N, M> md, const __grid_constant__ LtiIo io) {
  constexpr int SN = N * (N + 1) / 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= io.nf) return;
  double iv[N], I[SN];
  load_soa<N>(iv, io.vec, io.nf, tid);
  {
    double full[N * N];
    load_soa<N * N>(full, io.mat, io.nf, tid);
    sym_pack_upper<N>(I, full);
  }
  int status = 0;
  for (int k = 0; k < io.steps; ++k) {
    double y[M], w[N], v[M], gu[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { w[i] = 0.0; gu[i] = 0.0; }
#pragma unroll
    for (int a = 0; a < M; ++a) v[a] = 0.0;
    int err = 0;
    load_inputs<N, M>(io, k, tid, y, w, v, err);
    if (md.need_ctrl && io.gu != nullptr) {
#pragma unroll
      for (int i = 0; i < N; ++i) gu[i] = __ldg(io.gu + (int64_t)k * N + i);
    }
    InfoOut<N, M> o;
    if (err == 0) err = info_step<N, M>(md, iv, I, y, gu, v, o);
    if (err != 0) {
      if (status == 0) status = err;
      fail_outputs<N, M, N>(io, k, tid);
      continue;
    }
    if (io.every_step || k == io.steps - 1) {
      const bool want_post = io.o_state != nullptr || io.o_covar != nullptr;
      write_out<M>(io.o_meas, k, io.every_step, o.yhat, io.nf, tid);
      write_out<N>(io.o_innov, k, io.every_step, iv, io.nf, tid);  // Innovation() = i+ (information.go:272-274)
      if (want_post) {
        double Pc[N * N], xs[N];
        info_covariance<N>(Pc, I);  // information.go:276-293
        double ivv[N];
#pragma unroll
        for (int i = 0; i < N; ++i) ivv[i] = iv[i];
        mulvec<N, N>(xs, Pc, ivv);
        write_out<N>(io.o_state, k, io.every_step, xs, io.nf, tid);
        write_out<N * N>(io.o_covar, k, io.every_step, Pc, io.nf, tid);
      }
      if (io.o_pred != nullptr) {
        double Pc[N * N];
        info_covariance<N>(Pc, o.Ipred);
        write_out<N * N>(io.o_pred, k, io.every_step, Pc, io.nf, tid);
      }
    }
  }
  store_soa<N>(io.vec, iv, io.nf, tid);
  {
    double full[N * N];
    sym_expand<N>(full, I);
    store_soa<N * N>(io.mat, full, io.nf, tid);
  }
  if (io.status != nullptr && status != 0 && io.status[tid] == 0) io.status[tid] = status;
}

// ---- Square root ------------------------------------------------------------------------------------
template <int N, int M>
__global__ void __launch_bounds__(kThreads)
sqrt_update_kernel(const __grid_constant__ SqrtModel<N, M> md, const __grid_constant__ LtiIo io) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= io.nf) return;
  double x[N], S[N * N];
  load_soa<N>(x, io.vec, io.nf, tid);
  load_soa<N * N>(S, io.mat, io.nf, tid);
  int status = 0;
  for (int k = 0; k < io.steps; ++k) {
    double y[M], w[N], v[M], gu[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { w[i] = 0.0; gu[i] = 0.0; }
#pragma unroll
    for (int a = 0; a < M; ++a) v[a] = 0.0;
    int err = 0;
    load_inputs<N, M>(io, k, tid, y, w, v, err);
    if (md.need_ctrl && io.gu != nullptr) {
#pragma unroll
      for (int i = 0; i < N; ++i) gu[i] = __ldg(io.gu + (int64_t)k * N + i);
    }
    SqrtOut<N, M> o;
    if (err == 0) err = sqrt_step<N, M>(md, x, S, y, gu, w, v, o);
    if (err != 0) {
      if (status == 0) status = err;
      fail_outputs<N, M, M>(io, k, tid);
      continue;
    }
    if (io.every_step || k == io.steps - 1) {
      write_out<N>(io.o_state, k, io.every_step, x, io.nf, tid);
      write_out<M>(io.o_meas, k, io.every_step, o.yhat, io.nf, tid);
      write_out<M>(io.o_innov, k, io.every_step, o.innov, io.nf, tid);
      write_out<N * M>(io.o_gain, k, io.every_step, o.K, io.nf, tid);
      if (io.o_covar != nullptr) {  // squareroot.go:317-327: P = S S^T
        double Pc[N * N];
        mul_nt<N, N, N>(Pc, S, S);
        write_out<N * N>(io.o_covar, k, io.every_step, Pc, io.nf, tid);
      }
      if (io.o_pred != nullptr) {  // squareroot.go:330-340: P- = S- S-^T
        double Pc[N * N];
        mul_nt<N, N, N>(Pc, o.Spred, o.Spred);
        write_out<N * N>(io.o_pred, k, io.every_step, Pc, io.nf, tid);
      }
    }
  }
  store_soa<N>(io.vec, x, io.nf, tid);
  store_soa<N * N>(io.mat, S, io.nf, tid);
  if (io.status != nullptr && status != 0 && io.status[tid] == 0) io.status[tid] = status;
}

// ---- control term: gu[k][i] = sum_j G[i][j] u[k][j], once per step for the whole batch ----------------
#if GKB_LTI_PART == 0
struct GuParams { double G[GKB_MAX_N * GKB_MAX_C]; };
__global__ void gu_kernel_p(const __grid_constant__ GuParams p, int n, int c, const double* __restrict__ u, int steps,
                            double* __restrict__ gu) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= steps * n) return;
  const int k = idx / n, i = idx % n;
  double s = 0.0;
  for (int j = 0; j < c; ++j) s = fma(p.G[i * c + j], u[(int64_t)k * c + j], s);
  gu[idx] = s;
}

int launch_gu(const double* G_host, int n, int c, const double* u_dev, int steps, double* gu_dev, cudaStream_t s) {
  GuParams p;
  for (int i = 0; i < GKB_MAX_N * GKB_MAX_C; ++i) p.G[i] = 0.0;
  for (int i = 0; i < n * c; ++i) p.G[i] = G_host[i];
  const int total = steps * n;
  gu_kernel_p<<<(total + 127) / 128, 128, 0, s>>>(p, n, c, u_dev, steps, gu_dev);
  return 0;
}
#endif

// ---- model marshalling + dispatch -------------------------------------------------------------------
template <int N, int M>
static void fill_common(const HostModel& hm, double* G, double* H, int& c, int& need_ctrl) {
  for (int i = 0; i < N * GKB_MAX_C; ++i) G[i] = 0.0;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < hm.c; ++j) G[i * hm.c + j] = hm.G[i * hm.c + j];
  for (int i = 0; i < M * N; ++i) H[i] = hm.H[i];
  c = hm.c;
  need_ctrl = hm.need_ctrl;
}

template <int N, int M>
static int launch_shape(const HostModel& hm, const LtiIo& io, cudaStream_t s) {
  const unsigned grid = (unsigned)((io.nf + kThreads - 1) / kThreads);
  switch (hm.kind) {
    case GKB_VANILLA:
    case GKB_PREDICTOR: {
      if (hm.m_r != M) return GKB_ERR_DIMS;
      VanillaModel<N, M> md;
      for (int i = 0; i < N * N; ++i) { md.F[i] = hm.F[i]; md.Q[i] = hm.Q[i]; }
      for (int i = 0; i < M * M; ++i) md.R[i] = hm.R[i];
      fill_common<N, M>(hm, md.G, md.H, md.c, md.need_ctrl);
      if (hm.kind == GKB_VANILLA) vanilla_update_kernel<N, M, false><<<grid, kThreads, 0, s>>>(md, io);
      else vanilla_update_kernel<N, M, true><<<grid, kThreads, 0, s>>>(md, io);
      return 0;
    }
    case GKB_INFORMATION: {
      if (hm.rinv_dim != 1 && hm.rinv_dim != M) return GKB_ERR_DIMS;
      InfoModel<N, M> md;
      for (int i = 0; i < N * N; ++i) { md.Finv[i] = hm.Finv[i]; md.Qinv[i] = hm.Qinv[i]; }
      for (int i = 0; i < M * M; ++i) md.Rinv[i] = 0.0;
      if (hm.rinv_dim == 1) md.Rinv[0] = hm.Rinv[0];
      else for (int i = 0; i < M * M; ++i) md.Rinv[i] = hm.Rinv[i];
      md.rinv_dim = hm.rinv_dim;
      for (int i = 0; i < M * M; ++i) md.R[i] = hm.R[i];
      fill_common<N, M>(hm, md.G, md.H, md.c, md.need_ctrl);
      info_update_kernel<N, M><<<grid, kThreads, 0, s>>>(md, io);
      return 0;
    }
    case GKB_SQRT: {
      if (hm.m_r != M) return GKB_ERR_DIMS;
      SqrtModel<N, M> md;
      for (int i = 0; i < N * N; ++i) { md.F[i] = hm.F[i]; md.sqrtQ[i] = hm.sqrtQ[i]; }
      for (int i = 0; i < M * M; ++i) md.sqrtR[i] = hm.sqrtR[i];
      fill_common<N, M>(hm, md.G, md.H, md.c, md.need_ctrl);
      sqrt_update_kernel<N, M><<<grid, kThreads, 0, s>>>(md, io);
      return 0;
    }
    default: return GKB_ERR_UNSUPPORTED;
  }
}

#ifndef GKB_LTI_PART
#error "compile with -DGKB_LTI_PART=0 (n <= 6 + the control-term kernels) or 1 (n = 7, 8): see Makefile"
#endif
#if GKB_LTI_PART == 0
int launch_lti_update(const HostModel& hm, const LtiIo& io, cudaStream_t s) {
#define GKB_CASE(NN, MM) \
  if (hm.n == NN && hm.m == MM) return launch_shape<NN, MM>(hm, io, s);
  GKB_FOR_EACH_SHAPE(GKB_CASE)
#undef GKB_CASE
  return launch_lti_update_big(hm, io, s);
}
#else
int launch_lti_update_big(const HostModel& hm, const LtiIo& io, cudaStream_t s) {
#define GKB_CASE(NN, MM) \
  if (hm.n == NN && hm.m == MM) return launch_shape<NN, MM>(hm, io, s);
  GKB_FOR_EACH_BIG_SHAPE(GKB_CASE)
#undef GKB_CASE
  return GKB_ERR_UNSUPPORTED;
}
#endif

}  // namespace gkb
