// kernels_mc.cu -- fused Monte Carlo truth generation + filter + chi-square (NEES / NIS) reduction.
//
// Replaces NewMonteCarloRuns (montecarlo.go:92-119: samples x steps pure-predictor Vanilla updates,
// every Estimate retained) followed by NewChiSquare (chisquare.go:16-95: the tested filter re-run on
// each run's measurements, NEES and NIS per (run, step), per-step mean over runs).  Here one
// thread owns one trial: it advances the truth state, draws the AWGN samples from Philox keyed by
// (global trial, step), feeds the measurement to the tested filter held in registers and reduces
// NEES / NIS over the trials of the CTA per step; nothing is stored per (trial, step).
//
// Reduction: warp shuffle tree -> one shared-memory slot per (warp, step) -> per-CTA partial rows in
// global memory, flushed every kChunk steps -> a second tiny kernel sums the CTA rows in CTA order.
// No atomics: the result is bit-reproducible for a given (trials, grid).
#include "engine_internal.h"
#include "filters.cuh"
#include "filters_info_sqrt.cuh"
#include "philox.cuh"

namespace gkb {

constexpr int kChunk = kMcChunk;          // steps per shared-memory accumulation chunk
constexpr int kWarps = kThreads / 32;
constexpr int kMcMaxCtasPerSm = 8;

template <int N, int M>
struct McModel {
  double F[N * N];
  double G[N * GKB_MAX_C];
  double H[M * N];
  double LQ[N * N];
  double LR[M * M];
  double x0_truth[N];
  double x0_filter[N];
  double A0[N * N];  // initial P (vanilla), I (information) or S (sqrt) of the tested filter
  int c;
  int need_ctrl;
};

GKB_DEV double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Tested-filter adaptors: uniform interface over the three LDKF kinds.
template <int N, int M>
struct VanillaTested {
  using Model = VanillaModel<N, M>;
  double x[N];
  double P[N * (N + 1) / 2];
  GKB_DEV void init(const McModel<N, M>& mm) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = mm.x0_filter[i];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) P[sym_idx<N>(i, j)] = mm.A0[i * N + j];
  }
  // One Update(); returns NEES / NIS ingredients. chisquare.go:45-77.
  GKB_DEV int update(const Model& md, const double (&y)[M], const double (&gu)[N], const double (&xt)[N],
                     bool with_nees, bool with_nis, double& nees, double& nis) {
    double w0[N], v0[M];
#pragma unroll
    for (int i = 0; i < N; ++i) w0[i] = 0.0;  // the tested filter carries Noiseless(Q, R)
#pragma unroll
    for (int a = 0; a < M; ++a) v0[a] = 0.0;
    StepOut<N, M> o;
    int err = vanilla_step<N, M, false>(md, x, P, y, gu, w0, v0, o);
    if (err != 0) return err;
    if (with_nees) {
      double e[N];
#pragma unroll
      for (int i = 0; i < N; ++i) e[i] = xt[i] - x[i];
      bool ok;
      nees = spd_quadform<N>(P, e, ok);
    }
    if (with_nis) {
      // nu^T inv(H P- H^T + R) nu: the update already inverted exactly this matrix
      double t[M];
#pragma unroll
      for (int a = 0; a < M; ++a) {
        double s = o.Sinv[a * M] * o.innov[0];
#pragma unroll
        for (int b = 1; b < M; ++b) s = fma(o.Sinv[a * M + b], o.innov[b], s);
        t[a] = s;
      }
      double q = 0.0;
#pragma unroll
      for (int a = 0; a < M; ++a) q = fma(o.innov[a], t[a], q);
      nis = q;
    }
    return 0;
  }
};

template <int N, int M>
struct InfoTested {
  using Model = InfoModel<N, M>;
  double iv[N];
  double I[N * (N + 1) / 2];
  GKB_DEV void init(const McModel<N, M>& mm) {
    // NewInformationFromState (information.go:65-81): A0 already holds I0 = inv(P0), x0_filter holds i0
#pragma unroll
    for (int i = 0; i < N; ++i) iv[i] = mm.x0_filter[i];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) I[sym_idx<N>(i, j)] = mm.A0[i * N + j];
  }
  GKB_DEV int update(const Model& md, const double (&y)[M], const double (&gu)[N], const double (&xt)[N],
                     bool with_nees, bool with_nis, double& nees, double& nis) {
    double v0[M];
#pragma unroll
    for (int a = 0; a < M; ++a) v0[a] = 0.0;
    InfoOut<N, M> o;
    int err = info_step<N, M>(md, iv, I, y, gu, v0, o, /*want_yhat=*/false);
    if (err != 0) return err;
    if (with_nees) {
      // chisquare.go:51-58 with est.Covariance() = inv(I+) (zeros while singular) and est.State() = P i
      double Pc[N * N], xs[N], e[N], t[N];
      info_covariance<N>(Pc, I);
      mulvec<N, N>(xs, Pc, iv);
#pragma unroll
      for (int i = 0; i < N; ++i) e[i] = xt[i] - xs[i];
      (void)inverse_lu<N>(Pc);  // PInv.Inverse(est.Covariance()), error ignored
      mulvec<N, N>(t, Pc, e);
      double q = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) q = fma(e[i], t[i], q);
      nees = q;
    }
    if (with_nis) nis = 0.0;  // Innovation() is the n-vector i+: the reference's NIS product panics unless n == m
    return 0;
  }
};

template <int N, int M>
struct SqrtTested {
  using Model = SqrtModel<N, M>;
  double x[N];
  double S[N * N];
  GKB_DEV void init(const McModel<N, M>& mm) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = mm.x0_filter[i];
#pragma unroll
    for (int i = 0; i < N * N; ++i) S[i] = mm.A0[i];
  }
  GKB_DEV int update(const Model& md, const double (&y)[M], const double (&gu)[N], const double (&xt)[N],
                     bool with_nees, bool with_nis, double& nees, double& nis) {
    double w0[N], v0[M];
#pragma unroll
    for (int i = 0; i < N; ++i) w0[i] = 0.0;
#pragma unroll
    for (int a = 0; a < M; ++a) v0[a] = 0.0;
    SqrtOut<N, M> o;
    int err = sqrt_step<N, M>(md, x, S, y, gu, w0, v0, o);
    if (err != 0) return err;
    if (with_nees) {
      // P = S S^T with S lower triangular: e^T P^-1 e = |S^-1 e|^2 by forward substitution
      double z[N];
      double q = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double s = xt[i] - x[i];
#pragma unroll
        for (int l = 0; l < i; ++l) s = fma(-S[i * N + l], z[l], s);
        z[i] = s / S[i * N + i];
        q = fma(z[i], z[i], q);
      }
      nees = q;
    }
    if (with_nis) {
      // chisquare.go:67-76: Pyy = H PredCovariance H^T + R, PredCovariance = S- S-^T (squareroot.go:330-340)
      double B[M * N];  // H S-
#pragma unroll
      for (int a = 0; a < M; ++a)
#pragma unroll
        for (int j = 0; j < N; ++j) {
          double s = 0.0;
#pragma unroll
          for (int l = 0; l < N; ++l)
            if (l <= j) s = fma(md.H[a * N + l], o.Spred[l * N + j], s);
          B[a * N + j] = s;
        }
      double Pyy[M * M];
#pragma unroll
      for (int a = 0; a < M; ++a)
#pragma unroll
        for (int b = 0; b < M; ++b) {
          double s = 0.0;
#pragma unroll
          for (int l = 0; l < M; ++l) s = fma(md.sqrtR[a * M + l], md.sqrtR[b * M + l], s);  // R = sqrtR sqrtR^T
#pragma unroll
          for (int j = 0; j < N; ++j) s = fma(B[a * N + j], B[b * N + j], s);
          Pyy[a * M + b] = s;
        }
      (void)inverse_lu<M>(Pyy);
      double q = 0.0;
#pragma unroll
      for (int a = 0; a < M; ++a) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < M; ++b) s = fma(Pyy[a * M + b], o.innov[b], s);
        q = fma(o.innov[a], s, q);
      }
      nis = q;
    }
    return 0;
  }
};

template <int N, int M, class Tested>
__global__ void __launch_bounds__(kThreads)
mc_chisquare_kernel(const __grid_constant__ McModel<N, M> mm, const __grid_constant__ typename Tested::Model md,
                    const __grid_constant__ McIo io) {
  const int cols = kMcBaseCols + (io.want_xstats ? 3 * N : 0);
  extern __shared__ double acc[];  // [kWarps][kChunk][cols]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* wacc = acc + (size_t)warp * kChunk * cols;
  for (int i = threadIdx.x; i < kWarps * kChunk * cols; i += blockDim.x) acc[i] = 0.0;
  __syncthreads();
  double* prow = io.partial + (size_t)blockIdx.x * io.steps * cols;

  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < io.trials; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = base + threadIdx.x;
    const bool active = t < io.trials;
    const int64_t tl = active ? t : io.trials - 1;  // inactive lanes shadow the last trial, contribute 0
    const uint64_t gtrial = (uint64_t)(io.trial_offset + tl);
    double xt[N], xref[N];  // xref: the noise-free trajectory, the pivot of the Mean/StdDev sums
#pragma unroll
    for (int i = 0; i < N; ++i) xt[i] = xref[i] = mm.x0_truth[i];
    Tested kf;
    kf.init(mm);
    int status = 0;
    for (int k0 = 0; k0 < io.steps; k0 += kChunk) {
      const int kend = min(kChunk, io.steps - k0);
      for (int kk = 0; kk < kend; ++kk) {
        const int k = k0 + kk;
        // ---- AWGN samples: Process(k) then Measurement(k) (vanilla.go:146,157; noise.go:127-137)
        double w[N], v[M];
        if (io.noise_mode == GKB_NOISE_PHILOX) {
          double z[N + M];
          philox_normals<N + M>(io.seed, gtrial, (uint32_t)k, z);
#pragma unroll
          for (int i = 0; i < N; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j)
              if (j <= i) s = fma(mm.LQ[i * N + j], z[j], s);
            w[i] = s;
          }
#pragma unroll
          for (int a = 0; a < M; ++a) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < M; ++b)
              if (b <= a) s = fma(mm.LR[a * M + b], z[N + b], s);
            v[a] = s;
          }
        } else {
#pragma unroll
          for (int i = 0; i < N; ++i) w[i] = io.w[((int64_t)k * N + i) * io.trials + tl];
#pragma unroll
          for (int a = 0; a < M; ++a) v[a] = io.v[((int64_t)k * M + a) * io.trials + tl];
        }
        double gu[N];
#pragma unroll
        for (int i = 0; i < N; ++i) gu[i] = 0.0;
        if (mm.need_ctrl && io.u != nullptr) control_term<N>(gu, mm.G, mm.c, io.u + (int64_t)k * mm.c);
        // ---- truth: pure-predictor Vanilla.Update (vanilla.go:138-179): measurement from the
        //      PREVIOUS state, then the state advances (montecarlo.go:110-113)
        double yt[M];
#pragma unroll
        for (int a = 0; a < M; ++a) {
          double s = mm.H[a * N] * xt[0];
#pragma unroll
          for (int j = 1; j < N; ++j) s = fma(mm.H[a * N + j], xt[j], s);
          yt[a] = s + v[a];
        }
        {
          double xn[N];
#pragma unroll
          for (int i = 0; i < N; ++i) {
            double s = mm.F[i * N] * xt[0];
#pragma unroll
            for (int j = 1; j < N; ++j) s = fma(mm.F[i * N + j], xt[j], s);
            if (mm.need_ctrl) s += gu[i];
            xn[i] = s + w[i];
          }
#pragma unroll
          for (int i = 0; i < N; ++i) xt[i] = xn[i];
        }
        if (active) {
          if (io.truth_x) {
#pragma unroll
            for (int i = 0; i < N; ++i) io.truth_x[((int64_t)k * N + i) * io.trials + t] = xt[i];
          }
          if (io.truth_y) {
#pragma unroll
            for (int a = 0; a < M; ++a) io.truth_y[((int64_t)k * M + a) * io.trials + t] = yt[a];
          }
          if (io.noise_w) {
#pragma unroll
            for (int i = 0; i < N; ++i) io.noise_w[((int64_t)k * N + i) * io.trials + t] = w[i];
          }
          if (io.noise_v) {
#pragma unroll
            for (int a = 0; a < M; ++a) io.noise_v[((int64_t)k * M + a) * io.trials + t] = v[a];
          }
        }
        // ---- tested filter + chi-square samples (chisquare.go:39-77)
        double nees = 0.0, nis = 0.0;
        int err = kf.update(md, yt, gu, xt, io.with_nees != 0, io.with_nis != 0, nees, nis);
        if (err != 0) {
          if (status == 0) status = err;
          nees = 0.0;
          nis = 0.0;
        }
        if (!active) { nees = 0.0; nis = 0.0; }
        // ---- per-step reduction over the warp's trials
        double s_nis = warp_sum(nis), s_nees = warp_sum(nees);
        if (lane == 0) {
          wacc[kk * cols + 0] += s_nis;
          wacc[kk * cols + 1] += s_nees;
        }
        if (io.want_xstats) {
          // MonteCarloRuns.Mean / StdDev (montecarlo.go:18-59): sums of d = x - xref and d^2, where
          // xref is the noise-free trajectory (identical in every trial), so that the variance
          // (sum d^2 - (sum d)^2 / N) / (N - 1) does not cancel catastrophically.
          double xn[N];
#pragma unroll
          for (int i = 0; i < N; ++i) {
            double s = mm.F[i * N] * xref[0];
#pragma unroll
            for (int j = 1; j < N; ++j) s = fma(mm.F[i * N + j], xref[j], s);
            if (mm.need_ctrl) s += gu[i];
            xn[i] = s;
          }
#pragma unroll
          for (int i = 0; i < N; ++i) xref[i] = xn[i];
#pragma unroll
          for (int i = 0; i < N; ++i) {
            double d = active ? (xt[i] - xref[i]) : 0.0;
            double s1 = warp_sum(d), s2 = warp_sum(d * d);
            if (lane == 0) {
              wacc[kk * cols + kMcBaseCols + i] += s1;
              wacc[kk * cols + kMcBaseCols + N + i] += s2;
              if (base == 0 && warp == 0 && blockIdx.x == 0) wacc[kk * cols + kMcBaseCols + 2 * N + i] = xref[i];
            }
          }
        }
      }
      if (io.steps > kChunk) {  // flush this chunk into the CTA's partial row
        __syncthreads();
        for (int i = threadIdx.x; i < kend * cols; i += blockDim.x) {
          double s = 0.0;
#pragma unroll
          for (int wv = 0; wv < kWarps; ++wv) {
            s += acc[(size_t)wv * kChunk * cols + i];
            acc[(size_t)wv * kChunk * cols + i] = 0.0;
          }
          prow[(size_t)k0 * cols + i] += s;
        }
        __syncthreads();
      }
    }
    if (io.status != nullptr && active && status != 0) io.status[t] = status;
  }
  if (io.steps <= kChunk) {
    __syncthreads();
    for (int i = threadIdx.x; i < io.steps * cols; i += blockDim.x) {
      double s = 0.0;
#pragma unroll
      for (int wv = 0; wv < kWarps; ++wv) s += acc[(size_t)wv * kChunk * cols + i];
      prow[i] = s;
    }
  }
}

// out[col][k] = scale * sum_b partial[b][k][col], CTA rows added in CTA order.
__global__ void mc_finish_kernel(const double* __restrict__ partial, int grid, int steps, int cols, double scale,
                                 double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= steps * cols) return;
  double s = 0.0;
  for (int b = 0; b < grid; ++b) s += partial[(size_t)b * steps * cols + idx];
  const int k = idx / cols, col = idx % cols;
  out[(size_t)col * steps + k] = (col < kMcBaseCols) ? s * scale : s;  // only NIS / NEES become means
}

template <int N, int M>
static void fill_mc_model(const HostModel& hm, const McIo& io, McModel<N, M>& mm, const double* A0) {
  for (int i = 0; i < N * N; ++i) { mm.F[i] = hm.F[i]; mm.LQ[i] = io.LQ[i]; mm.A0[i] = A0[i]; }
  for (int i = 0; i < M * M; ++i) mm.LR[i] = io.LR[i];
  for (int i = 0; i < N * GKB_MAX_C; ++i) mm.G[i] = 0.0;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < hm.c; ++j) mm.G[i * hm.c + j] = hm.G[i * hm.c + j];
  for (int i = 0; i < M * N; ++i) mm.H[i] = hm.H[i];
  for (int i = 0; i < N; ++i) { mm.x0_truth[i] = io.x0_truth[i]; mm.x0_filter[i] = io.x0_filter[i]; }
  mm.c = hm.c;
  mm.need_ctrl = hm.need_ctrl;
}

template <class Kern>
static int pick_grid(Kern kern, size_t smem, int64_t trials, int device) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int sms = 148, per_sm = 1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > kMcMaxCtasPerSm) per_sm = kMcMaxCtasPerSm;
  // persistent CTAs: SM count x resident CTAs, no more than the work needs
  int64_t need = (trials + kThreads - 1) / kThreads;
  int64_t g = (int64_t)sms * per_sm;
  if (g > need) g = need;
  return (int)(g < 1 ? 1 : g);
}

template <int N, int M>
static int launch_mc_shape(const HostModel& hm, const McIo& io, int device, int* grid_out, cudaStream_t s) {
  const int cols = mc_cols(N, io.want_xstats);
  const size_t smem = sizeof(double) * kWarps * kChunk * cols;
  int grid = 1;
  McModel<N, M> mm;
  fill_mc_model<N, M>(hm, io, mm, io.P0);
  switch (hm.kind) {
    case GKB_VANILLA: {
      VanillaModel<N, M> md;
      for (int i = 0; i < N * N; ++i) { md.F[i] = hm.F[i]; md.Q[i] = hm.Q[i]; }
      for (int i = 0; i < M * M; ++i) md.R[i] = hm.R[i];
      for (int i = 0; i < N * GKB_MAX_C; ++i) md.G[i] = mm.G[i];
      for (int i = 0; i < M * N; ++i) md.H[i] = hm.H[i];
      md.c = hm.c; md.need_ctrl = hm.need_ctrl;
      auto kern = mc_chisquare_kernel<N, M, VanillaTested<N, M>>;
      *grid_out = grid = pick_grid(kern, smem, io.trials, device);
      kern<<<grid, kThreads, smem, s>>>(mm, md, io);
      return 0;
    }
    case GKB_INFORMATION: {
      InfoModel<N, M> md;
      for (int i = 0; i < N * N; ++i) { md.Finv[i] = hm.Finv[i]; md.Qinv[i] = hm.Qinv[i]; }
      for (int i = 0; i < M * M; ++i) md.Rinv[i] = hm.Rinv[i];
      md.rinv_dim = hm.rinv_dim;
      for (int i = 0; i < N * GKB_MAX_C; ++i) md.G[i] = mm.G[i];
      for (int i = 0; i < M * N; ++i) md.H[i] = hm.H[i];
      md.c = hm.c; md.need_ctrl = hm.need_ctrl;
      auto kern = mc_chisquare_kernel<N, M, InfoTested<N, M>>;
      *grid_out = grid = pick_grid(kern, smem, io.trials, device);
      kern<<<grid, kThreads, smem, s>>>(mm, md, io);
      return 0;
    }
    case GKB_SQRT: {
      SqrtModel<N, M> md;
      for (int i = 0; i < N * N; ++i) { md.F[i] = hm.F[i]; md.sqrtQ[i] = hm.sqrtQ[i]; }
      for (int i = 0; i < M * M; ++i) md.sqrtR[i] = hm.sqrtR[i];
      for (int i = 0; i < N * GKB_MAX_C; ++i) md.G[i] = mm.G[i];
      for (int i = 0; i < M * N; ++i) md.H[i] = hm.H[i];
      md.c = hm.c; md.need_ctrl = hm.need_ctrl;
      auto kern = mc_chisquare_kernel<N, M, SqrtTested<N, M>>;
      *grid_out = grid = pick_grid(kern, smem, io.trials, device);
      kern<<<grid, kThreads, smem, s>>>(mm, md, io);
      return 0;
    }
    default: return GKB_ERR_UNSUPPORTED;
  }
}

int mc_max_grid(int device) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return sms * kMcMaxCtasPerSm;
}

int launch_mc(const HostModel& hm, const McIo& io, int device, int* grid_out, cudaStream_t s) {
#define GKB_CASE(NN, MM) \
  if (hm.n == NN && hm.m == MM) return launch_mc_shape<NN, MM>(hm, io, device, grid_out, s);
  GKB_FOR_EACH_SHAPE(GKB_CASE)
#undef GKB_CASE
  return GKB_ERR_UNSUPPORTED;
}

int launch_mc_finish(const double* partial, int grid, int steps, int cols, double scale, double* out_cols,
                     cudaStream_t s) {
  const int total = steps * cols;
  mc_finish_kernel<<<(total + 255) / 256, 256, 0, s>>>(partial, grid, steps, cols, scale, out_cols);
  return 0;
}

}  // namespace gkb
