// kernels_mc.cuh -- fused Monte Carlo truth generation + filter + chi-square (NEES / NIS) reduction.
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
#pragma once
#ifndef GKB_MC_MIN_CTAS
// resident CTAs per SM the register allocator must leave room for (measured on B200 for n = 3:
// 5 CTAs of 128 threads = 96 registers is the fastest point, see DESIGN.md "tuning log")
#define GKB_MC_MIN_CTAS(n, m, lean) (!(lean) ? 1 : ((n) <= 3 && (m) == 1) ? 5 : (n) <= 3 ? 4 : (n) == 4 ? 3 : 2)
// The vanilla tested filter at n <= 3, m = 1 fits 72 registers without spills: 7 CTAs per SM run as fast as 5 on a full
// GPU (18.95 ms for 10^6 x 1000) and keep a 125 000-trial shard (one eighth of 10^6: strong scaling on
// 8 GPUs = 977 CTAs) resident in ONE wave of 148 x 7 = 1036 CTAs instead of 1.32 waves of 740.
#define GKB_MC_MIN_CTAS_VANILLA(n, m, lean) (((lean) && (n) <= 3 && (m) == 1) ? 7 : GKB_MC_MIN_CTAS(n, m, lean))
#endif
#ifndef GKB_MC_HOIST_MODEL
#define GKB_MC_HOIST_MODEL 0
#endif
#include "engine_internal.h"
#include "filters.cuh"
#include "filters_info_sqrt.cuh"
#include "philox.cuh"
#include <type_traits>

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

// Sums two quantities over the warp with one shuffle tree: after the first exchange even lanes carry
// `a`, odd lanes carry `b`; on return lane 0 holds sum(a) and lane 1 holds sum(b).
GKB_DEV double warp_sum2(double a, double b, int lane) {
  const bool odd = lane & 1;
  double keep = odd ? b : a;
  const double give = odd ? a : b;
  keep += __shfl_xor_sync(0xffffffffu, give, 1);
#pragma unroll
  for (int off = 2; off < 32; off <<= 1) keep += __shfl_xor_sync(0xffffffffu, keep, off);
  return keep;
}

template <int K>
GKB_DEV void opaque(double (&a)[K]) {  // pins values in registers: the compiler cannot re-derive them
#pragma unroll
  for (int i = 0; i < K; ++i) asm volatile("mov.b64 %0, %0;" : "+d"(a[i]));
}

GKB_DEV double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Tested-filter adaptors: uniform interface over the three LDKF kinds.
template <int N, int M>
struct VanillaTested {
  using Model = VanillaModel<N, M>;
  static constexpr int min_ctas(bool lean) { return GKB_MC_MIN_CTAS_VANILLA(N, M, lean); }
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
    double w0[N], v0[M];  // the tested filter carries Noiseless(Q, R): never read (NOISY = false)
#pragma unroll
    for (int i = 0; i < N; ++i) w0[i] = 0.0;
#pragma unroll
    for (int a = 0; a < M; ++a) v0[a] = 0.0;
    StepOut<N, M> o;
    int err = vanilla_step<N, M, false, /*NOISY=*/false, /*CHECK=*/false>(md, x, P, y, gu, w0, v0, o);
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
  static constexpr int min_ctas(bool lean) { return GKB_MC_MIN_CTAS(N, M, lean); }
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
      (void)inverse_lu_fast<N>(Pc);  // PInv.Inverse(est.Covariance()), error ignored
      mulvec<N, N>(t, Pc, e);
      double q = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) q = fma(e[i], t[i], q);
      nees = q;
    }
    if (with_nis) {
      // chisquare.go:61-77 with est.Innovation() = i+ (information.go:272-274): the product Pyy^-1 i+ only has
      // matching dimensions when n == m (the host rejects NIS for every other shape, as mat64 would panic)
      if constexpr (N == M) {
        double Pp[N * N], PHt[N * M], Pyy[M * M], t[M];
        info_covariance<N>(Pp, o.Ipred);  // PredCovariance() = inv(I-), zeros when not invertible
        mul_nt<N, N, M>(PHt, Pp, md.H);
#pragma unroll
        for (int a = 0; a < M; ++a)
#pragma unroll
          for (int b = 0; b < M; ++b) {
            double s2 = md.H[a * N] * PHt[b];
#pragma unroll
            for (int l = 1; l < N; ++l) s2 = fma(md.H[a * N + l], PHt[l * M + b], s2);
            Pyy[a * M + b] = s2 + md.R[a * M + b];
          }
        (void)inverse_lu<M>(Pyy);
        mulvec<M, M>(t, Pyy, iv);
        double q = 0.0;
#pragma unroll
        for (int a = 0; a < M; ++a) q = fma(iv[a], t[a], q);
        nis = q;
      } else {
        nis = 0.0;
      }
    }
    return 0;
  }
};

template <int N, int M>
struct SqrtTested {
  using Model = SqrtModel<N, M>;
  static constexpr int min_ctas(bool lean) { return GKB_MC_MIN_CTAS(N, M, lean); }
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
        z[i] = s * rcp_nr(S[i * N + i]);
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

// LEAN = the production configuration (Philox noise, no dumps, no Mean/StdDev sums, no per-trial
// status): the optional paths are compiled out so that the time loop stays small in the
// instruction cache.  Results are identical to the general instantiation.
// NOCTRL (LEAN only) = the run has no control term (io.gu == nullptr: no G, or all-zero controls as in
// montecarlo.go:98-104): the per-step control loads and selects are compiled out.
template <int N, int M, class Tested, bool LEAN, bool NOCTRL = false>
__global__ void __launch_bounds__(kThreads, Tested::min_ctas(LEAN))
mc_chisquare_kernel(const __grid_constant__ McModel<N, M> mm_c, const __grid_constant__ typename Tested::Model md_c,
                    const __grid_constant__ McIo io) {
#if GKB_MC_HOIST_MODEL
  McModel<N, M> mm = mm_c;
  typename Tested::Model md = md_c;
  opaque(mm.F); opaque(mm.H); opaque(mm.LQ); opaque(mm.LR);
  if constexpr (std::is_same<typename Tested::Model, VanillaModel<N, M>>::value) {
    opaque(md.F); opaque(md.H); opaque(md.Q); opaque(md.R);
  }
#else
  const McModel<N, M>& mm = mm_c;
  const typename Tested::Model& md = md_c;
#endif
  const int cols = kMcBaseCols + ((!LEAN && io.want_xstats) ? 3 * N : 0);
  extern __shared__ __align__(16) double smem_mc[];  // [icdf table][kWarps][kChunk][cols]
  double* icdf_tab = smem_mc;
  double* acc = smem_mc + kIcdfSegments * kIcdfCoefs;
  icdf_load(icdf_tab);
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
    PhiloxTrial pt;
    if (LEAN && N + M <= 4) pt.init(io.seed, gtrial);
    for (int k0 = 0; k0 < io.steps; k0 += kChunk) {
      const int kend = min(kChunk, io.steps - k0);
      for (int kk = 0; kk < kend; ++kk) {
        const int k = k0 + kk;
        // ---- AWGN samples: Process(k) then Measurement(k) (vanilla.go:146,157; noise.go:127-137)
        double w[N], v[M];
        if (LEAN || io.noise_mode == GKB_NOISE_PHILOX) {
          double z[N + M];
          if constexpr (LEAN && N + M <= 4) philox_normals_trial<N + M>(icdf_tab, pt, io.seed, (uint32_t)k, z);
          else philox_normals<N + M>(icdf_tab, io.seed, gtrial, (uint32_t)k, z);
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
        const bool ctrl = !(LEAN && NOCTRL) && mm.need_ctrl && io.gu != nullptr;  // uniform
        if (ctrl) {
#pragma unroll
          for (int i = 0; i < N; ++i) gu[i] = __ldg(io.gu + (int64_t)k * N + i);
        }
        // the tested filter's own control term (its G may differ from the truth's; LEAN runs share one stream)
        double guf[N];
#pragma unroll
        for (int i = 0; i < N; ++i) guf[i] = gu[i];
        if (!LEAN && io.gu_f != io.gu) {
#pragma unroll
          for (int i = 0; i < N; ++i) guf[i] = io.gu_f ? __ldg(io.gu_f + (int64_t)k * N + i) : 0.0;
        }
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
            if (ctrl) s += gu[i];
            xn[i] = s + w[i];
          }
#pragma unroll
          for (int i = 0; i < N; ++i) xt[i] = xn[i];
        }
        if (!LEAN && active) {
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
        int err = kf.update(md, yt, guf, xt, io.with_nees != 0, io.with_nis != 0, nees, nis);
        if (err != 0) {
          if (status == 0) status = err;
          nees = 0.0;
          nis = 0.0;
        }
        if (!active) { nees = 0.0; nis = 0.0; }
        // ---- per-step reduction over the warp's trials
        const double s2 = warp_sum2(nis, nees, lane);
        if (lane < 2) wacc[kk * cols + lane] += s2;
        if (!LEAN && io.want_xstats) {
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
    if (active && status != 0) {
      if (io.first_error != nullptr) atomicMin(io.first_error, status);  // rare: any failed Update (chisquare.go:40-42 panics)
      if (!LEAN && io.status != nullptr) io.status[t] = status;
    }
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

}  // namespace gkb
