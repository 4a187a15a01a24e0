// kernels_noise.cu -- AWGN (noise.go:109-159) for ordinary filter handles: Noise.Process(k) / Noise.Measurement(k)
// samples generated on the device from the engine's counter-based stream.
//
// The reference's AWGN draws L z from a clock-seeded math/rand stream, a fresh sample on every call, ignoring k
// (noise.go:127-137).  Here the standard normals of (filter, step) are the Philox4x32-10 stream the Monte Carlo
// kernel uses for (trial, step) -- normal j of (seed, filter_offset + filter, step) -- so that
//   first  Process(k)      = chol(Q) z[0 : n]              (vanilla.go:146; the only one of a pure predictor / sqrt)
//          Measurement(k)  = chol(R) z[n : n + m]          (vanilla.go:157)
//   second Process(k)      = chol(Q) z[n + m : 2 n + m]    (vanilla.go:195: a different draw, as with the reference)
// and a pure predictor with this noise reproduces the truth trajectories of NewMonteCarloRuns sample for sample.
// The samples of one gkb_update call are written to the handle's replay arrays by this kernel and consumed by the
// ordinary replay path of the update kernels.
#include "engine_internal.h"
#include "philox.cuh"

namespace gkb {

struct AwgnParams {
  int n, m;
  unsigned long long seed;
  long long filter_offset;
  double LQ[GKB_MAX_N * GKB_MAX_N];
  double LR[GKB_MAX_M * GKB_MAX_M];
};

__global__ void __launch_bounds__(kThreads)
awgn_fill_kernel(const __grid_constant__ AwgnParams p, int64_t nf, int steps, int step0, double* __restrict__ w,
                 double* __restrict__ v, double* __restrict__ w2) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nf) return;
  const int n = p.n, m = p.m, count = 2 * n + m;
  const uint64_t gf = (uint64_t)(p.filter_offset + tid);
  for (int k = 0; k < steps; ++k) {
    double z[2 * GKB_MAX_N + GKB_MAX_M];
    for (int b = 0; 4 * b < count; ++b) {
      uint32_t o[4];
      philox4x32_10((uint32_t)gf, (uint32_t)(gf >> 32), (uint32_t)(step0 + k), (uint32_t)b, (uint32_t)p.seed,
                    (uint32_t)(p.seed >> 32), o);
      for (int i = 0; i < 4; ++i)
        if (4 * b + i < count) z[4 * b + i] = icdf_normal(o[i], kIcdfTable);
    }
    // the colouring is written exactly like the Monte Carlo kernel's (fma chain over j <= i), so the two agree bitwise
    for (int i = 0; i < n; ++i) {
      double s1 = 0.0, s2 = 0.0;
      for (int j = 0; j <= i; ++j) {
        s1 = fma(p.LQ[i * n + j], z[j], s1);
        s2 = fma(p.LQ[i * n + j], z[n + m + j], s2);
      }
      if (w) w[((int64_t)k * n + i) * nf + tid] = s1;
      if (w2) w2[((int64_t)k * n + i) * nf + tid] = s2;
    }
    for (int a = 0; a < m; ++a) {
      double s = 0.0;
      for (int b = 0; b <= a; ++b) s = fma(p.LR[a * m + b], z[n + b], s);
      if (v) v[((int64_t)k * m + a) * nf + tid] = s;
    }
  }
}

int launch_awgn_fill(int n, int m, const double* LQ, const double* LR, unsigned long long seed, long long filter_offset,
                     int64_t nf, int steps, int step0, double* w, double* v, double* w2, cudaStream_t s) {
  if (n < 1 || n > GKB_MAX_N || m < 1 || m > GKB_MAX_M) return GKB_ERR_UNSUPPORTED;
  AwgnParams p;
  p.n = n; p.m = m; p.seed = seed; p.filter_offset = filter_offset;
  for (int i = 0; i < GKB_MAX_N * GKB_MAX_N; ++i) p.LQ[i] = i < n * n ? LQ[i] : 0.0;
  for (int i = 0; i < GKB_MAX_M * GKB_MAX_M; ++i) p.LR[i] = i < m * m ? LR[i] : 0.0;
  awgn_fill_kernel<<<(unsigned)((nf + kThreads - 1) / kThreads), kThreads, 0, s>>>(p, nf, steps, step0, w, v, w2);
  return 0;
}

}  // namespace gkb
