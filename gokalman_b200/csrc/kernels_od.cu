// kernels_od.cu -- orbit-determination inputs on the device (od_synth.cuh) and the FUSED OD filter run.
//
//   od_synth_kernel   writes the per-filter Phi / Htilde / real / computed streams of gkb_nl_run ([step][component]
//                     [filter], coalesced streaming stores) from the filters' initial reference orbits: the stream
//                     generator of the bench / tests, and the "precompute once, filter many times" path.
//   od_run_kernel     the same synthesis FUSED with the hybrid CKF / EKF step (hybrid.go:104-204): every epoch's Phi,
//                     Htilde and observations are produced in the thread that consumes them and never touch HBM.
//                     Host traffic of a whole run: 48 B per filter in (initial orbit), the final estimate out.
// Both call the same non-inlined od_step, so the fused run is bit-identical to gkb_nl_run on the synthesised streams.
#include <cstring>

#include "engine_internal.h"
#include "filters_nl.cuh"
#include "filters_strict.cuh"
#include "od_synth.cuh"

namespace gkb {

__global__ void __launch_bounds__(kThreads)
od_synth_kernel(const __grid_constant__ OdParams c, int64_t nf, int steps, double* __restrict__ orbit,
                const double* __restrict__ station, const double* __restrict__ tobs, double* __restrict__ Phi,
                double* __restrict__ Ht, double* __restrict__ real_obs, double* __restrict__ comp_obs) {
  extern __shared__ __align__(16) double icdf_tab[];
  icdf_load(icdf_tab);
  __syncthreads();
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nf) return;
  double X[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) X[i] = orbit[(int64_t)i * nf + tid];
  const uint64_t gf = (uint64_t)(c.filter_offset + tid);
  for (int k = 0; k < steps; ++k) {
    double st[6], to[2], z[2], o[kOdRows];
#pragma unroll
    for (int i = 0; i < 6; ++i) st[i] = __ldg(station + (int64_t)k * 6 + i);
    to[0] = __ldg(tobs + (int64_t)k * 2);
    to[1] = __ldg(tobs + (int64_t)k * 2 + 1);
    philox_normals<2>(icdf_tab, c.seed, gf, (uint32_t)k, z);
    od_step(c, X, st, to, z[0], z[1], o);
    double* p = Phi + (int64_t)k * 36 * nf + tid;
#pragma unroll
    for (int i = 0; i < 36; ++i) __stcs(p + (int64_t)i * nf, o[kOdPhi + i]);
    p = Ht + (int64_t)k * 12 * nf + tid;
#pragma unroll
    for (int i = 0; i < 12; ++i) __stcs(p + (int64_t)i * nf, o[kOdH + i]);
    p = real_obs + (int64_t)k * 2 * nf + tid;
    __stcs(p, o[kOdReal]);
    __stcs(p + nf, o[kOdReal + 1]);
    p = comp_obs + (int64_t)k * 2 * nf + tid;
    __stcs(p, o[kOdComp]);
    __stcs(p + nf, o[kOdComp + 1]);
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) orbit[(int64_t)i * nf + tid] = X[i];
}

// Fused run: N = 6, M = 2 (the statOD shape).  STRICT selects the reference-order filter step (filters_strict.cuh).
template <bool STRICT>
__global__ void __launch_bounds__(kThreads)
od_run_kernel(const __grid_constant__ NlModel<6, 2> md, const __grid_constant__ OdParams c,
              const __grid_constant__ NlIo io, double* __restrict__ orbit, const double* __restrict__ station,
              const double* __restrict__ tobs) {
  constexpr int N = 6, M = 2, SN = N * (N + 1) / 2, PN = STRICT ? N * N : SN;
  extern __shared__ __align__(16) double icdf_tab[];
  icdf_load(icdf_tab);
  __syncthreads();
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= io.nf) return;
  double X[6], x[N], P[PN];
#pragma unroll
  for (int i = 0; i < 6; ++i) X[i] = orbit[(int64_t)i * io.nf + tid];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = io.vec[(int64_t)i * io.nf + tid];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const double v = io.mat[(int64_t)((i <= j) ? (i * N + j) : (j * N + i)) * io.nf + tid];
      if constexpr (STRICT) P[i * N + j] = v;
      else if (i <= j) P[sym_idx<N>(i, j)] = v;
    }
  const uint64_t gf = (uint64_t)(c.filter_offset + tid);
  int status = 0;
  for (int k = 0; k < io.steps; ++k) {
    const unsigned fl = io.flags ? io.flags[k] : (unsigned)GKB_F_MEAS;
    const bool has_meas = (fl & GKB_F_MEAS) != 0, ekf = (fl & GKB_F_EKF) != 0;
    double st[6], to[2], z[2], o[kOdRows];
#pragma unroll
    for (int i = 0; i < 6; ++i) st[i] = __ldg(station + (int64_t)k * 6 + i);
    to[0] = __ldg(tobs + (int64_t)k * 2);
    to[1] = __ldg(tobs + (int64_t)k * 2 + 1);
    philox_normals<2>(icdf_tab, c.seed, gf, (uint32_t)k, z);
    od_step(c, X, st, to, z[0], z[1], o);
    double Phi[N * N], Ht[M * N], ro[M], co[M];
#pragma unroll
    for (int i = 0; i < N * N; ++i) Phi[i] = o[kOdPhi + i];
#pragma unroll
    for (int i = 0; i < M * N; ++i) Ht[i] = has_meas ? o[kOdH + i] : 0.0;
#pragma unroll
    for (int a = 0; a < M; ++a) { ro[a] = has_meas ? o[kOdReal + a] : 0.0; co[a] = has_meas ? o[kOdComp + a] : 0.0; }
    int err;
    if constexpr (STRICT) {
      double Ppred[N * N], K[N * M], innov[M], obsdev[M];
      err = strict::hybrid_step<N, M>(md, x, P, Phi, Ht, ro, co, nullptr, has_meas, ekf, false, Ppred, K, innov, obsdev);
    } else {
      NlOut<N, M> no;
      err = hybrid_step<N, M>(md, x, P, Phi, Ht, ro, co, nullptr, has_meas, ekf, false, no);
    }
    if (err != 0) {
      if (status == 0) status = err;
      continue;
    }
    if (io.every_step) {
      if (io.o_state != nullptr) {
        double* dst = io.o_state + (int64_t)k * N * io.nf + tid;
#pragma unroll
        for (int i = 0; i < N; ++i) __stcs(dst + (int64_t)i * io.nf, x[i]);
      }
      if (io.o_covar != nullptr) {
        double* dst = io.o_covar + (int64_t)k * N * N * io.nf + tid;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
          for (int j = 0; j < N; ++j) {
            if constexpr (STRICT) __stcs(dst + (int64_t)(i * N + j) * io.nf, P[i * N + j]);
            else __stcs(dst + (int64_t)(i * N + j) * io.nf, P[sym_idx<N>(i, j)]);
          }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) orbit[(int64_t)i * io.nf + tid] = X[i];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    io.vec[(int64_t)i * io.nf + tid] = x[i];
    if (!io.every_step && io.o_state != nullptr) io.o_state[(int64_t)i * io.nf + tid] = x[i];
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double v;
      if constexpr (STRICT) v = P[i * N + j];
      else v = P[sym_idx<N>(i, j)];
      io.mat[(int64_t)(i * N + j) * io.nf + tid] = v;
      if (!io.every_step && io.o_covar != nullptr) io.o_covar[(int64_t)(i * N + j) * io.nf + tid] = v;
    }
  if (io.status != nullptr && status != 0 && io.status[tid] == 0) io.status[tid] = status;
}

static constexpr size_t kOdSmem = sizeof(double) * kIcdfSegments * kIcdfCoefs;

int launch_od_synth(const OdParams& c, int64_t nf, int steps, double* orbit, const double* station, const double* tobs,
                    double* Phi, double* Ht, double* real_obs, double* comp_obs, cudaStream_t s) {
  const unsigned grid = (unsigned)((nf + kThreads - 1) / kThreads);
  od_synth_kernel<<<grid, kThreads, kOdSmem, s>>>(c, nf, steps, orbit, station, tobs, Phi, Ht, real_obs, comp_obs);
  return 0;
}

int launch_od_run(const HostModel& hm, const OdParams& c, const NlIo& io, double* orbit, const double* station,
                  const double* tobs, cudaStream_t s) {
  if (hm.kind != GKB_HYBRID || hm.n != 6 || hm.m != 2) return GKB_ERR_UNSUPPORTED;
  NlModel<6, 2> md;
  memset(&md, 0, sizeof md);
  for (int i = 0; i < hm.q * hm.q; ++i) md.Q[i] = hm.Q[i];
  for (int i = 0; i < 4; ++i) { md.R[i] = hm.R[i]; md.L[i] = hm.L[i]; }
  md.q = hm.q;
  const unsigned grid = (unsigned)((io.nf + kThreads - 1) / kThreads);
  if (io.strict) od_run_kernel<true><<<grid, kThreads, kOdSmem, s>>>(md, c, io, orbit, station, tobs);
  else od_run_kernel<false><<<grid, kThreads, kOdSmem, s>>>(md, c, io, orbit, station, tobs);
  return 0;
}

}  // namespace gkb
