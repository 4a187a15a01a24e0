// kernels_od.cu -- orbit-determination inputs on the device (od_synth.cuh) and the FUSED OD filter run.
//
//   od_synth_kernel   writes the per-filter Phi / Htilde / real / computed streams of gkb_nl_run ([step][component]
//                     [filter], coalesced streaming stores) from the filters' initial reference orbits: the stream
//                     generator of the bench / tests, and the "precompute once, filter many times" path.
//   od_run_kernel     the same synthesis FUSED with the hybrid CKF / EKF step (hybrid.go:104-204): every epoch's Phi,
//                     Htilde and observations are produced in the thread that consumes them and never touch HBM.
//                     Host traffic of a whole run: 48 B per filter in (initial orbit), the final estimate out.
// od_step is inlined into both with explicit fma()s, so the fused run is bit-identical to gkb_nl_run on the synthesised
// streams.  The fused run is scheduled in (chunk, group) tasks when there are more groups than resident warps (sched.cuh).
#include <cstring>

#include "engine_internal.h"
#include "filters_nl.cuh"
#include "filters_strict.cuh"
#include "od_synth.cuh"
#include "sched.cuh"

namespace gkb {

using strict::kStrictThreads;  // (GKB_SM's stride)

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
// Epochs [k0, k1) of filter `tid`: reference orbit, filter state and covariance come from and go back to the handle's
// arrays; `last` = these are the final epochs of the call (the final-estimate outputs are written).
template <bool STRICT, bool SCHED>
__device__ __forceinline__ void od_task(const NlModel<6, 2>& md, const OdParams& c, const NlIo& io, double* __restrict__ orbit,
                                        const double* __restrict__ station, const double* __restrict__ tobs,
                                        const double* icdf_tab, double* sm, int64_t tid, int k0, int k1, bool last) {
  // STRICT: the covariance (packed), the work matrix and the K R slot live in lane-private shared-memory columns at `sm`
  // (filters_strict.cuh: hybrid_sm_predict / hybrid_sm_update), exactly as in hybrid_run_strict_kernel
  constexpr int N = 6, M = 2, SN = N * (N + 1) / 2, PN = STRICT ? 1 : SN;
  static_assert(strict::kStrictThreads == kThreads, "the strict step's shared-memory stride is the CTA size");
  double* Ps = sm;
  double* Ws = sm + SN * kThreads;
  double* Hs = Ws + N * N * kThreads;
  auto ld_state = [](const double* p) { return SCHED ? __ldcg(p) : *p; };  // SCHED: written by another SM -> read from L2
  double X[6], x[N], P[PN];
#pragma unroll
  for (int i = 0; i < 6; ++i) X[i] = ld_state(orbit + (int64_t)i * io.nf + tid);
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = ld_state(io.vec + (int64_t)i * io.nf + tid);
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (i > j) continue;  // the stored matrix is the mirrored upper triangle (AsSymDense)
      const double v = ld_state(io.mat + (int64_t)(i * N + j) * io.nf + tid);
      if constexpr (STRICT) GKB_SM(Ps, sym_idx<N>(i, j)) = v;
      else P[sym_idx<N>(i, j)] = v;
    }
  const uint64_t gf = (uint64_t)(c.filter_offset + tid);
  int status = 0;
  for (int k = k0; k < k1; ++k) {
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
      double xbar[N], K[N * M], innov[M], obsdev[M];
      strict::hybrid_sm_predict<N, M>(md, x, Ps, Ws, Phi, nullptr, false, ekf, xbar);
      err = strict::hybrid_sm_update<N, M>(md, x, Ws, Hs, xbar, Ht, ro, co, has_meas, ekf, nullptr, io.nf, K, innov, obsdev);
      if (err == 0) strict::hybrid_sm_commit<N>(Ps, Ws);
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
            if constexpr (STRICT) __stcs(dst + (int64_t)(i * N + j) * io.nf, GKB_SM(Ps, sym_idx<N>(i, j)));
            else __stcs(dst + (int64_t)(i * N + j) * io.nf, P[sym_idx<N>(i, j)]);
          }
      }
    }
  }
  const bool final_out = last && !io.every_step;
#pragma unroll
  for (int i = 0; i < 6; ++i) orbit[(int64_t)i * io.nf + tid] = X[i];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    io.vec[(int64_t)i * io.nf + tid] = x[i];
    if (final_out && io.o_state != nullptr) io.o_state[(int64_t)i * io.nf + tid] = x[i];
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double v;
      if constexpr (STRICT) v = GKB_SM(Ps, sym_idx<N>(i, j));
      else v = P[sym_idx<N>(i, j)];
      io.mat[(int64_t)(i * N + j) * io.nf + tid] = v;
      if (final_out && io.o_covar != nullptr) io.o_covar[(int64_t)(i * N + j) * io.nf + tid] = v;
    }
  if (io.status != nullptr && status != 0 && io.status[tid] == 0) io.status[tid] = status;
}

// io.chunks == 0: one whole-run task per warp (grid = all groups); otherwise persistent warps under the (chunk, group)
// scheduler of sched.cuh -- same per-filter arithmetic, hence the same bits.
template <bool STRICT>
__global__ void __launch_bounds__(kThreads)
od_run_kernel(const __grid_constant__ NlModel<6, 2> md, const __grid_constant__ OdParams c,
              const __grid_constant__ NlIo io, double* __restrict__ orbit, const double* __restrict__ station,
              const double* __restrict__ tobs) {
  extern __shared__ __align__(16) double icdf_tab[];
  icdf_load(icdf_tab);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double* sm = STRICT ? icdf_tab + kIcdfSegments * kIcdfCoefs + threadIdx.x : nullptr;  // the strict step's matrix columns
  if (io.chunks <= 0) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < io.nf) od_task<STRICT, false>(md, c, io, orbit, station, tobs, icdf_tab, sm, tid, 0, io.steps, true);
    return;
  }
  const int groups = (int)((io.nf + 31) / 32);
  const int n_tasks = groups * io.chunks;
  int ch, g;
  while (sched_claim(io.sched, n_tasks, groups, lane, ch, g)) {
    const int k0 = ch * io.chunk_len, k1 = min(io.steps, k0 + io.chunk_len);
    if (ch > 0) sched_acquire_group(io.sched + 1 + g, ch, lane);
    const int64_t tid = (int64_t)g * 32 + lane;
    if (tid < io.nf) od_task<STRICT, true>(md, c, io, orbit, station, tobs, icdf_tab, sm, tid, k0, k1, ch == io.chunks - 1);
    if (ch != io.chunks - 1) sched_release_group(io.sched + 1 + g, ch + 1, lane);
  }
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
  auto launch = [&](auto kern, size_t smem) {
    int sms = 148, device = 0, per_sm = 1;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (smem > 48 * 1024) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
    if (per_sm < 1) per_sm = 1;
    constexpr int kWarps = kThreads / 32;
    const int64_t groups = (io.nf + 31) / 32;
    int64_t ctas = (groups + kWarps - 1) / kWarps;
    NlIo io2 = io;
    bool forced = false;
    const int chunks = io.sched != nullptr ? sched_pick_chunks(groups, (int64_t)sms * per_sm * kWarps, io.steps, 16.0, &forced) : 1;
    const bool scheduled = io.sched != nullptr && (chunks > 1 || forced);
    sched_set_chunks(io2, chunks, scheduled);
    if (scheduled) {
      if (ctas > (int64_t)sms * per_sm) ctas = (int64_t)sms * per_sm;
      cudaMemsetAsync(io.sched, 0, sizeof(int) * (size_t)(groups + 1), s);
    }
    kern<<<(unsigned)ctas, kThreads, smem, s>>>(md, c, io2, orbit, station, tobs);
  };
  // strict: + packed P (21), work matrix (36) and the K R slot (12) per thread = 70.7 KB per CTA; two CTAs per SM still fit
  if (io.strict) launch(od_run_kernel<true>, kOdSmem + sizeof(double) * (21 + 36 + 12) * kThreads);
  else launch(od_run_kernel<false>, kOdSmem);
  return 0;
}

}  // namespace gkb
