// kernels_nl_tma.cu -- the production path of the NLDKF kernels (HybridKF, SRIF): per-filter Phi / Htilde /
// observation streams staged through shared memory by the TMA, outputs after the last epoch only.
// Split from kernels_nl.cu (which keeps the general kernels) so that the two halves compile in parallel.
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through the runtime, no libcuda link)

#include <cstdlib>
#include <cstring>

#include "engine_internal.h"
#include "filters_nl.cuh"

namespace gkb {

// ---- TMA-staged hybrid kernel --------------------------------------------------------------------------
// Production configuration of the hybrid filter (per-filter Phi / Htilde / observation streams, no
// SNC, outputs after the last epoch only).  A CTA owns 128 consecutive filters; the 52 (n=6, m=2)
// input rows of an epoch are 1 KB contiguous segments of the SoA streams, copied into shared memory
// by the TMA (cp.async.bulk, one 1 KB bulk copy per row, completion counted on an mbarrier) two
// epochs ahead of the arithmetic, so HBM latency overlaps the FP64 work instead of stalling the two
// resident warps per scheduler.  Each thread then reads its own column with conflict-free LDS.64.
namespace tma {

GKB_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
GKB_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
GKB_DEV void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
GKB_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
GKB_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
GKB_DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 2-D tiled TMA load (cp.async.bulk.tensor): box {32 filters, rows} of a [rows_total][nf] stream.
GKB_DEV void tensor_g2s_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace tma

template <int N, int M>
__global__ void __launch_bounds__(kThreads)
hybrid_run_tma_kernel(const __grid_constant__ NlModel<N, M> md, const __grid_constant__ NlIo io) {
  constexpr int SN = N * (N + 1) / 2;
  constexpr int ROWS_PHI = N * N, ROWS_H = M * N, ROWS = ROWS_PHI + ROWS_H + 2 * M;
  constexpr int STAGES = 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stage = reinterpret_cast<double*>(smem_raw);                       // [STAGES][ROWS][kThreads]
  uint64_t* full = reinterpret_cast<uint64_t*>(stage + (size_t)STAGES * ROWS * kThreads);  // [STAGES]
  const int64_t cta_base = (int64_t)blockIdx.x * kThreads;
  const int64_t tid = cta_base + threadIdx.x;
  const bool active = tid < io.nf;
  const uint32_t cnt = (uint32_t)min((int64_t)kThreads, io.nf - cta_base);  // filters of this CTA (even)
  const uint32_t row_bytes = cnt * 8u;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) tma::mbar_init(&full[s], 1);
    tma::fence_barrier_init();
  }
  __syncthreads();

  // warp 0 issues the bulk copies of one epoch: lane r copies rows r, r + 32, ...
  auto issue = [&](int k, int s) {
    const bool has_meas = io.flags ? ((io.flags[k] & GKB_F_MEAS) != 0) : true;
    const int rows = has_meas ? ROWS : ROWS_PHI;
    double* dst = stage + (size_t)s * ROWS * kThreads;
    if (threadIdx.x == 0) tma::mbar_expect_tx(&full[s], (uint32_t)rows * row_bytes);
    __syncwarp();
    for (int r = threadIdx.x; r < rows; r += 32) {
      const double* src;
      if (r < ROWS_PHI) src = io.Phi + ((int64_t)k * ROWS_PHI + r) * io.nf;
      else if (r < ROWS_PHI + ROWS_H) src = io.Htilde + ((int64_t)k * ROWS_H + (r - ROWS_PHI)) * io.nf;
      else if (r < ROWS_PHI + ROWS_H + M) src = io.real_obs + ((int64_t)k * M + (r - ROWS_PHI - ROWS_H)) * io.nf;
      else src = io.computed_obs + ((int64_t)k * M + (r - ROWS_PHI - ROWS_H - M)) * io.nf;
      tma::bulk_g2s(dst + (size_t)r * kThreads, src + cta_base, row_bytes, &full[s]);
    }
  };

  double x[N], P[SN];
  if (active) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = io.vec[(int64_t)i * io.nf + tid];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) P[sym_idx<N>(i, j)] = io.mat[(int64_t)(i * N + j) * io.nf + tid];
  }
  if (threadIdx.x < 32) {
    issue(0, 0);
    if (io.steps > 1) issue(1, 1);
  }
  int status = 0;
  for (int k = 0; k < io.steps; ++k) {
    const int s = k & 1;
    const unsigned fl = io.flags ? io.flags[k] : (unsigned)GKB_F_MEAS;
    const bool has_meas = (fl & GKB_F_MEAS) != 0, ekf = (fl & GKB_F_EKF) != 0;
    tma::mbar_wait(&full[s], (uint32_t)((k >> 1) & 1));
    const double* col = stage + (size_t)s * ROWS * kThreads + threadIdx.x;
    double Phi[N * N], Ht[M * N], ro[M], co[M];
#pragma unroll
    for (int i = 0; i < N * N; ++i) Phi[i] = col[(size_t)i * kThreads];
    if (has_meas) {
#pragma unroll
      for (int i = 0; i < M * N; ++i) Ht[i] = col[(size_t)(ROWS_PHI + i) * kThreads];
#pragma unroll
      for (int a = 0; a < M; ++a) {
        ro[a] = col[(size_t)(ROWS_PHI + ROWS_H + a) * kThreads];
        co[a] = col[(size_t)(ROWS_PHI + ROWS_H + M + a) * kThreads];
      }
    } else {
#pragma unroll
      for (int i = 0; i < M * N; ++i) Ht[i] = 0.0;
#pragma unroll
      for (int a = 0; a < M; ++a) { ro[a] = 0.0; co[a] = 0.0; }
    }
    __syncthreads();  // every thread holds its epoch-k inputs in registers: the stage can be refilled
    if (threadIdx.x < 32 && k + STAGES < io.steps) issue(k + STAGES, s);
    if (active) {
      NlOut<N, M> o;
      int err = hybrid_step<N, M>(md, x, P, Phi, Ht, ro, co, nullptr, has_meas, ekf, false, o);
      if (err != 0 && status == 0) status = err;
    }
  }
  if (active) {
    if (io.o_state != nullptr) {
#pragma unroll
      for (int i = 0; i < N; ++i) io.o_state[(int64_t)i * io.nf + tid] = x[i];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) io.vec[(int64_t)i * io.nf + tid] = x[i];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const double v = P[sym_idx<N>(i, j)];
        io.mat[(int64_t)(i * N + j) * io.nf + tid] = v;
        if (io.o_covar != nullptr) io.o_covar[(int64_t)(i * N + j) * io.nf + tid] = v;
      }
    if (io.status != nullptr && status != 0 && io.status[tid] == 0) io.status[tid] = status;
  }
}

// ---- warp-private TMA pipelines (tensor maps) ------------------------------------------------------------
// Same production configuration as above, but no CTA-wide synchronisation at all: every warp owns 32
// consecutive filters and a private ring of kWStages shared-memory stages with one mbarrier each.  An epoch
// of a warp is four tiled TMA loads (cp.async.bulk.tensor.2d): the {32 filters x N*N rows} box of the Phi
// stream, {32 x M*N} of Htilde and {32 x M} of each observation stream.  As soon as a lane has moved its
// column of a stage into registers the stage is re-armed for epoch k + kWStages, so every warp always has
// one to two epochs (13 KB each at n = 6, m = 2) in flight while it does the FP64 work of the current one.
// Out-of-range filters of a ragged last warp are zero-filled by the TMA and never written back.
// The SRIF's GENERAL epoch (srif_step: LU with interchanges, votes, a full R), kept OUT OF LINE: it runs for Predict()
// epochs, for the first epoch of a run (R is a full matrix until the first Householder update) and when the
// speculative epoch gives up -- rare, and inlined it costs the hot loop its registers (the straight-line epoch
// srif_step_tri needs all 255).  State and inputs travel through local / shared memory on this path only.
template <int N, int M>
__device__ __noinline__ int srif_epoch_general(const NlModel<N, M>& md, double* __restrict__ b_io, double* __restrict__ R_io,
                                               const double* __restrict__ col, bool has_meas, bool pad_lane) {
  constexpr int ROWS_PHI = N * N, ROWS_H = M * N;
  double b[N], R[N * N], Phi[N * N], Ht[M * N], ro[M], co[M];
#pragma unroll
  for (int i = 0; i < N; ++i) b[i] = b_io[i];
#pragma unroll
  for (int i = 0; i < N * N; ++i) { R[i] = R_io[i]; Phi[i] = col[i * 32]; }
#pragma unroll
  for (int i = 0; i < M * N; ++i) Ht[i] = has_meas ? col[(ROWS_PHI + i) * 32] : 0.0;
#pragma unroll
  for (int a = 0; a < M; ++a) {
    ro[a] = has_meas ? col[(ROWS_PHI + ROWS_H + a) * 32] : 0.0;
    co[a] = has_meas ? col[(ROWS_PHI + ROWS_H + M + a) * 32] : 0.0;
  }
  if (pad_lane) {  // keep the padding lanes' Phi invertible (zero-filled by the TMA)
#pragma unroll
    for (int i = 0; i < N; ++i) Phi[i * N + i] = 1.0;
  }
  NlOut<N, M> o;
  const int err = srif_step<N, M>(md, b, R, Phi, Ht, ro, co, has_meas, o);
#pragma unroll
  for (int i = 0; i < N; ++i) b_io[i] = b[i];
#pragma unroll
  for (int i = 0; i < N * N; ++i) R_io[i] = R[i];
  return err;
}

struct NlTensorMaps {
  alignas(64) CUtensorMap phi;
  alignas(64) CUtensorMap h;
  alignas(64) CUtensorMap real_obs;
  alignas(64) CUtensorMap computed_obs;
};
constexpr int kWStages = 2;

// Work distribution: a TASK is (chunk c of the epochs, group g of 32 consecutive filters).  Persistent warps claim
// tasks from one atomic counter in chunk-major order (all groups' chunk 0, then chunk 1, ...), so that the epochs of
// the last, partially filled wave of groups are spread over every resident warp instead of leaving a tail: with
// 10^5 filters = 3125 groups over 1184 resident warps a static one-group-per-warp grid takes 3 rounds for 2.64
// rounds of work; split in 3 chunks it takes 7.92 -> 8 thirds.  A group's state travels between its chunks through
// the handle's state arrays (L2) with a release / acquire flag per group; chunk c of a group is claimed a full pass
// over all groups after chunk c-1, so the acquire practically never spins.  chunks = 1 is the plain one-pass run.
// Spin (lane 0) until the group's flag reaches `value`, then make the state written by the previous owner of the
// group visible to the whole warp.
__device__ __forceinline__ void acquire_group_state(const int* flag, int value, int lane) {
  if (lane == 0) {
    int seen;
    do {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
      if (seen < value) __nanosleep(200);
    } while (seen < value);
  }
  __syncwarp();
  __threadfence();
}

// Epochs [k0, k1) of the 32 filters of group g, by one warp: the input boxes stream through the warp's private ring
// of kWStages stages (s / phase: the ring position, carried from task to task), the state comes from and goes back
// to the handle's state arrays; `last` = these are the final epochs of the call (the Estimate read-outs are written).
template <int N, int M, bool SRIF, bool SCHED>
__device__ __forceinline__ void nl_wtma_task(const NlModel<N, M>& md, const NlIo& io, const NlTensorMaps& maps, double* ring,
                                             uint64_t* full, int& s, uint32_t& phase, int lane, int g, int k0, int k1,
                                             bool last, const int* wait_for, int wait_value) {
  constexpr int SN = SRIF ? N * N : N * (N + 1) / 2;  // SRIF keeps the full sqrt-information matrix R
  constexpr int ROWS_PHI = N * N, ROWS_H = M * N, ROWS = ROWS_PHI + ROWS_H + 2 * M;
  constexpr uint32_t kBytesPhi = ROWS_PHI * 32 * 8, kBytesAll = ROWS * 32 * 8;
  const int64_t warp_base = (int64_t)g * 32;
  const int64_t tid = warp_base + lane;
  const bool active = tid < io.nf;

  auto issue = [&](int k, int st) {  // lane 0 only
    const bool has_meas = io.flags ? ((io.flags[k] & GKB_F_MEAS) != 0) : true;
    double* dst = ring + (size_t)st * ROWS * 32;
    tma::mbar_expect_tx(&full[st], has_meas ? kBytesAll : kBytesPhi);
    tma::tensor_g2s_2d(dst, &maps.phi, (int)warp_base, k * ROWS_PHI, &full[st]);
    if (has_meas) {
      tma::tensor_g2s_2d(dst + ROWS_PHI * 32, &maps.h, (int)warp_base, k * ROWS_H, &full[st]);
      tma::tensor_g2s_2d(dst + (ROWS_PHI + ROWS_H) * 32, &maps.real_obs, (int)warp_base, k * M, &full[st]);
      tma::tensor_g2s_2d(dst + (ROWS_PHI + ROWS_H + M) * 32, &maps.computed_obs, (int)warp_base, k * M, &full[st]);
    }
  };
  // the input streams do not depend on the previous chunk: start them before waiting for the state
  if (lane == 0) {
    if constexpr (SCHED) {
      int st = s;
#pragma unroll
      for (int j = 0; j < kWStages; ++j) {
        if (k0 + j < k1) issue(k0 + j, st);
        if (++st == kWStages) st = 0;
      }
    } else {  // one task per warp: the ring starts at stage 0
#pragma unroll
      for (int j = 0; j < kWStages; ++j)
        if (j < k1) issue(j, j);
    }
  }
  if (SCHED && wait_for != nullptr) acquire_group_state(wait_for, wait_value, lane);
  // SCHED: the state may have been written by another SM -> read it from L2 (__ldcg), never from a stale L1 line
  auto ld_state = [](const double* p) { return SCHED ? __ldcg(p) : *p; };

  double x[N], P[SN];  // hybrid: x, P (packed upper);  SRIF: b, R
  if (active) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = ld_state(io.vec + (int64_t)i * io.nf + tid);
    if constexpr (SRIF) {
#pragma unroll
      for (int i = 0; i < N * N; ++i) P[i] = ld_state(io.mat + (int64_t)i * io.nf + tid);
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i; j < N; ++j) P[sym_idx<N>(i, j)] = ld_state(io.mat + (int64_t)(i * N + j) * io.nf + tid);
    }
  } else {  // lanes past the last filter run on an identity problem (the TMA zero-fills their columns)
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = 0.0;
#pragma unroll
    for (int i = 0; i < SN; ++i) P[i] = 0.0;
    if constexpr (SRIF) {
#pragma unroll
      for (int i = 0; i < N; ++i) P[i * N + i] = 1.0;
    }
  }
  int status = 0;
  bool try_fast = SRIF && N >= 3;  // the speculative SRIF epoch (srif_step_tri); given up for the task once it fails
  for (int k = k0; k < k1; ++k) {
    if constexpr (SRIF && N >= 3) {
      // Runs of measurement epochs with R upper triangular in every lane (all of them, once the first Householder
      // update has happened) go through the straight-line epoch with R packed; anything else -- a Predict() epoch,
      // a full R, a lane that needs a pivot interchange or hits an error -- falls through to the general epoch
      // below with the stage untouched.
      bool lower_zero = true;
#pragma unroll
      for (int i = 1; i < N; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) lower_zero = lower_zero && (P[i * N + j] == 0.0);
      if (try_fast && __all_sync(0xffffffffu, lower_zero)) {
        double U[N * (N + 1) / 2];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
          for (int j = i; j < N; ++j) U[sym_idx<N>(i, j)] = P[i * N + j];
        while (k < k1) {
          if (io.flags && (io.flags[k] & GKB_F_MEAS) == 0) break;
          tma::mbar_wait(&full[s], phase);
          if (!srif_step_tri<N, M, 32>(md, x, U, ring + (size_t)s * ROWS * 32 + lane, !active)) {
            try_fast = false;
            break;
          }
          __syncwarp();  // every lane is done with the stage: re-arm it
          if (lane == 0 && k + kWStages < k1) issue(k + kWStages, s);
          if (++s == kWStages) { s = 0; phase ^= 1u; }
          ++k;
        }
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
          for (int j = i; j < N; ++j) P[i * N + j] = U[sym_idx<N>(i, j)];
        if (k >= k1) break;
      }
    }
    const unsigned fl = io.flags ? io.flags[k] : (unsigned)GKB_F_MEAS;
    const bool has_meas = (fl & GKB_F_MEAS) != 0, ekf = (fl & GKB_F_EKF) != 0;
    tma::mbar_wait(&full[s], phase);
    const double* col = ring + (size_t)s * ROWS * 32 + lane;
    if constexpr (SRIF) {
      // the general epoch, out of line (reads its inputs from the stage, state through a local copy)
      double tb[N], tR[N * N];
#pragma unroll
      for (int i = 0; i < N; ++i) tb[i] = x[i];
#pragma unroll
      for (int i = 0; i < N * N; ++i) tR[i] = P[i];
      const int err = srif_epoch_general<N, M>(md, tb, tR, col, has_meas, !active);
#pragma unroll
      for (int i = 0; i < N; ++i) x[i] = tb[i];
#pragma unroll
      for (int i = 0; i < N * N; ++i) P[i] = tR[i];
      __syncwarp();  // every lane is done with the stage: re-arm it
      if (lane == 0 && k + kWStages < k1) issue(k + kWStages, s);
      if (++s == kWStages) { s = 0; phase ^= 1u; }
      if (err != 0 && status == 0) status = err;
    } else {
    double Phi[N * N], Ht[M * N], ro[M], co[M];
#pragma unroll
    for (int i = 0; i < N * N; ++i) Phi[i] = col[i * 32];
    if (has_meas) {
#pragma unroll
      for (int i = 0; i < M * N; ++i) Ht[i] = col[(ROWS_PHI + i) * 32];
#pragma unroll
      for (int a = 0; a < M; ++a) {
        ro[a] = col[(ROWS_PHI + ROWS_H + a) * 32];
        co[a] = col[(ROWS_PHI + ROWS_H + M + a) * 32];
      }
    } else {
#pragma unroll
      for (int i = 0; i < M * N; ++i) Ht[i] = 0.0;
#pragma unroll
      for (int a = 0; a < M; ++a) { ro[a] = 0.0; co[a] = 0.0; }
    }
    __syncwarp();  // all 32 columns of the stage are in registers: re-arm it
    if (lane == 0 && k + kWStages < k1) issue(k + kWStages, s);
    if (++s == kWStages) { s = 0; phase ^= 1u; }
    NlOut<N, M> o;
    int err;
    {
      err = hybrid_step<N, M>(md, x, P, Phi, Ht, ro, co, nullptr, has_meas, ekf, false, o);
      if (io.every_step && active && err == 0) {  // Estimate k: streamed out once, never read back here
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
            for (int j = 0; j < N; ++j) __stcs(dst + (int64_t)(i * N + j) * io.nf, P[sym_idx<N>(i, j)]);
        }
      }
    }
    if (err != 0 && status == 0) status = err;
    }  // !SRIF
  }
  if constexpr (SRIF) {
    // read-outs of the last estimate: State() = inv(R) b (srif.go:223-235), Covariance() = inv(R) inv(R)^T (253-265)
    if (last && io.o_state != nullptr) {
      double xs[N];
      if (!srif_state<N>(xs, P, x)) {
        if (status == 0) status = GKB_ERR_SINGULAR_R;
#pragma unroll
        for (int i = 0; i < N; ++i) xs[i] = 0.0;
      }
      if (active) {
#pragma unroll
        for (int i = 0; i < N; ++i) io.o_state[(int64_t)i * io.nf + tid] = xs[i];
      }
    }
    if (last && io.o_covar != nullptr) {
      double Pc[N * N];
      srif_covariance<N>(Pc, P);
      if (active) {
#pragma unroll
        for (int i = 0; i < N * N; ++i) io.o_covar[(int64_t)i * io.nf + tid] = Pc[i];
      }
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < N; ++i) io.vec[(int64_t)i * io.nf + tid] = x[i];
#pragma unroll
      for (int i = 0; i < N * N; ++i) io.mat[(int64_t)i * io.nf + tid] = P[i];
    }
  } else if (active) {
    if (last && !io.every_step && io.o_state != nullptr) {
#pragma unroll
      for (int i = 0; i < N; ++i) io.o_state[(int64_t)i * io.nf + tid] = x[i];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) io.vec[(int64_t)i * io.nf + tid] = x[i];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const double v = P[sym_idx<N>(i, j)];
        io.mat[(int64_t)(i * N + j) * io.nf + tid] = v;
        if (last && !io.every_step && io.o_covar != nullptr) io.o_covar[(int64_t)(i * N + j) * io.nf + tid] = v;
      }
  }
  if (active && io.status != nullptr && status != 0 && io.status[tid] == 0) io.status[tid] = status;
}

// One task per warp, taken from the block index: grid = all groups, no scheduler.  Used when the groups fit in one
// round of resident warps.
template <int N, int M, bool SRIF>
__global__ void __launch_bounds__(kThreads)
nl_run_wtma_kernel(const __grid_constant__ NlModel<N, M> md, const __grid_constant__ NlIo io,
                   const __grid_constant__ NlTensorMaps maps) {
  constexpr int ROWS = N * N + M * N + 2 * M;
  constexpr int kWarpsPerCta = kThreads / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* ring = reinterpret_cast<double*>(smem_raw) + (size_t)warp * kWStages * ROWS * 32;  // [stage][row][lane]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + sizeof(double) * kWarpsPerCta * kWStages * ROWS * 32) +
                   warp * kWStages;
  const int g = blockIdx.x * kWarpsPerCta + warp;
  if ((int64_t)g * 32 >= io.nf) return;  // whole warp out of range (no CTA-wide barrier anywhere below)
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kWStages; ++st) tma::mbar_init(&full[st], 1);
    tma::fence_barrier_init();
  }
  __syncwarp();
  int s = 0;
  uint32_t phase = 0;
  nl_wtma_task<N, M, SRIF, false>(md, io, maps, ring, full, s, phase, lane, g, 0, io.steps, true, nullptr, 0);
}

// Work distribution for more than one round of groups: a TASK is (chunk c of the epochs, group g of 32 consecutive
// filters).  Persistent warps claim tasks from one atomic counter in chunk-major order (all groups' chunk 0, then
// chunk 1, ...), so that the epochs of the last, partially filled wave of groups are spread over every resident warp
// instead of leaving a tail: 10^5 filters = 3125 groups over 1184 resident warps take 3 rounds for 2.64 rounds of
// work when every warp owns whole groups; split in 3 chunks it is 7.92 -> 8 thirds.  A group's state travels between
// its chunks through the handle's state arrays (L2) with a release / acquire flag per group; chunk c of a group is
// claimed a full pass over all groups after chunk c - 1, so the acquire practically never spins.
template <int N, int M, bool SRIF>
__global__ void __launch_bounds__(kThreads)
nl_run_wtma_sched_kernel(const __grid_constant__ NlModel<N, M> md, const __grid_constant__ NlIo io,
                         const __grid_constant__ NlTensorMaps maps) {
  constexpr int ROWS = N * N + M * N + 2 * M;
  constexpr int kWarpsPerCta = kThreads / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* ring = reinterpret_cast<double*>(smem_raw) + (size_t)warp * kWStages * ROWS * 32;  // [stage][row][lane]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + sizeof(double) * kWarpsPerCta * kWStages * ROWS * 32) +
                   warp * kWStages;
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kWStages; ++st) tma::mbar_init(&full[st], 1);
    tma::fence_barrier_init();
  }
  __syncwarp();
  const int groups = (int)((io.nf + 31) / 32);
  const int n_tasks = groups * io.chunks;
  int* next_task = io.sched;  // [0]: task counter;  [1 + g]: chunks of group g already written back
  int* done = io.sched + 1;
  int s = 0;
  uint32_t phase = 0;
  for (;;) {
    int task = 0;
    if (lane == 0) task = atomicAdd(next_task, 1);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= n_tasks) break;
    const int c = task / groups, g = task - c * groups;
    const int k0 = c * io.chunk_len, k1 = min(io.steps, k0 + io.chunk_len);
    const bool last = (c == io.chunks - 1);
    nl_wtma_task<N, M, SRIF, true>(md, io, maps, ring, full, s, phase, lane, g, k0, k1, last, c > 0 ? done + g : nullptr, c);
    if (!last) {  // publish the state to whichever warp claims the next chunk of this group
      __threadfence();
      __syncwarp();
      if (lane == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(done + g), "r"(c + 1) : "memory");
    }
  }
}

// cuTensorMapEncodeTiled, fetched through the runtime so the library carries no link-time libcuda dependency.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// [rows_total][nf] FP64 stream, box = {32 filters, box_rows}; false when TMA cannot describe it.
static bool make_stream_map(CUtensorMap* map, const double* base, int64_t nf, int64_t rows_total, int box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || rows_total < 1 || rows_total > 0x7fffffffLL || box_rows > 256) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)nf, (cuuint64_t)rows_total};
  const cuuint64_t strides[1] = {(cuuint64_t)nf * sizeof(double)};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int N, int M>
static int launch_nl_tma_shape(const HostModel& hm, const NlIo& io, cudaStream_t s) {
  const unsigned grid = (unsigned)((io.nf + kThreads - 1) / kThreads);
  NlModel<N, M> md;
  for (int i = 0; i < GKB_MAX_Q * GKB_MAX_Q; ++i) md.Q[i] = 0.0;
  for (int i = 0; i < hm.q * hm.q; ++i) md.Q[i] = hm.Q[i];
  for (int i = 0; i < M * M; ++i) { md.R[i] = hm.R[i]; md.L[i] = hm.L[i]; }
  md.q = hm.q;
  // TMA-staged fast path (the production configuration of both NLDKF kinds): per-filter streams, no SNC
  // epochs, final-estimate outputs only, streams that satisfy the TMA's 16-byte rules (even filter count,
  // 16-byte aligned bases).
  constexpr int ROWS = N * N + M * N + 2 * M;
  const size_t smem = sizeof(double) * 2 * ROWS * kThreads + 2 * sizeof(uint64_t);
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  // every-step outputs (state / covariance of each epoch, the input of SmoothAll) are streamed by the hybrid kernel;
  // the SRIF's per-epoch read-outs need inv(R) each time and stay on the general kernel
  const bool fast = !io.phi_shared && !io.h_shared && io.Gamma == nullptr && (!io.every_step || hm.kind == GKB_HYBRID) &&
                    io.Htilde != nullptr &&
                    io.real_obs != nullptr && io.computed_obs != nullptr && (io.nf % 2 == 0) && aligned(io.Phi) &&
                    aligned(io.Htilde) && aligned(io.real_obs) && aligned(io.computed_obs) && io.o_meas == nullptr &&
                    io.o_innov == nullptr && io.o_pred == nullptr && io.o_gain == nullptr && io.o_obsdev == nullptr &&
                    smem <= 110 * 1024 && io.steps >= 2 && io.nf < 0x7fffffffLL;
  const char* path = getenv("GKB_NL_PATH");  // A/B switch for tests and profiling: plain | bulk | tensor
  const bool want_bulk = path && !strcmp(path, "bulk"), want_plain = path && !strcmp(path, "plain");
  const bool srif = hm.kind == GKB_SRIF;
  if (hm.kind != GKB_HYBRID && !srif) return 1;
  if (fast && !want_plain && !(want_bulk && !srif)) {
    NlTensorMaps maps;
    const int64_t st = io.steps;
    if (make_stream_map(&maps.phi, io.Phi, io.nf, st * N * N, N * N) &&
        make_stream_map(&maps.h, io.Htilde, io.nf, st * M * N, M * N) &&
        make_stream_map(&maps.real_obs, io.real_obs, io.nf, st * M, M) &&
        make_stream_map(&maps.computed_obs, io.computed_obs, io.nf, st * M, M)) {
      const size_t wsmem = sizeof(double) * (kThreads / 32) * kWStages * ROWS * 32 + (kThreads / 32) * kWStages * sizeof(uint64_t);
      // persistent grid: as many CTAs as stay resident; epochs split into the smallest number of chunks (<= 8) that
      // makes (groups x chunks) fill whole rounds of the resident warps to >= 97 %
      int sms = 148, device = 0;
      cudaGetDevice(&device);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
      auto launch_with = [&](bool persist, auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, wsmem);
        if (per_sm < 1) per_sm = 1;
        const int64_t groups = (io.nf + 31) / 32;
        const int64_t slots = (int64_t)sms * per_sm * (kThreads / 32);
        // More groups than resident warps: dynamic claiming needs ~8 tasks per warp to even out warps that run at
        // different speeds (coarser tasks let a fast warp start one more whole task at the very end); among the
        // neighbouring chunk counts take the one whose task count fills whole rounds best.  Chunks stay >= 16 epochs.
        int chunks = 1;
        if (persist && groups > slots && io.sched != nullptr) {
          const double per_slot = (double)groups / (double)slots;
          const int base = (int)(8.0 / per_slot + 0.5);
          double best = -1.0;
          for (int c = base - 1; c <= base + 1; ++c) {
            if (c < 1 || (c > 1 && io.steps / c < 16)) continue;
            const double rounds = per_slot * c;
            const double eff = rounds / (double)(int64_t)(rounds + 0.999999);
            if (eff > best + 1e-9) { best = eff; chunks = c; }
          }
        }
        if (const char* e = getenv("GKB_NL_CHUNKS")) { const int c = atoi(e); if (persist && c >= 1 && c <= io.steps) chunks = c; }
        NlIo io2 = io;
        io2.chunks = chunks;
        io2.chunk_len = (io.steps + chunks - 1) / chunks;
        io2.chunks = (io.steps + io2.chunk_len - 1) / io2.chunk_len;
        int64_t ctas = (groups + (kThreads / 32) - 1) / (kThreads / 32);
        if (persist) {
          if (ctas > (int64_t)sms * per_sm) ctas = (int64_t)sms * per_sm;
          cudaMemsetAsync(io.sched, 0, sizeof(int) * (size_t)(groups + 1), s);
        }
        kern<<<(unsigned)ctas, kThreads, wsmem, s>>>(md, io2, maps);
      };
      // the scheduler only pays when there is more than one round of groups (or when a test forces it)
      const bool forced = getenv("GKB_NL_CHUNKS") != nullptr;
      const bool multi_round = (io.nf + 31) / 32 > (int64_t)sms * 2 * (kThreads / 32);
      if (srif) {
        if (forced || multi_round) launch_with(true, nl_run_wtma_sched_kernel<N, M, true>);
        else launch_with(false, nl_run_wtma_kernel<N, M, true>);
      } else {
        if (forced || multi_round) launch_with(true, nl_run_wtma_sched_kernel<N, M, false>);
        else launch_with(false, nl_run_wtma_kernel<N, M, false>);
      }
      return 0;
    }
  }
  if (fast && want_bulk && !srif) {
    auto kern = hybrid_run_tma_kernel<N, M>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    kern<<<grid, kThreads, smem, s>>>(md, io);
    return 0;
  }
  return 1;  // not this path: the caller falls back to the general kernels
}

// 0 = launched; 1 = the call is not in the production configuration (or GKB_NL_PATH=plain): use the general kernels.
int launch_nl_tma(const HostModel& hm, const NlIo& io, cudaStream_t s) {
#define GKB_CASE(NN, MM) \
  if (hm.n == NN && hm.m == MM) return launch_nl_tma_shape<NN, MM>(hm, io, s);
  GKB_FOR_EACH_SHAPE(GKB_CASE)
#undef GKB_CASE
  return 1;
}

}  // namespace gkb
