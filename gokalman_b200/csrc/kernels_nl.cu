// kernels_nl.cu -- batched NLDKF kernels: HybridKF (CKF / EKF, hybrid.go:104-204) and SRIF
// (srif.go:101-160), one filter per thread, time loop in the kernel.
//
// Per epoch each filter consumes its own Phi (n x n), Htilde (m x n), real and computed
// observation, streamed from HBM in SoA [step][component][filter]: a warp reads one coalesced
// 256-byte row per component.  Per-epoch flags (Predict vs Update, EKF on/off, SNC on/off) are
// shared by all filters, so the control flow is warp-uniform.
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through the runtime, no libcuda link)

#include <cstdlib>
#include <cstring>

#include "engine_internal.h"
#include "filters_nl.cuh"

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

// ---- SmoothAll (hybrid.go:209-238, srif.go:165-192) ---------------------------------------------------------
// Backward sweep over a stored history: for k = steps-2 .. 0:  S = inv(Phi_{k+1}),  x_k = S x_{k+1},
// P_k = (S P_{k+1}) S^T (upper triangle kept, AsSymDense).  x_{k+1} / P_{k+1} are the values the previous
// iteration wrote, so every estimate is the final one mapped back through the STMs -- exactly what the
// reference's in-place loop does.  One filter per thread; per (filter, step) the kernel reads one Phi and
// writes one state + covariance: HBM-bound streaming, no reuse.
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
  for (int k = steps - 2; k >= 0; --k) {
    double S[N * N];
    nl_load<N * N>(S, Phi, phi_shared, k + 1, nf, tid);
    if (inverse_lu<N>(S) != 0) {  // "provided STM is not invertible": the reference stops here
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
    GKB_SM(1) GKB_SM(2) GKB_SM(3) GKB_SM(4) GKB_SM(5) GKB_SM(6)
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
  GKB_FOR_EACH_SHAPE(GKB_CASE)
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
__global__ void __launch_bounds__(kThreads)
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
  for (int k = 0; k < steps; ++k) {
    double Hk[M * N], ro[M], co[M];
    nl_load<M * N>(Hk, H, h_shared, k, nf, tid);
    nl_load<M>(ro, real_obs, 0, k, nf, tid);
    nl_load<M>(co, computed_obs, 0, k, nf, tid);
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
  GKB_FOR_EACH_SHAPE(GKB_CASE)
#undef GKB_CASE
  return GKB_ERR_UNSUPPORTED;
}

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
struct NlTensorMaps {
  alignas(64) CUtensorMap phi;
  alignas(64) CUtensorMap h;
  alignas(64) CUtensorMap real_obs;
  alignas(64) CUtensorMap computed_obs;
};
constexpr int kWStages = 2;

template <int N, int M, bool SRIF>
__global__ void __launch_bounds__(kThreads)
nl_run_wtma_kernel(const __grid_constant__ NlModel<N, M> md, const __grid_constant__ NlIo io,
                   const __grid_constant__ NlTensorMaps maps) {
  constexpr int SN = SRIF ? N * N : N * (N + 1) / 2;  // SRIF keeps the full sqrt-information matrix R
  constexpr int ROWS_PHI = N * N, ROWS_H = M * N, ROWS = ROWS_PHI + ROWS_H + 2 * M;
  constexpr int kWarpsPerCta = kThreads / 32;
  constexpr uint32_t kBytesPhi = ROWS_PHI * 32 * 8, kBytesAll = ROWS * 32 * 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* ring = reinterpret_cast<double*>(smem_raw) + (size_t)warp * kWStages * ROWS * 32;  // [stage][row][lane]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + sizeof(double) * kWarpsPerCta * kWStages * ROWS * 32) +
                   warp * kWStages;
  const int64_t warp_base = ((int64_t)blockIdx.x * kWarpsPerCta + warp) * 32;
  if (warp_base >= io.nf) return;  // whole warp out of range (no CTA-wide barrier anywhere below)
  const int64_t tid = warp_base + lane;
  const bool active = tid < io.nf;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kWStages; ++s) tma::mbar_init(&full[s], 1);
    tma::fence_barrier_init();
  }
  __syncwarp();

  auto issue = [&](int k, int s) {  // lane 0 only
    const bool has_meas = io.flags ? ((io.flags[k] & GKB_F_MEAS) != 0) : true;
    double* dst = ring + (size_t)s * ROWS * 32;
    tma::mbar_expect_tx(&full[s], has_meas ? kBytesAll : kBytesPhi);
    tma::tensor_g2s_2d(dst, &maps.phi, (int)warp_base, k * ROWS_PHI, &full[s]);
    if (has_meas) {
      tma::tensor_g2s_2d(dst + ROWS_PHI * 32, &maps.h, (int)warp_base, k * ROWS_H, &full[s]);
      tma::tensor_g2s_2d(dst + (ROWS_PHI + ROWS_H) * 32, &maps.real_obs, (int)warp_base, k * M, &full[s]);
      tma::tensor_g2s_2d(dst + (ROWS_PHI + ROWS_H + M) * 32, &maps.computed_obs, (int)warp_base, k * M, &full[s]);
    }
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kWStages; ++s)
      if (s < io.steps) issue(s, s);
  }

  double x[N], P[SN];  // hybrid: x, P (packed upper);  SRIF: b, R
  if (active) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = io.vec[(int64_t)i * io.nf + tid];
    if constexpr (SRIF) {
#pragma unroll
      for (int i = 0; i < N * N; ++i) P[i] = io.mat[(int64_t)i * io.nf + tid];
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i; j < N; ++j) P[sym_idx<N>(i, j)] = io.mat[(int64_t)(i * N + j) * io.nf + tid];
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
  int s = 0;
  uint32_t phase = 0;
  for (int k = 0; k < io.steps; ++k) {
    const unsigned fl = io.flags ? io.flags[k] : (unsigned)GKB_F_MEAS;
    const bool has_meas = (fl & GKB_F_MEAS) != 0, ekf = (fl & GKB_F_EKF) != 0;
    tma::mbar_wait(&full[s], phase);
    const double* col = ring + (size_t)s * ROWS * 32 + lane;
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
    if (lane == 0 && k + kWStages < io.steps) issue(k + kWStages, s);
    if (++s == kWStages) { s = 0; phase ^= 1u; }
    NlOut<N, M> o;
    int err;
    if constexpr (SRIF) {
      if (!active) {  // keep the padding lanes' Phi invertible (zero-filled by the TMA)
#pragma unroll
        for (int i = 0; i < N; ++i) Phi[i * N + i] = 1.0;
      }
      err = srif_step<N, M>(md, x, P, Phi, Ht, ro, co, has_meas, o);
    } else {
      err = hybrid_step<N, M>(md, x, P, Phi, Ht, ro, co, nullptr, has_meas, ekf, false, o);
    }
    if (err != 0 && status == 0) status = err;
  }
  if constexpr (SRIF) {
    // read-outs of the last estimate: State() = inv(R) b (srif.go:223-235), Covariance() = inv(R) inv(R)^T (253-265)
    if (io.o_state != nullptr) {
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
    if (io.o_covar != nullptr) {
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
  }
  if (active && io.status != nullptr && status != 0 && io.status[tid] == 0) io.status[tid] = status;
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
static int launch_nl_shape(const HostModel& hm, const NlIo& io, cudaStream_t s) {
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
  const bool fast = !io.phi_shared && !io.h_shared && io.Gamma == nullptr && !io.every_step && io.Htilde != nullptr &&
                    io.real_obs != nullptr && io.computed_obs != nullptr && (io.nf % 2 == 0) && aligned(io.Phi) &&
                    aligned(io.Htilde) && aligned(io.real_obs) && aligned(io.computed_obs) && io.o_meas == nullptr &&
                    io.o_innov == nullptr && io.o_pred == nullptr && io.o_gain == nullptr && io.o_obsdev == nullptr &&
                    smem <= 110 * 1024 && io.steps >= 2 && io.nf < 0x7fffffffLL;
  const char* path = getenv("GKB_NL_PATH");  // A/B switch for tests and profiling: plain | bulk | tensor
  const bool want_bulk = path && !strcmp(path, "bulk"), want_plain = path && !strcmp(path, "plain");
  const bool srif = hm.kind == GKB_SRIF;
  if (hm.kind != GKB_HYBRID && !srif) return GKB_ERR_UNSUPPORTED;
  if (fast && !want_plain && !(want_bulk && !srif)) {
    NlTensorMaps maps;
    const int64_t st = io.steps;
    if (make_stream_map(&maps.phi, io.Phi, io.nf, st * N * N, N * N) &&
        make_stream_map(&maps.h, io.Htilde, io.nf, st * M * N, M * N) &&
        make_stream_map(&maps.real_obs, io.real_obs, io.nf, st * M, M) &&
        make_stream_map(&maps.computed_obs, io.computed_obs, io.nf, st * M, M)) {
      const size_t wsmem = sizeof(double) * (kThreads / 32) * kWStages * ROWS * 32 + (kThreads / 32) * kWStages * sizeof(uint64_t);
      auto launch = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        kern<<<grid, kThreads, wsmem, s>>>(md, io, maps);
      };
      if (srif) launch(nl_run_wtma_kernel<N, M, true>);
      else launch(nl_run_wtma_kernel<N, M, false>);
      return 0;
    }
  }
  if (srif) {
    srif_run_kernel<N, M><<<grid, kThreads, 0, s>>>(md, io);
    return 0;
  }
  if (fast && want_bulk) {
    auto kern = hybrid_run_tma_kernel<N, M>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    kern<<<grid, kThreads, smem, s>>>(md, io);
    return 0;
  }
  hybrid_run_kernel<N, M><<<grid, kThreads, 0, s>>>(md, io);
  return 0;
}

int launch_nl_run(const HostModel& hm, const NlIo& io, cudaStream_t s) {
#define GKB_CASE(NN, MM) \
  if (hm.n == NN && hm.m == MM) return launch_nl_shape<NN, MM>(hm, io, s);
  GKB_FOR_EACH_SHAPE(GKB_CASE)
#undef GKB_CASE
  return GKB_ERR_UNSUPPORTED;
}

}  // namespace gkb
