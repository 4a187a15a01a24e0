// kernels_setup.cu -- constructor-time algebra on the device (one thread), so that the engine has a
// single implementation of every factorisation: inv(F), inv(Q), inv(R) for NewInformation
// (information.go:38-50), chol(P0), chol(Q), chol(R) for NewSquareRoot (squareroot.go:35-40,
// 100-114) and the AWGN colouring (noise.go:146-153), I0 = inv(P0) for NewInformationFromState
// (information.go:65-81), R0 / b0 / L for NewSRIF (srif.go:20-45).
#include "engine_internal.h"
#include "smallmat.cuh"

namespace gkb {

struct SetupData {
  int ops;
  int status;
  double F[GKB_MAX_N * GKB_MAX_N], Q[GKB_MAX_N * GKB_MAX_N], R[GKB_MAX_M * GKB_MAX_M];
  double x0[GKB_MAX_N], A0[GKB_MAX_N * GKB_MAX_N];
  double Finv[GKB_MAX_N * GKB_MAX_N], Qinv[GKB_MAX_N * GKB_MAX_N], Rinv[GKB_MAX_M * GKB_MAX_M];
  double sqrtQ[GKB_MAX_N * GKB_MAX_N], sqrtR[GKB_MAX_M * GKB_MAX_M];
};

template <int K>
GKB_DEV void ld(double (&dst)[K * K], const double* src) {
#pragma unroll
  for (int i = 0; i < K * K; ++i) dst[i] = src[i];
}
template <int K>
GKB_DEV void st(double* dst, const double (&src)[K * K]) {
#pragma unroll
  for (int i = 0; i < K * K; ++i) dst[i] = src[i];
}

template <int N, int M>
__global__ void setup_kernel(SetupData* d) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int ops = d->ops;
  int status = 0;
  if (ops & kOpFinv) {
    double a[N * N];
    ld<N>(a, d->F);
    (void)inverse_lu<N>(a);  // error only printed by the reference
    st<N>(d->Finv, a);
  }
  if (ops & kOpQinv) {
    double a[N * N];
    ld<N>(a, d->Q);
    (void)inverse_lu<N>(a);
    st<N>(d->Qinv, a);
  }
  if (ops & kOpRinv) {
    double a[M * M];
    ld<M>(a, d->R);
    (void)inverse_lu<M>(a);
    st<M>(d->Rinv, a);
  }
  if (ops & kOpSqrtQ) {
    double a[N * N], L[N * N];
    ld<N>(a, d->Q);
    (void)chol_lower<N>(L, a);  // `ok` ignored like squareroot.go:102-105
    st<N>(d->sqrtQ, L);
  }
  if (ops & kOpSqrtR) {
    double a[M * M], L[M * M];
    ld<M>(a, d->R);
    (void)chol_lower<M>(L, a);
    st<M>(d->sqrtR, L);
  }
  if (ops & kOpFromState) {
    double a[N * N], x[N], y[N];
    ld<N>(a, d->A0);
    if (inverse_lu<N>(a) != 0) {
#pragma unroll
      for (int i = 0; i < N * N; ++i) a[i] = 0.0;  // information.go:69-72
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) a[i * N + j] = a[j * N + i];  // AsSymDense keeps the upper triangle
    }
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = d->x0[i];
    mulvec<N, N>(y, a, x);
    st<N>(d->A0, a);
#pragma unroll
    for (int i = 0; i < N; ++i) d->x0[i] = y[i];
  }
  if (ops & kOpCholA0) {
    double a[N * N], L[N * N];
    ld<N>(a, d->A0);
    (void)chol_lower<N>(L, a);
    st<N>(d->A0, L);
  }
  if (ops & kOpSrifInit) {
    double a[N * N], L[N * N], x[N], y[N];
#pragma unroll
    for (int i = 0; i < N * N; ++i) a[i] = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) a[i * N + i] = 1.0 / d->A0[i * N + i];  // srif.go:23-26
    (void)chol_lower<N>(L, a);
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = d->x0[i];
    mulvec<N, N>(y, L, x);  // b0 = R0 x0
    st<N>(d->A0, L);
#pragma unroll
    for (int i = 0; i < N; ++i) d->x0[i] = y[i];
    // srif.go:39-45: L = chol(R_meas) (kOpSqrtR above), its inverse must exist
    double l[M * M];
    ld<M>(l, d->sqrtR);
    if (inverse_lu<M>(l) != 0) status = GKB_ERR_SINGULAR_R;
  }
  d->status = status;
}

template <int N, int M>
static int run_setup(SetupData* dev, cudaStream_t s) {
  setup_kernel<N, M><<<1, 32, 0, s>>>(dev);
  return 0;
}

int launch_model_setup(HostModel& hm, int ops, double* x0, double* A0, cudaStream_t s) {
  const int n = hm.n, m = hm.m_r;
  SetupData h;
  h.ops = ops;
  h.status = 0;
  for (int i = 0; i < n * n; ++i) { h.F[i] = hm.F[i]; h.Q[i] = hm.Q[i]; }
  for (int i = 0; i < m * m; ++i) h.R[i] = hm.R[i];
  if (x0) for (int i = 0; i < n; ++i) h.x0[i] = x0[i];
  if (A0) for (int i = 0; i < n * n; ++i) h.A0[i] = A0[i];
  // one small device scratch per (host thread, device), kept for the life of the thread
  static thread_local SetupData* dev_cache[64] = {nullptr};
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur < 0 || cur >= 64) return GKB_ERR_CUDA;
  if (!dev_cache[cur] && cudaMalloc(&dev_cache[cur], sizeof(SetupData)) != cudaSuccess) return GKB_ERR_CUDA;
  SetupData* dev = dev_cache[cur];
  cudaMemcpyAsync(dev, &h, sizeof(SetupData), cudaMemcpyHostToDevice, s);
  int rc = GKB_ERR_UNSUPPORTED;
#define GKB_CASE(NN, MM) \
  if (n == NN && m == MM) rc = run_setup<NN, MM>(dev, s);
  GKB_FOR_EACH_LTI_SHAPE(GKB_CASE)
#undef GKB_CASE
  if (rc == 0) {
    cudaMemcpyAsync(&h, dev, sizeof(SetupData), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) rc = GKB_ERR_CUDA;
  }
  if (rc != 0) return rc;
  if (ops & kOpFinv) for (int i = 0; i < n * n; ++i) hm.Finv[i] = h.Finv[i];
  if (ops & kOpQinv) for (int i = 0; i < n * n; ++i) hm.Qinv[i] = h.Qinv[i];
  if (ops & kOpRinv) for (int i = 0; i < m * m; ++i) hm.Rinv[i] = h.Rinv[i];
  if (ops & kOpSqrtQ) for (int i = 0; i < n * n; ++i) hm.sqrtQ[i] = h.sqrtQ[i];
  if (ops & kOpSqrtR) for (int i = 0; i < m * m; ++i) hm.sqrtR[i] = h.sqrtR[i];
  if (ops & (kOpFromState | kOpCholA0 | kOpSrifInit)) {
    if (A0) for (int i = 0; i < n * n; ++i) A0[i] = h.A0[i];
  }
  if (ops & (kOpFromState | kOpSrifInit)) {
    if (x0) for (int i = 0; i < n; ++i) x0[i] = h.x0[i];
  }
  return h.status;
}

}  // namespace gkb
