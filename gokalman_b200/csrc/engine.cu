// engine.cu -- host runtime and C-ABI of libgokalman_b200.so (see include/gokalman_b200.h).
//
// A handle owns the device-resident state of a batch of filters (SoA [component][filter]), a host
// copy of the shared model and growable device staging buffers for host-side callers.  Calls with
// host pointers copy in, launch, copy out and synchronise; calls with device pointers only enqueue
// work on the handle's stream (the legacy default stream unless gkb_set_stream is used), so a
// caller can bracket them with its own CUDA events.  There is no CPU path anywhere in this file.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "engine_internal.h"

using namespace gkb;

namespace {

thread_local std::string g_last_error;
thread_local float g_last_ms = 0.f;
thread_local int g_last_launches = 0;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define GKB_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(GKB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) {
      if (cudaMalloc(&p, bytes) != cudaSuccess) return fail(GKB_ERR_CUDA, "cudaMalloc(%zu) failed", bytes);
      want = bytes;
    }
    cap = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() { return reinterpret_cast<T*>(p); }
};

// CUDA-event bracket around the kernels of the last call on this thread.  The events are kept so
// that gkb_last_kernel_ms() can be asked later, also after an asynchronous (device-pointer) call.
// Events belong to the device that was current when they were created and can only be recorded on streams of
// that device: one set per device (a host thread may drive handles on several GPUs), created on first use.
struct LastTiming {
  cudaEvent_t e0 = nullptr, e1 = nullptr, m0 = nullptr, m1 = nullptr;  // whole call / dominant kernel
};
constexpr int kMaxDevices = 64;
thread_local LastTiming g_timing_dev[kMaxDevices];
thread_local struct { int device = -1; bool valid = false, main_valid = false; } g_timing;

struct Timer {
  cudaStream_t s;
  LastTiming* t = nullptr;
  bool ok = false;
  // the calling code has made the handle's device current (check_device / cudaSetDevice) before timing
  explicit Timer(cudaStream_t st) : s(st) {
    int dev = -1;
    g_timing.valid = g_timing.main_valid = false;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return;
    t = &g_timing_dev[dev];
    if (!t->e0) {
      if (cudaEventCreate(&t->e0) != cudaSuccess || cudaEventCreate(&t->e1) != cudaSuccess ||
          cudaEventCreate(&t->m0) != cudaSuccess || cudaEventCreate(&t->m1) != cudaSuccess) {
        t->e0 = nullptr;
        (void)cudaGetLastError();
        return;
      }
    }
    g_timing.device = dev;
    ok = cudaEventRecord(t->e0, s) == cudaSuccess;
    if (!ok) (void)cudaGetLastError();  // timing is best effort: never let it leak into the call's error check
  }
  void main_begin() {
    if (ok && cudaEventRecord(t->m0, s) != cudaSuccess) { ok = false; (void)cudaGetLastError(); }
  }
  void main_end() {
    if (!ok) return;
    if (cudaEventRecord(t->m1, s) == cudaSuccess) g_timing.main_valid = true;
    else (void)cudaGetLastError();
  }
  void stop(int launches, bool /*sync*/) {
    g_last_launches = launches;
    if (!ok) return;
    if (cudaEventRecord(t->e1, s) == cudaSuccess) g_timing.valid = true;
    else (void)cudaGetLastError();
  }
};

void sym_from_upper(double* dst, const double* src, int n) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) dst[i * n + j] = (j >= i) ? src[i * n + j] : src[j * n + i];
}

bool is_nil(const double* M, int len) {  // helper.go:50-63
  if (!M) return true;
  for (int i = 0; i < len; ++i)
    if (M[i] != 0.0) return false;
  return true;
}

// vec[i][f] = x0[i] (or per-filter copy), mat[i][f] = A0[i]
__global__ void fill_state_kernel(double* vec, double* mat, const double* x0, int x0_per_filter, const double* A0,
                                  int n, int64_t nf) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  for (int i = 0; i < n; ++i) vec[(int64_t)i * nf + f] = x0_per_filter ? x0[(int64_t)i * nf + f] : x0[i];
  for (int i = 0; i < n * n; ++i) mat[(int64_t)i * nf + f] = A0[i];
}

int check_device(int device) {
  // cudaGetDeviceProperties costs milliseconds: validate each device once per process.
  static int verdict[64] = {0};  // 0 = unknown, 1 = ok
  if (device < 0 || device >= 64) return fail(GKB_ERR_ARG, "device %d out of range", device);
  if (verdict[device] != 1) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
      return fail(GKB_ERR_CUDA, "no CUDA device available: this engine has no CPU fallback");
    if (device >= count) return fail(GKB_ERR_ARG, "device %d out of range (0..%d)", device, count - 1);
    int major = 0, minor = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device) != cudaSuccess)
      return fail(GKB_ERR_CUDA, "cannot query device %d", device);
    if (major != 10)
      return fail(GKB_ERR_CUDA, "device %d is sm_%d%d; the kernels are built for sm_100a only", device, major, minor);
    verdict[device] = 1;
  }
  if (cudaSetDevice(device) != cudaSuccess) return fail(GKB_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  return 0;
}

}  // namespace

struct gkb_filter {
  HostModel hm;
  int64_t nf = 0;
  int device = 0;
  cudaStream_t stream = cudaStreamLegacy;
  int step = 0;
  bool ekf = false;
  bool strict = false;  // gkb_set_strict: reference-order arithmetic (hybrid)
  DevBuf vec, mat, vec0, mat0, status;
  DevBuf replay_w, replay_v, replay_w2;
  int replay_steps = 0;
  bool has_w = false, has_v = false;
  // AWGN on this handle (gkb_set_philox_noise): the samples of each call are generated into the replay arrays
  bool philox = false;
  unsigned long long philox_seed = 0;
  long long philox_offset = 0;
  bool awgn_valid = false;  // chol(Q), chol(R) of the current noise matrices (recomputed after gkb_set_noise)
  double awgn_LQ[GKB_MAX_N * GKB_MAX_N], awgn_LR[GKB_MAX_M * GKB_MAX_M];
  DevBuf in_y, in_u, in_gu, in_a, in_b, in_c, in_d, in_e, in_f;  // staging for host inputs
  DevBuf o_state, o_meas, o_innov, o_covar, o_pred, o_gain, o_obsdev;
  // large-state handles (kernels_tile.cu): filter-major arrays, model kept on the device
  bool tile = false;
  DevBuf tile_model;  // F [n*n], Q [n*n], H [8][n], R [8][8]
  DevBuf sched;       // NLDKF production kernel: task counter + one flag per group of 32 filters
  // host-stream pipeline of gkb_nl_run: two staging sets, a copy stream, "copied" / "consumed" events per set
  DevBuf st_phi[2], st_h[2], st_r[2], st_c[2];
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  DevBuf orbit, od_tab;  // gkb_od_run: reference orbits [6][nf]; per-epoch station + truth-observation tables
  bool has_orbit = false;
};

extern "C" {

const char* gkb_version(void) { return "gokalman_b200 0.1 (sm_100a)"; }
const char* gkb_last_error(void) { return g_last_error.c_str(); }
float gkb_last_kernel_ms(void) {
  if (!g_timing.valid) return -1.f;
  const LastTiming& t = g_timing_dev[g_timing.device];
  if (cudaEventSynchronize(t.e1) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&g_last_ms, t.e0, t.e1) != cudaSuccess) return -1.f;
  return g_last_ms;
}
float gkb_last_main_kernel_ms(void) {
  if (!g_timing.main_valid) return gkb_last_kernel_ms();
  const LastTiming& t = g_timing_dev[g_timing.device];
  float ms = -1.f;
  if (cudaEventSynchronize(t.m1) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, t.m0, t.m1) != cudaSuccess) return -1.f;
  return ms;
}
int gkb_last_kernel_launches(void) { return g_last_launches; }

int gkb_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
  return c;
}

int gkb_shape_supported(int kind, int n, int m) {
  if (kind < GKB_VANILLA || kind > GKB_SRIF) return 0;
  if (kind == GKB_VANILLA && tile_shape_supported(n, m)) return 1;
#define GKB_CASE(NN, MM) \
  if (n == NN && m == MM) return 1;
  if (kind == GKB_HYBRID || kind == GKB_SRIF) {  // n = 7, 8: general kernels only (kernels_nl_big.cu)
    GKB_FOR_EACH_LTI_SHAPE(GKB_CASE)
  } else {  // the LDKF kinds also have n = 7, 8 (batched Update kernels; the Monte Carlo harness stops at n = 6)
    GKB_FOR_EACH_LTI_SHAPE(GKB_CASE)
  }
#undef GKB_CASE
  return 0;
}

static int finish_create(gkb_filter* f, const double* x0, int x0_per_filter, const double* A0) {
  const int n = f->hm.n;
  const int64_t nf = f->nf;
  int rc;
  if ((rc = f->vec.ensure(sizeof(double) * n * nf))) return rc;
  if ((rc = f->mat.ensure(sizeof(double) * n * n * nf))) return rc;
  if ((rc = f->vec0.ensure(sizeof(double) * n * nf))) return rc;
  if ((rc = f->mat0.ensure(sizeof(double) * n * n * nf))) return rc;
  if ((rc = f->status.ensure(sizeof(int32_t) * nf))) return rc;
  DevBuf dx, dA;
  const size_t xbytes = sizeof(double) * (x0_per_filter ? (size_t)n * nf : (size_t)n);
  if ((rc = dx.ensure(xbytes))) return rc;
  if ((rc = dA.ensure(sizeof(double) * n * n))) { dx.release(); return rc; }
  cudaMemcpyAsync(dx.p, x0, xbytes, cudaMemcpyHostToDevice, f->stream);
  cudaMemcpyAsync(dA.p, A0, sizeof(double) * n * n, cudaMemcpyHostToDevice, f->stream);
  fill_state_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, f->stream>>>(f->vec0.as<double>(), f->mat0.as<double>(),
                                                                        dx.as<double>(), x0_per_filter, dA.as<double>(), n, nf);
  cudaMemcpyAsync(f->vec.p, f->vec0.p, sizeof(double) * n * nf, cudaMemcpyDeviceToDevice, f->stream);
  cudaMemcpyAsync(f->mat.p, f->mat0.p, sizeof(double) * n * n * nf, cudaMemcpyDeviceToDevice, f->stream);
  cudaMemsetAsync(f->status.p, 0, sizeof(int32_t) * nf, f->stream);
  cudaError_t e = cudaStreamSynchronize(f->stream);
  dx.release();
  dA.release();
  if (e != cudaSuccess) return fail(GKB_ERR_CUDA, "state initialisation failed: %s", cudaGetErrorString(e));
  return 0;
}

static void destroy_filter(gkb_filter* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  if (f->copy_stream) {
    cudaStreamSynchronize(f->copy_stream);
    cudaStreamDestroy(f->copy_stream);
    for (int b = 0; b < 2; ++b) { cudaEventDestroy(f->ev_h2d[b]); cudaEventDestroy(f->ev_done[b]); }
  }
  DevBuf* bufs[] = {&f->st_phi[0], &f->st_phi[1], &f->st_h[0], &f->st_h[1], &f->st_r[0], &f->st_r[1], &f->st_c[0], &f->st_c[1],
                    &f->replay_w2, &f->orbit, &f->od_tab, &f->sched, &f->tile_model, &f->vec, &f->mat, &f->vec0, &f->mat0, &f->status, &f->replay_w, &f->replay_v, &f->in_y, &f->in_u, &f->in_gu,
                    &f->in_a, &f->in_b, &f->in_c, &f->in_d, &f->in_e, &f->in_f, &f->o_state, &f->o_meas, &f->o_innov,
                    &f->o_covar, &f->o_pred, &f->o_gain, &f->o_obsdev};
  for (DevBuf* b : bufs) b->release();
  delete f;
}

// vec[f][i] = x0[i] (or the per-filter copy), mat[f][i*n+j] = A0[i*n+j]: filter-major (large-state handles)
__global__ void fill_state_tile_kernel(double* vec, double* mat, const double* x0, int x0_per_filter, const double* A0,
                                       int n, int64_t nf) {
  const int64_t total = nf * n * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    mat[idx] = A0[idx % (n * n)];
    if (idx < nf * n) vec[idx] = x0_per_filter ? x0[idx] : x0[idx % n];
  }
}

// NewVanilla for n in {16, 24, ... 64}, m <= 8: the warp-per-filter tensor-core path (kernels_tile.cu).
static int create_tile(int n, int m, int c, int64_t n_filters, int device, const double* x0, int x0_per_filter,
                       const double* P0, const double* F, const double* G, const double* H, const double* Q,
                       const double* R, gkb_filter** out) {
  if (c < 0 || c > GKB_MAX_C) return fail(GKB_ERR_UNSUPPORTED, "control size %d outside 0..%d", c, GKB_MAX_C);
  int rc = check_device(device);
  if (rc) return rc;
  gkb_filter* f = new gkb_filter();
  f->nf = n_filters;
  f->device = device;
  f->tile = true;
  memset(&f->hm, 0, sizeof f->hm);
  f->hm.kind = GKB_VANILLA; f->hm.n = n; f->hm.m = m; f->hm.m_r = m;  // dimensions only: the arrays of hm are too small
  f->hm.c = c;
  f->hm.need_ctrl = !(G == nullptr || c == 0 || is_nil(G, n * c));  // vanilla.go:39
  std::vector<double> model((size_t)2 * n * n + 8 * n + 64 + (size_t)n * GKB_MAX_C, 0.0), A0((size_t)n * n);
  double* mF = model.data();
  double* mQ = mF + n * n;
  double* mH = mQ + n * n;
  double* mR = mH + 8 * n;
  double* mG = mR + 64;
  if (G && c > 0) memcpy(mG, G, sizeof(double) * n * c);
  memcpy(mF, F, sizeof(double) * n * n);
  sym_from_upper(mQ, Q, n);
  memcpy(mH, H, sizeof(double) * m * n);
  for (int a = 0; a < 8; ++a)
    for (int b = 0; b < 8; ++b) mR[a * 8 + b] = (a < m && b < m) ? (b >= a ? R[a * m + b] : R[b * m + a]) : (a == b ? 1.0 : 0.0);
  sym_from_upper(A0.data(), P0, n);
  const size_t xbytes = sizeof(double) * (x0_per_filter ? (size_t)n * n_filters : (size_t)n);
  DevBuf dx, dA;
  if ((rc = f->tile_model.ensure(sizeof(double) * model.size())) || (rc = f->vec.ensure(sizeof(double) * n * n_filters)) ||
      (rc = f->mat.ensure(sizeof(double) * n * n * n_filters)) || (rc = f->vec0.ensure(sizeof(double) * n * n_filters)) ||
      (rc = f->mat0.ensure(sizeof(double) * n * n * n_filters)) || (rc = f->status.ensure(sizeof(int32_t) * n_filters)) ||
      (rc = dx.ensure(xbytes)) || (rc = dA.ensure(sizeof(double) * n * n))) {
    dx.release(); dA.release(); destroy_filter(f);
    return rc;
  }
  cudaMemcpyAsync(f->tile_model.p, model.data(), sizeof(double) * model.size(), cudaMemcpyHostToDevice, f->stream);
  cudaMemcpyAsync(dx.p, x0, xbytes, cudaMemcpyHostToDevice, f->stream);
  cudaMemcpyAsync(dA.p, A0.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice, f->stream);
  fill_state_tile_kernel<<<1184, 256, 0, f->stream>>>(f->vec0.as<double>(), f->mat0.as<double>(), dx.as<double>(),
                                                      x0_per_filter, dA.as<double>(), n, n_filters);
  cudaMemcpyAsync(f->vec.p, f->vec0.p, sizeof(double) * n * n_filters, cudaMemcpyDeviceToDevice, f->stream);
  cudaMemcpyAsync(f->mat.p, f->mat0.p, sizeof(double) * n * n * n_filters, cudaMemcpyDeviceToDevice, f->stream);
  cudaMemsetAsync(f->status.p, 0, sizeof(int32_t) * n_filters, f->stream);
  cudaError_t e = cudaStreamSynchronize(f->stream);
  dx.release();
  dA.release();
  if (e != cudaSuccess) {
    destroy_filter(f);
    return fail(GKB_ERR_CUDA, "state initialisation failed: %s", cudaGetErrorString(e));
  }
  *out = f;
  return 0;
}

int gkb_filter_major(const gkb_filter* f) { return (f && f->tile) ? 1 : 0; }

int gkb_create_lti(int kind, int n, int m, int c, int64_t n_filters, int device, const double* x0, int x0_per_filter,
                   const double* P0, const double* F, const double* G, const double* H, const double* Q,
                   const double* R, gkb_filter** out) {
  if (!out) return fail(GKB_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (kind != GKB_VANILLA && kind != GKB_PREDICTOR && kind != GKB_INFORMATION && kind != GKB_SQRT)
    return fail(GKB_ERR_ARG, "gkb_create_lti: kind %d is not an LDKF kind", kind);
  if (!x0 || !P0 || !F || !H || !Q || !R) return fail(GKB_ERR_ARG, "gkb_create_lti: NULL model array");
  if (n_filters < 1) return fail(GKB_ERR_ARG, "n_filters must be >= 1");
  if (kind == GKB_VANILLA && tile_shape_supported(n, m))
    return create_tile(n, m, c, n_filters, device, x0, x0_per_filter, P0, F, G, H, Q, R, out);
  if (c < 0 || c > GKB_MAX_C) return fail(GKB_ERR_UNSUPPORTED, "control size %d outside 0..%d", c, GKB_MAX_C);
  if (!gkb_shape_supported(kind, n, m))
    return fail(GKB_ERR_UNSUPPORTED, "no compiled kernel for n=%d m=%d (kind %d)", n, m, kind);
  int rc = check_device(device);
  if (rc) return rc;
  gkb_filter* f = new gkb_filter();
  f->nf = n_filters;
  f->device = device;
  HostModel& hm = f->hm;
  memset(&hm, 0, sizeof hm);
  hm.kind = kind; hm.n = n; hm.m = m; hm.c = c; hm.m_r = m;
  memcpy(hm.F, F, sizeof(double) * n * n);
  if (G && c > 0) memcpy(hm.G, G, sizeof(double) * n * c);
  hm.need_ctrl = !(G == nullptr || c == 0 || is_nil(G, n * c));  // vanilla.go:39
  memcpy(hm.H, H, sizeof(double) * m * n);
  sym_from_upper(hm.Q, Q, n);
  sym_from_upper(hm.R, R, m);
  double A0[GKB_MAX_N * GKB_MAX_N];
  sym_from_upper(A0, P0, n);
  std::vector<double> x0v(x0, x0 + (x0_per_filter ? (size_t)n * n_filters : (size_t)n));
  int ops = 0;
  if (kind == GKB_INFORMATION) { ops = kOpFinv | kOpQinv | kOpRinv; hm.rinv_dim = m; }
  if (kind == GKB_SQRT) ops = kOpSqrtQ | kOpSqrtR | kOpCholA0;
  if (ops) {
    rc = launch_model_setup(hm, ops, nullptr, A0, f->stream);
    if (rc) { destroy_filter(f); return fail(rc, "model setup failed (%d)", rc); }
  }
  rc = finish_create(f, x0v.data(), x0_per_filter, A0);
  if (rc) { destroy_filter(f); return rc; }
  *out = f;
  return 0;
}

int gkb_create_information_from_state(int n, int m, int c, int64_t n_filters, int device, const double* x0,
                                      const double* P0, const double* F, const double* G, const double* H,
                                      const double* Q, const double* R, gkb_filter** out) {
  if (!out) return fail(GKB_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (!x0 || !P0 || !F || !H || !Q || !R) return fail(GKB_ERR_ARG, "NULL model array");
  if (!gkb_shape_supported(GKB_INFORMATION, n, m)) return fail(GKB_ERR_UNSUPPORTED, "no compiled kernel for n=%d m=%d", n, m);
  int rc = check_device(device);
  if (rc) return rc;
  // I0 = inv(P0) (zeros if singular), i0 = I0 x0 (information.go:65-81), on the device
  HostModel tmp;
  memset(&tmp, 0, sizeof tmp);
  tmp.kind = GKB_INFORMATION; tmp.n = n; tmp.m = m; tmp.m_r = m;
  double A0[GKB_MAX_N * GKB_MAX_N], xv[GKB_MAX_N];
  sym_from_upper(A0, P0, n);
  memcpy(xv, x0, sizeof(double) * n);
  rc = launch_model_setup(tmp, kOpFromState, xv, A0, cudaStreamLegacy);
  if (rc) return fail(rc, "information-from-state setup failed (%d)", rc);
  return gkb_create_lti(GKB_INFORMATION, n, m, c, n_filters, device, xv, 0, A0, F, G, H, Q, R, out);
}

int gkb_create_hybrid(int n, int m, int q, int64_t n_filters, int device, const double* x0, int x0_per_filter,
                      const double* P0, const double* Q, const double* R, gkb_filter** out) {
  if (!out) return fail(GKB_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (!x0 || !P0 || !R) return fail(GKB_ERR_ARG, "NULL model array");
  if (q < 0 || q > GKB_MAX_Q) return fail(GKB_ERR_UNSUPPORTED, "process-noise size %d outside 0..%d", q, GKB_MAX_Q);
  if (n_filters < 1) return fail(GKB_ERR_ARG, "n_filters must be >= 1");
  if (!gkb_shape_supported(GKB_HYBRID, n, m)) return fail(GKB_ERR_UNSUPPORTED, "no compiled kernel for n=%d m=%d", n, m);
  int rc = check_device(device);
  if (rc) return rc;
  gkb_filter* f = new gkb_filter();
  f->nf = n_filters;
  f->device = device;
  HostModel& hm = f->hm;
  memset(&hm, 0, sizeof hm);
  hm.kind = GKB_HYBRID; hm.n = n; hm.m = m; hm.q = q; hm.m_r = m;
  if (Q && q > 0) sym_from_upper(hm.Q, Q, q);
  sym_from_upper(hm.R, R, m);
  double A0[GKB_MAX_N * GKB_MAX_N];
  sym_from_upper(A0, P0, n);
  rc = finish_create(f, x0, x0_per_filter, A0);
  if (rc) { destroy_filter(f); return rc; }
  // A single-filter handle is the reference-shaped use (one HybridKF, Prepare + Update per epoch through the shim): it
  // starts in reference-order arithmetic, so that its estimates are the reference's bit for bit whatever the conditioning
  // of the run.  Batched handles start in the production (FMA) mode; gkb_set_strict switches either.
  f->strict = (n_filters == 1);
  *out = f;
  return 0;
}

int gkb_create_srif(int n, int m, int64_t n_filters, int device, const double* x0, int x0_per_filter,
                    const double* P0, const double* R, int non_tri_r, gkb_filter** out) {
  if (!out) return fail(GKB_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (!x0 || !P0 || !R) return fail(GKB_ERR_ARG, "NULL model array");
  if (x0_per_filter) return fail(GKB_ERR_UNSUPPORTED, "per-filter x0 is not supported for SRIF (b0 = R0 x0 is formed once)");
  if (n_filters < 1) return fail(GKB_ERR_ARG, "n_filters must be >= 1");
  if (!gkb_shape_supported(GKB_SRIF, n, m)) return fail(GKB_ERR_UNSUPPORTED, "no compiled kernel for n=%d m=%d", n, m);
  int rc = check_device(device);
  if (rc) return rc;
  gkb_filter* f = new gkb_filter();
  f->nf = n_filters;
  f->device = device;
  HostModel& hm = f->hm;
  memset(&hm, 0, sizeof hm);
  hm.kind = GKB_SRIF; hm.n = n; hm.m = m; hm.m_r = m; hm.non_tri_r = non_tri_r;
  sym_from_upper(hm.R, R, m);
  double A0[GKB_MAX_N * GKB_MAX_N], xv[GKB_MAX_N];
  memset(A0, 0, sizeof A0);
  for (int i = 0; i < n * n; ++i) A0[i] = P0[i];
  memcpy(xv, x0, sizeof(double) * n);
  rc = launch_model_setup(hm, kOpSqrtR | kOpSrifInit, xv, A0, f->stream);
  if (rc) { destroy_filter(f); return fail(rc, "NewSRIF: sqrt of the measurement noise is not invertible (%d)", rc); }
  memcpy(hm.L, hm.sqrtR, sizeof(double) * m * m);  // srif.go:48 keeps L, not its inverse
  rc = finish_create(f, xv, 0, A0);
  if (rc) { destroy_filter(f); return rc; }
  *out = f;
  return 0;
}

void gkb_destroy(gkb_filter* f) { destroy_filter(f); }

int64_t gkb_n_filters(const gkb_filter* f) { return f ? f->nf : 0; }
int gkb_step(const gkb_filter* f) { return f ? f->step : 0; }

int gkb_set_stream(gkb_filter* f, void* stream) {
  if (!f) return fail(GKB_ERR_ARG, "NULL handle");
  f->stream = stream ? reinterpret_cast<cudaStream_t>(stream) : cudaStreamLegacy;
  return 0;
}

int gkb_set_strict(gkb_filter* f, int on) {
  if (!f) return fail(GKB_ERR_ARG, "NULL handle");
  if (f->hm.kind != GKB_HYBRID && f->hm.kind != GKB_SRIF)
    return fail(GKB_ERR_UNSUPPORTED, "strict (reference-order) arithmetic is built for GKB_HYBRID and GKB_SRIF handles");
  f->strict = on != 0;
  return 0;
}

int gkb_set_state_transition(gkb_filter* f, const double* F) {
  if (!f || !F) return fail(GKB_ERR_ARG, "NULL argument");
  if (f->tile) {  // the model of a large-state handle lives on the device: F [n*n] first
    cudaSetDevice(f->device);
    GKB_CUDA(cudaMemcpy(f->tile_model.p, F, sizeof(double) * f->hm.n * f->hm.n, cudaMemcpyHostToDevice));
    return 0;
  }
  cudaSetDevice(f->device);
  memcpy(f->hm.F, F, sizeof(double) * f->hm.n * f->hm.n);
  if (f->hm.kind == GKB_INFORMATION) {  // information.go:117-123
    int rc = launch_model_setup(f->hm, kOpFinv, nullptr, nullptr, f->stream);
    if (rc) return fail(rc, "inv(F) setup failed");
  }
  return 0;
}

int gkb_set_input_control(gkb_filter* f, int c, const double* G) {
  if (!f) return fail(GKB_ERR_ARG, "NULL handle");
  if (f->tile) {
    if (c < 0 || c > GKB_MAX_C) return fail(GKB_ERR_UNSUPPORTED, "control size %d outside 0..%d", c, GKB_MAX_C);
    cudaSetDevice(f->device);
    const int n = f->hm.n;
    f->hm.c = c;  // needCtrl is not re-evaluated (vanilla.go:99-101)
    if (G && c > 0)
      GKB_CUDA(cudaMemcpy(f->tile_model.as<double>() + 2 * n * n + 8 * n + 64, G, sizeof(double) * n * c, cudaMemcpyHostToDevice));
    return 0;
  }
  if (c < 0 || c > GKB_MAX_C) return fail(GKB_ERR_UNSUPPORTED, "control size %d outside 0..%d", c, GKB_MAX_C);
  f->hm.c = c;  // needCtrl is not re-evaluated (vanilla.go:99-101)
  if (G && c > 0) memcpy(f->hm.G, G, sizeof(double) * f->hm.n * c);
  return 0;
}

int gkb_set_measurement_matrix(gkb_filter* f, int m, const double* H) {
  if (!f || !H) return fail(GKB_ERR_ARG, "NULL argument");
  if (f->tile) {
    if (!tile_shape_supported(f->hm.n, m)) return fail(GKB_ERR_UNSUPPORTED, "no large-state kernel for n=%d m=%d", f->hm.n, m);
    cudaSetDevice(f->device);
    const int n = f->hm.n;
    std::vector<double> Hp((size_t)8 * n, 0.0);  // rows >= m stay zero
    memcpy(Hp.data(), H, sizeof(double) * m * n);
    GKB_CUDA(cudaMemcpy(f->tile_model.as<double>() + 2 * n * n, Hp.data(), sizeof(double) * 8 * n, cudaMemcpyHostToDevice));
    f->hm.m = m;
    return 0;
  }
  if (!gkb_shape_supported(f->hm.kind, f->hm.n, m)) return fail(GKB_ERR_UNSUPPORTED, "no compiled kernel for n=%d m=%d", f->hm.n, m);
  f->hm.m = m;
  memcpy(f->hm.H, H, sizeof(double) * m * f->hm.n);
  return 0;
}

int gkb_set_noise(gkb_filter* f, const double* Q, int m_r, const double* R) {
  if (!f || !R) return fail(GKB_ERR_ARG, "NULL argument");
  if (f->tile) {
    if (m_r < 1 || m_r > 8) return fail(GKB_ERR_UNSUPPORTED, "R dimension %d outside 1..8", m_r);
    cudaSetDevice(f->device);
    const int n = f->hm.n;
    if (Q) {
      std::vector<double> Qs((size_t)n * n);
      sym_from_upper(Qs.data(), Q, n);
      GKB_CUDA(cudaMemcpy(f->tile_model.as<double>() + n * n, Qs.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    }
    double Rp[64];
    for (int a = 0; a < 8; ++a)
      for (int b = 0; b < 8; ++b)
        Rp[a * 8 + b] = (a < m_r && b < m_r) ? (b >= a ? R[a * m_r + b] : R[b * m_r + a]) : (a == b ? 1.0 : 0.0);
    GKB_CUDA(cudaMemcpy(f->tile_model.as<double>() + 2 * n * n + 8 * n, Rp, sizeof Rp, cudaMemcpyHostToDevice));
    f->hm.m_r = m_r;
    return 0;
  }
  if (m_r < 1 || m_r > GKB_MAX_M) return fail(GKB_ERR_UNSUPPORTED, "R dimension %d outside 1..%d", m_r, GKB_MAX_M);
  if (f->hm.kind == GKB_SRIF) return fail(GKB_ERR_UNSUPPORTED, "noise not yet supported for SRIF (srif.go:77-79 panics)");
  cudaSetDevice(f->device);
  const int qd = (f->hm.kind == GKB_HYBRID) ? f->hm.q : f->hm.n;
  if (Q && qd > 0) sym_from_upper(f->hm.Q, Q, qd);
  sym_from_upper(f->hm.R, R, m_r);
  f->hm.m_r = m_r;
  if (f->hm.kind == GKB_SQRT) {  // squareroot.go:100-114
    if (!gkb_shape_supported(GKB_SQRT, f->hm.n, m_r)) return fail(GKB_ERR_UNSUPPORTED, "no compiled kernel for n=%d m=%d", f->hm.n, m_r);
    int rc = launch_model_setup(f->hm, kOpSqrtQ | kOpSqrtR, nullptr, nullptr, f->stream);
    if (rc) return fail(rc, "chol(Q), chol(R) setup failed");
  }
  // GKB_INFORMATION: inv(Q), inv(R) deliberately left stale (information.go:136-138)
  f->has_w = f->has_v = false;
  f->replay_steps = 0;
  f->philox = false;  // SetNoise replaces the Noise object: an AWGN one is re-armed by gkb_set_philox_noise
  f->awgn_valid = false;
  return 0;
}

int gkb_set_philox_noise(gkb_filter* f, uint64_t seed, int64_t filter_offset) {
  if (!f) return fail(GKB_ERR_ARG, "NULL handle");
  if (f->tile) return fail(GKB_ERR_UNSUPPORTED, "large-state handles (n=%d) carry Noiseless noise only", f->hm.n);
  if (f->hm.kind == GKB_HYBRID || f->hm.kind == GKB_SRIF)
    return fail(GKB_ERR_UNSUPPORTED, "the NLDKF kinds never draw noise samples (hybrid.go / srif.go call neither Process nor Measurement)");
  f->philox = true;
  f->philox_seed = seed;
  f->philox_offset = filter_offset;
  f->has_w = f->has_v = false;
  f->replay_steps = 0;
  return 0;
}

namespace {
// chol(Q), chol(R) of a model for the AWGN colouring (distmv.NewNormal, noise.go:146-153); GKB_ERR_ARG when either is
// not positive definite (NewAWGN panics there, noise.go:149-156).
int awgn_factors(const HostModel& hm, double* LQ, double* LR, cudaStream_t s) {
  HostModel t = hm;
  t.m = hm.m_r;
  int rc = launch_model_setup(t, kOpSqrtQ | kOpSqrtR, nullptr, nullptr, s);
  if (rc) return fail(rc, "noise factorisation failed (%d)", rc);
  for (int i = 0; i < hm.n * hm.n; ++i) {
    if (!std::isfinite(t.sqrtQ[i])) return fail(GKB_ERR_ARG, "process noise invalid: Q is not positive definite (noise.go:149-151)");
    LQ[i] = t.sqrtQ[i];
  }
  for (int i = 0; i < hm.m_r * hm.m_r; ++i) {
    if (!std::isfinite(t.sqrtR[i])) return fail(GKB_ERR_ARG, "measurement noise invalid: R is not positive definite (noise.go:154-156)");
    LR[i] = t.sqrtR[i];
  }
  return 0;
}
}  // namespace

int gkb_awgn_sample(int n, int m, const double* Q, const double* R, uint64_t seed, int64_t filter, int step, int device,
                    double* w, double* v, double* w2) {
  if (!Q || !R) return fail(GKB_ERR_ARG, "NULL argument");
  if (n < 1 || n > GKB_MAX_N || m < 1 || m > GKB_MAX_M) return fail(GKB_ERR_UNSUPPORTED, "n=%d m=%d outside 1..%d / 1..%d", n, m, GKB_MAX_N, GKB_MAX_M);
  if (!gkb_shape_supported(GKB_VANILLA, n, m)) return fail(GKB_ERR_UNSUPPORTED, "no compiled setup kernel for n=%d m=%d", n, m);
  int rc = check_device(device);
  if (rc) return rc;
  cudaStream_t s = cudaStreamLegacy;
  HostModel hm;
  memset(&hm, 0, sizeof hm);
  hm.n = n; hm.m = m; hm.m_r = m;
  sym_from_upper(hm.Q, Q, n);
  sym_from_upper(hm.R, R, m);
  double LQ[GKB_MAX_N * GKB_MAX_N], LR[GKB_MAX_M * GKB_MAX_M];
  if ((rc = awgn_factors(hm, LQ, LR, s))) return rc;
  DevBuf d;
  if ((rc = d.ensure(sizeof(double) * (size_t)(2 * n + m)))) return rc;
  double* dw = d.as<double>();
  launch_awgn_fill(n, m, LQ, LR, seed, filter, 1, 1, step, dw, dw + n, dw + n + m, s);
  double h[2 * GKB_MAX_N + GKB_MAX_M];
  cudaError_t e = cudaMemcpy(h, dw, sizeof(double) * (size_t)(2 * n + m), cudaMemcpyDeviceToHost);
  d.release();
  if (e != cudaSuccess) return fail(GKB_ERR_CUDA, "AWGN sample failed: %s", cudaGetErrorString(e));
  if (w) memcpy(w, h, sizeof(double) * n);
  if (v) memcpy(v, h + n, sizeof(double) * m);
  if (w2) memcpy(w2, h + n + m, sizeof(double) * n);
  return 0;
}

int gkb_set_replay_noise(gkb_filter* f, int steps, const double* w, const double* v, int mem) {
  if (!f) return fail(GKB_ERR_ARG, "NULL handle");
  if (f->tile) return fail(GKB_ERR_UNSUPPORTED, "large-state handles (n=%d) carry Noiseless noise only (no replay samples)", f->hm.n);
  if (steps < 1) return fail(GKB_ERR_ARG, "steps must be >= 1");
  cudaSetDevice(f->device);
  const cudaMemcpyKind kind = mem == GKB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  int rc;
  f->has_w = f->has_v = false;
  f->philox = false;
  if (w) {
    const size_t bytes = sizeof(double) * (size_t)steps * f->hm.n * f->nf;
    if ((rc = f->replay_w.ensure(bytes))) return rc;
    GKB_CUDA(cudaMemcpyAsync(f->replay_w.p, w, bytes, kind, f->stream));
    f->has_w = true;
  }
  if (v) {
    const size_t bytes = sizeof(double) * (size_t)steps * f->hm.m * f->nf;
    if ((rc = f->replay_v.ensure(bytes))) return rc;
    GKB_CUDA(cudaMemcpyAsync(f->replay_v.p, v, bytes, kind, f->stream));
    f->has_v = true;
  }
  GKB_CUDA(cudaStreamSynchronize(f->stream));
  f->replay_steps = steps;
  return 0;
}

int gkb_reset(gkb_filter* f) {
  if (!f) return fail(GKB_ERR_ARG, "NULL handle");
  cudaSetDevice(f->device);
  const int n = f->hm.n;
  GKB_CUDA(cudaMemcpyAsync(f->vec.p, f->vec0.p, sizeof(double) * n * f->nf, cudaMemcpyDeviceToDevice, f->stream));
  GKB_CUDA(cudaMemcpyAsync(f->mat.p, f->mat0.p, sizeof(double) * n * n * f->nf, cudaMemcpyDeviceToDevice, f->stream));
  GKB_CUDA(cudaMemsetAsync(f->status.p, 0, sizeof(int32_t) * f->nf, f->stream));
  f->step = 0;
  f->ekf = false;
  return 0;
}

int gkb_get_state(const gkb_filter* f, double* vec, double* mat) {
  if (!f) return fail(GKB_ERR_ARG, "NULL handle");
  cudaSetDevice(f->device);
  const int n = f->hm.n;
  GKB_CUDA(cudaStreamSynchronize(f->stream));
  if (vec) GKB_CUDA(cudaMemcpy(vec, f->vec.p, sizeof(double) * n * f->nf, cudaMemcpyDeviceToHost));
  if (mat) GKB_CUDA(cudaMemcpy(mat, f->mat.p, sizeof(double) * n * n * f->nf, cudaMemcpyDeviceToHost));
  return 0;
}

int gkb_set_state(gkb_filter* f, const double* vec, const double* mat) {
  if (!f) return fail(GKB_ERR_ARG, "NULL handle");
  cudaSetDevice(f->device);
  const int n = f->hm.n;
  if (vec) GKB_CUDA(cudaMemcpy(f->vec.p, vec, sizeof(double) * n * f->nf, cudaMemcpyHostToDevice));
  if (mat) GKB_CUDA(cudaMemcpy(f->mat.p, mat, sizeof(double) * n * n * f->nf, cudaMemcpyHostToDevice));
  return 0;
}

}  // extern "C"

namespace {

// Device-side views of a gkb_outputs request (staged in the handle when the caller passed host
// pointers) and the copy-back afterwards.
struct OutPlan {
  double *state = nullptr, *meas = nullptr, *innov = nullptr, *covar = nullptr, *pred = nullptr, *gain = nullptr,
         *obsdev = nullptr;
  size_t b_state = 0, b_meas = 0, b_innov = 0, b_covar = 0, b_pred = 0, b_gain = 0, b_obsdev = 0;
};

int plan_outputs(gkb_filter* f, const gkb_outputs* out, int steps, int innov_len, OutPlan& pl) {
  if (!out) return 0;
  const int n = f->hm.n, m = f->hm.m;
  const size_t rows = out->every_step ? (size_t)steps : 1;
  const size_t per = sizeof(double) * rows * f->nf;
  pl.b_state = per * n; pl.b_meas = per * m; pl.b_innov = per * innov_len; pl.b_covar = per * n * n;
  pl.b_pred = per * n * n; pl.b_gain = per * n * m; pl.b_obsdev = per * m;
  if (out->mem == GKB_DEVICE) {
    pl.state = out->state; pl.meas = out->meas; pl.innov = out->innov; pl.covar = out->covar;
    pl.pred = out->pred_covar; pl.gain = out->gain; pl.obsdev = out->obs_dev;
    return 0;
  }
  int rc;
#define GKB_STAGE(field, buf, bytes, dst)        \
  if (out->field) {                              \
    if ((rc = f->buf.ensure(bytes))) return rc;  \
    dst = f->buf.as<double>();                   \
  }
  GKB_STAGE(state, o_state, pl.b_state, pl.state)
  GKB_STAGE(meas, o_meas, pl.b_meas, pl.meas)
  GKB_STAGE(innov, o_innov, pl.b_innov, pl.innov)
  GKB_STAGE(covar, o_covar, pl.b_covar, pl.covar)
  GKB_STAGE(pred_covar, o_pred, pl.b_pred, pl.pred)
  GKB_STAGE(gain, o_gain, pl.b_gain, pl.gain)
  GKB_STAGE(obs_dev, o_obsdev, pl.b_obsdev, pl.obsdev)
#undef GKB_STAGE
  return 0;
}

int copy_back(gkb_filter* f, const gkb_outputs* out, const OutPlan& pl) {
  if (!out) return 0;
  if (out->mem == GKB_HOST) {
#define GKB_BACK(field, src, bytes) \
  if (out->field && src) GKB_CUDA(cudaMemcpyAsync(out->field, src, bytes, cudaMemcpyDeviceToHost, f->stream));
    GKB_BACK(state, pl.state, pl.b_state)
    GKB_BACK(meas, pl.meas, pl.b_meas)
    GKB_BACK(innov, pl.innov, pl.b_innov)
    GKB_BACK(covar, pl.covar, pl.b_covar)
    GKB_BACK(pred_covar, pl.pred, pl.b_pred)
    GKB_BACK(gain, pl.gain, pl.b_gain)
    GKB_BACK(obs_dev, pl.obsdev, pl.b_obsdev)
#undef GKB_BACK
    if (out->status)
      GKB_CUDA(cudaMemcpyAsync(out->status, f->status.p, sizeof(int32_t) * f->nf, cudaMemcpyDeviceToHost, f->stream));
  } else if (out->status) {
    GKB_CUDA(cudaMemcpyAsync(out->status, f->status.p, sizeof(int32_t) * f->nf, cudaMemcpyDeviceToDevice, f->stream));
  }
  return 0;
}

// Stage a host input array on the device (or pass a device pointer through).
int stage_in(gkb_filter* f, DevBuf& buf, const void* src, size_t bytes, int in_mem, const void** dev) {
  *dev = nullptr;
  if (!src) return 0;
  if (in_mem == GKB_DEVICE) {
    *dev = src;
    return 0;
  }
  int rc = buf.ensure(bytes);
  if (rc) return rc;
  GKB_CUDA(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, f->stream));
  *dev = buf.p;
  return 0;
}

}  // namespace

extern "C" {

int gkb_update(gkb_filter* f, int steps, const double* y, int y_shared, const double* u, int in_mem,
               const gkb_outputs* out) {
  if (!f) return fail(GKB_ERR_ARG, "NULL handle");
  const HostModel& hm = f->hm;
  if (hm.kind == GKB_HYBRID || hm.kind == GKB_SRIF) return fail(GKB_ERR_ARG, "gkb_update: handle is an NLDKF, use gkb_nl_run");
  if (steps < 1) return fail(GKB_ERR_ARG, "steps must be >= 1");
  if (!y) return fail(GKB_ERR_DIMS, "measurement (y) is NULL");  // checkMatDims(measurement, H) vanilla.go:133-135
  if (hm.need_ctrl && !u) return fail(GKB_ERR_DIMS, "dimensions must agree: control (u) is required because G is not nil");
  if ((hm.kind == GKB_VANILLA || hm.kind == GKB_PREDICTOR || hm.kind == GKB_SQRT) && hm.m_r != hm.m)
    return fail(GKB_ERR_DIMS, "dimensions must agree: H has %d rows but R is %dx%d", hm.m, hm.m_r, hm.m_r);
  if (hm.kind == GKB_INFORMATION && hm.rinv_dim != 1 && hm.rinv_dim != hm.m)
    return fail(GKB_ERR_DIMS, "dimensions must agree: H has %d rows but inv(R) is %dx%d", hm.m, hm.rinv_dim, hm.rinv_dim);
  int rc = 0;
  if (cudaSetDevice(f->device) != cudaSuccess) return fail(GKB_ERR_CUDA, "cudaSetDevice failed");
  const int n = hm.n, m = hm.m;
  if (f->tile) {
    TileIo tio;
    memset(&tio, 0, sizeof tio);
    tio.nf = f->nf; tio.steps = steps; tio.m = m;
    tio.x = f->vec.as<double>();
    tio.P = f->mat.as<double>();
    const void* dy = nullptr;
    if ((rc = stage_in(f, f->in_y, y, sizeof(double) * (size_t)steps * m * (y_shared ? 1 : f->nf), in_mem, &dy))) return rc;
    tio.y = static_cast<const double*>(dy);
    tio.y_shared = y_shared;
    tio.F = f->tile_model.as<double>();
    tio.Q = tio.F + n * n;
    tio.H = tio.Q + n * n;
    tio.R = tio.H + 8 * n;
    if (hm.need_ctrl && hm.c > 0) {  // u is required (checked above); G u once per step for the whole batch
      const void* du = nullptr;
      if ((rc = stage_in(f, f->in_u, u, sizeof(double) * (size_t)steps * hm.c, in_mem, &du))) return rc;
      if ((rc = f->in_gu.ensure(sizeof(double) * (size_t)steps * n))) return rc;
      launch_tile_gu(tio.R + 64, n, hm.c, static_cast<const double*>(du), steps, f->in_gu.as<double>(), f->stream);
      tio.gu = f->in_gu.as<double>();
    }
    OutPlan pl;
    if ((rc = plan_outputs(f, out, steps, m, pl))) return rc;
    tio.every_step = out ? out->every_step : 0;
    tio.o_state = pl.state; tio.o_meas = pl.meas; tio.o_innov = pl.innov; tio.o_covar = pl.covar;
    tio.o_pred = pl.pred; tio.o_gain = pl.gain;
    tio.status = f->status.as<int32_t>();
    const bool tsync = !(in_mem == GKB_DEVICE && (!out || out->mem == GKB_DEVICE));
    Timer tm(f->stream);
    rc = launch_tile_update(tio, n, f->device, f->stream);
    if (rc) return fail(rc, "no large-state kernel for n=%d m=%d", n, m);
    tm.stop(1, tsync);
    GKB_CUDA(cudaGetLastError());
    f->step += steps;
    if ((rc = copy_back(f, out, pl))) return rc;
    if (tsync) GKB_CUDA(cudaStreamSynchronize(f->stream));
    return 0;
  }
  LtiIo io;
  memset(&io, 0, sizeof io);
  io.nf = f->nf;
  io.steps = steps;
  io.step0 = f->step;
  io.vec = f->vec.as<double>();
  io.mat = f->mat.as<double>();
  const void* dy = nullptr;
  const void* du = nullptr;
  const size_t ybytes = sizeof(double) * (size_t)steps * m * (y_shared ? 1 : f->nf);
  if ((rc = stage_in(f, f->in_y, y, ybytes, in_mem, &dy))) return rc;
  if (u && hm.c > 0)
    if ((rc = stage_in(f, f->in_u, u, sizeof(double) * (size_t)steps * hm.c, in_mem, &du))) return rc;
  io.y = static_cast<const double*>(dy);
  io.y_shared = y_shared;
  if (du && hm.need_ctrl) {
    if ((rc = f->in_gu.ensure(sizeof(double) * (size_t)steps * n))) return rc;
    launch_gu(hm.G, n, hm.c, static_cast<const double*>(du), steps, f->in_gu.as<double>(), f->stream);
    io.gu = f->in_gu.as<double>();
  }
  io.w = f->has_w ? f->replay_w.as<double>() : nullptr;
  io.v = f->has_v ? f->replay_v.as<double>() : nullptr;
  io.replay_steps = f->replay_steps;
  if (f->philox) {
    // AWGN: this call's Process / Measurement samples, keyed by (filter, absolute step), into the replay arrays
    if (hm.m_r != m) return fail(GKB_ERR_DIMS, "dimensions must agree: H has %d rows but R is %dx%d", m, hm.m_r, hm.m_r);
    if (!f->awgn_valid) {  // one factorisation per SetNoise, not per Update()
      if ((rc = awgn_factors(hm, f->awgn_LQ, f->awgn_LR, f->stream))) return rc;
      f->awgn_valid = true;
    }
    const double *LQ = f->awgn_LQ, *LR = f->awgn_LR;
    const size_t wb = sizeof(double) * (size_t)steps * n * f->nf, vb = sizeof(double) * (size_t)steps * m * f->nf;
    if ((rc = f->replay_w.ensure(wb)) || (rc = f->replay_v.ensure(vb)) || (rc = f->replay_w2.ensure(wb))) return rc;
    const bool second = hm.kind == GKB_VANILLA;  // only Vanilla.Update calls Process(k) twice (vanilla.go:146,195)
    launch_awgn_fill(n, m, LQ, LR, f->philox_seed, f->philox_offset, f->nf, steps, f->step, f->replay_w.as<double>(),
                     f->replay_v.as<double>(), second ? f->replay_w2.as<double>() : nullptr, f->stream);
    io.w = f->replay_w.as<double>();
    io.v = f->replay_v.as<double>();
    io.w2 = second ? f->replay_w2.as<double>() : nullptr;
    io.step0 = 0;
    io.replay_steps = steps;
  }
  const int innov_len = (hm.kind == GKB_INFORMATION) ? n : m;
  OutPlan pl;
  if ((rc = plan_outputs(f, out, steps, innov_len, pl))) return rc;
  io.every_step = out ? out->every_step : 0;
  io.o_state = pl.state; io.o_meas = pl.meas; io.o_innov = pl.innov; io.o_covar = pl.covar;
  io.o_pred = pl.pred; io.o_gain = pl.gain; io.o_obsdev = nullptr;
  io.status = f->status.as<int32_t>();
  const bool sync = !(in_mem == GKB_DEVICE && (!out || out->mem == GKB_DEVICE));
  Timer tm(f->stream);
  rc = launch_lti_update(hm, io, f->stream);
  if (rc) return fail(rc, "no kernel for kind=%d n=%d m=%d", hm.kind, n, m);
  tm.stop(1, sync);
  GKB_CUDA(cudaGetLastError());
  f->step += steps;
  if ((rc = copy_back(f, out, pl))) return rc;
  if (sync) GKB_CUDA(cudaStreamSynchronize(f->stream));
  return 0;
}

// One launch of the NLDKF kernels over epochs [k0, k0 + len) of a call of `total` epochs: the stream pointers in `io`
// already point at epoch k0; every-step outputs are offset to row k0, final-estimate outputs are written by the
// launch that ends the call.  The state travels between launches through the handle's state arrays.
static int nl_launch_epochs(gkb_filter* f, NlIo io, const OutPlan& pl, int innov_len, int k0, int len, int total) {
  const HostModel& hm = f->hm;
  const int n = hm.n, m = hm.m;
  io.steps = len;
  const bool last = (k0 + len == total);
  const size_t row = (size_t)k0 * f->nf;
  auto at = [&](double* base, int comps) -> double* {
    if (!base) return nullptr;
    if (io.every_step) return base + row * comps;
    return last ? base : nullptr;
  };
  io.o_state = at(pl.state, n); io.o_meas = at(pl.meas, m); io.o_innov = at(pl.innov, innov_len);
  io.o_covar = at(pl.covar, n * n); io.o_pred = at(pl.pred, n * n); io.o_gain = at(pl.gain, n * m);
  io.o_obsdev = at(pl.obsdev, m);
  return launch_nl_run(hm, io, f->stream);
}

int gkb_nl_run(gkb_filter* f, int steps, const uint8_t* flags, const double* Phi, int phi_shared,
               const double* Htilde, int h_shared, const double* real_obs, const double* computed_obs,
               const double* Gamma, int in_mem, const gkb_outputs* out) {
  if (!f) return fail(GKB_ERR_ARG, "NULL handle");
  const HostModel& hm = f->hm;
  if (hm.kind != GKB_HYBRID && hm.kind != GKB_SRIF) return fail(GKB_ERR_ARG, "gkb_nl_run: handle is an LDKF, use gkb_update");
  if (steps < 1) return fail(GKB_ERR_ARG, "steps must be >= 1");
  if (!Phi) return fail(GKB_ERR_LOCKED, "kf is locked (call Prepare() first): Phi is NULL");  // hybrid.go:105-107
  bool any_meas = (flags == nullptr);
  bool any_snc = false;
  std::vector<uint8_t> hflags;
  if (flags) {
    if (in_mem == GKB_DEVICE) {
      hflags.resize(steps);
      GKB_CUDA(cudaMemcpy(hflags.data(), flags, steps, cudaMemcpyDeviceToHost));
    } else {
      hflags.assign(flags, flags + steps);
    }
    for (int k = 0; k < steps; ++k) {
      any_meas = any_meas || (hflags[k] & GKB_F_MEAS);
      any_snc = any_snc || (hflags[k] & GKB_F_SNC);
    }
  }
  if (any_meas && (!Htilde || !real_obs || !computed_obs))
    return fail(GKB_ERR_DIMS, "dimensions must agree: Update epochs need Htilde, real and computed observations");
  if (any_snc && hm.kind == GKB_HYBRID && (!Gamma || hm.q == 0))
    return fail(GKB_ERR_DIMS, "SNC epochs need Gamma and a q x q process noise matrix");
  if (cudaSetDevice(f->device) != cudaSuccess) return fail(GKB_ERR_CUDA, "cudaSetDevice failed");
  const int n = hm.n, m = hm.m;
  int rc;
  NlIo io;
  memset(&io, 0, sizeof io);
  io.nf = f->nf;
  io.steps = steps;
  io.vec = f->vec.as<double>();
  io.mat = f->mat.as<double>();
  io.phi_shared = phi_shared;
  io.h_shared = h_shared;
  if ((rc = f->sched.ensure(sizeof(int) * (size_t)((f->nf + 31) / 32 + 1)))) return rc;
  io.sched = f->sched.as<int>();
  const int innov_len = (hm.kind == GKB_SRIF) ? n : m;
  OutPlan pl;
  if ((rc = plan_outputs(f, out, steps, innov_len, pl))) return rc;
  io.every_step = out ? out->every_step : 0;
  io.strict = f->strict ? 1 : 0;
  io.status = f->status.as<int32_t>();
  const bool sync = !(in_mem == GKB_DEVICE && (!out || out->mem == GKB_DEVICE));

  // ---- host-resident per-filter streams: the call is PCIe-bound (416 B per filter-epoch at n = 6, m = 2 against
  // ~20 B/ns of PCIe), so the epochs are cut into chunks that travel through two staging sets on a copy stream
  // while the kernels of the previous chunk run: the device never waits for more than one chunk, the staging
  // memory is bounded (2 x ~256 MB instead of the whole stream), and the copy engine is kept busy back to back.
  const size_t per_epoch = sizeof(double) * (size_t)f->nf *
                           ((phi_shared ? 0 : n * n) + ((Htilde && !h_shared) ? m * n : 0) + (real_obs ? m : 0) + (computed_obs ? m : 0));
  int chunk = steps;
  if (in_mem != GKB_DEVICE && per_epoch > 0) {
    const size_t target = (size_t)256 << 20;
    chunk = (int)std::max<size_t>(2, target / per_epoch);
    if (const char* e = getenv("GKB_NL_H2D_CHUNK")) chunk = std::max(1, atoi(e));  // tests force tiny chunks
  }
  if (in_mem != GKB_DEVICE && chunk < steps) {
    if (!f->copy_stream) {
      GKB_CUDA(cudaStreamCreateWithFlags(&f->copy_stream, cudaStreamNonBlocking));
      for (int b = 0; b < 2; ++b) {
        GKB_CUDA(cudaEventCreateWithFlags(&f->ev_h2d[b], cudaEventDisableTiming));
        GKB_CUDA(cudaEventCreateWithFlags(&f->ev_done[b], cudaEventDisableTiming));
      }
    }
    // shared (small) streams and the flags are staged whole, once
    const void *dfl = nullptr, *dphi_sh = nullptr, *dh_sh = nullptr, *dg = nullptr;
    if ((rc = stage_in(f, f->in_a, flags, (size_t)steps, in_mem, &dfl))) return rc;
    if (phi_shared && (rc = stage_in(f, f->in_b, Phi, sizeof(double) * (size_t)steps * n * n, in_mem, &dphi_sh))) return rc;
    if (Htilde && h_shared && (rc = stage_in(f, f->in_c, Htilde, sizeof(double) * (size_t)steps * m * n, in_mem, &dh_sh))) return rc;
    if (hm.kind == GKB_HYBRID && Gamma && hm.q > 0)
      if ((rc = stage_in(f, f->in_f, Gamma, sizeof(double) * (size_t)steps * n * hm.q, in_mem, &dg))) return rc;
    const size_t nfz = (size_t)f->nf;
    const size_t b_phi = phi_shared ? 0 : sizeof(double) * n * n * nfz, b_h = (Htilde && !h_shared) ? sizeof(double) * m * n * nfz : 0;
    const size_t b_o = sizeof(double) * m * nfz;
    for (int b = 0; b < 2; ++b) {
      if (b_phi && (rc = f->st_phi[b].ensure(b_phi * chunk))) return rc;
      if (b_h && (rc = f->st_h[b].ensure(b_h * chunk))) return rc;
      if (real_obs && (rc = f->st_r[b].ensure(b_o * chunk))) return rc;
      if (computed_obs && (rc = f->st_c[b].ensure(b_o * chunk))) return rc;
    }
    GKB_CUDA(cudaStreamSynchronize(f->stream));  // the staging sets may still be read by an earlier asynchronous call
    Timer tm(f->stream);
    int launches = 0;
    for (int k0 = 0, c = 0; k0 < steps; k0 += chunk, ++c) {
      const int len = std::min(chunk, steps - k0), b = c & 1;
      cudaStream_t cs = f->copy_stream;
      if (c >= 2) GKB_CUDA(cudaStreamWaitEvent(cs, f->ev_done[b], 0));  // the kernels of chunk c - 2 are done with set b
      if (b_phi) GKB_CUDA(cudaMemcpyAsync(f->st_phi[b].p, Phi + (size_t)k0 * n * n * nfz, b_phi * len, cudaMemcpyHostToDevice, cs));
      if (b_h) GKB_CUDA(cudaMemcpyAsync(f->st_h[b].p, Htilde + (size_t)k0 * m * n * nfz, b_h * len, cudaMemcpyHostToDevice, cs));
      if (real_obs) GKB_CUDA(cudaMemcpyAsync(f->st_r[b].p, real_obs + (size_t)k0 * m * nfz, b_o * len, cudaMemcpyHostToDevice, cs));
      if (computed_obs) GKB_CUDA(cudaMemcpyAsync(f->st_c[b].p, computed_obs + (size_t)k0 * m * nfz, b_o * len, cudaMemcpyHostToDevice, cs));
      GKB_CUDA(cudaEventRecord(f->ev_h2d[b], cs));
      GKB_CUDA(cudaStreamWaitEvent(f->stream, f->ev_h2d[b], 0));
      NlIo ioc = io;
      ioc.flags = dfl ? static_cast<const uint8_t*>(dfl) + k0 : nullptr;
      ioc.Phi = phi_shared ? static_cast<const double*>(dphi_sh) + (size_t)k0 * n * n : f->st_phi[b].as<double>();
      ioc.Htilde = !Htilde ? nullptr : (h_shared ? static_cast<const double*>(dh_sh) + (size_t)k0 * m * n : f->st_h[b].as<double>());
      ioc.real_obs = real_obs ? f->st_r[b].as<double>() : nullptr;
      ioc.computed_obs = computed_obs ? f->st_c[b].as<double>() : nullptr;
      ioc.Gamma = dg ? static_cast<const double*>(dg) + (size_t)k0 * n * hm.q : nullptr;
      rc = nl_launch_epochs(f, ioc, pl, innov_len, k0, len, steps);
      if (rc) return fail(rc, "no kernel for kind=%d n=%d m=%d", hm.kind, n, m);
      GKB_CUDA(cudaEventRecord(f->ev_done[b], f->stream));
      ++launches;
    }
    tm.stop(launches, true);
    GKB_CUDA(cudaGetLastError());
    f->step += steps;
    if ((rc = copy_back(f, out, pl))) return rc;
    GKB_CUDA(cudaStreamSynchronize(f->stream));
    return 0;
  }

  const void *dfl = nullptr, *dphi = nullptr, *dh = nullptr, *dr = nullptr, *dc = nullptr, *dg = nullptr;
  if ((rc = stage_in(f, f->in_a, flags, (size_t)steps, in_mem, &dfl))) return rc;
  if ((rc = stage_in(f, f->in_b, Phi, sizeof(double) * (size_t)steps * n * n * (phi_shared ? 1 : f->nf), in_mem, &dphi))) return rc;
  if ((rc = stage_in(f, f->in_c, Htilde, sizeof(double) * (size_t)steps * m * n * (h_shared ? 1 : f->nf), in_mem, &dh))) return rc;
  if ((rc = stage_in(f, f->in_d, real_obs, sizeof(double) * (size_t)steps * m * f->nf, in_mem, &dr))) return rc;
  if ((rc = stage_in(f, f->in_e, computed_obs, sizeof(double) * (size_t)steps * m * f->nf, in_mem, &dc))) return rc;
  if (hm.kind == GKB_HYBRID && Gamma && hm.q > 0)
    if ((rc = stage_in(f, f->in_f, Gamma, sizeof(double) * (size_t)steps * n * hm.q, in_mem, &dg))) return rc;
  io.flags = static_cast<const uint8_t*>(dfl);
  io.Phi = static_cast<const double*>(dphi);
  io.Htilde = static_cast<const double*>(dh);
  io.real_obs = static_cast<const double*>(dr);
  io.computed_obs = static_cast<const double*>(dc);
  io.Gamma = static_cast<const double*>(dg);
  Timer tm(f->stream);
  rc = nl_launch_epochs(f, io, pl, innov_len, 0, steps, steps);
  if (rc) return fail(rc, "no kernel for kind=%d n=%d m=%d", hm.kind, n, m);
  tm.stop(1, sync);
  GKB_CUDA(cudaGetLastError());
  f->step += steps;
  if ((rc = copy_back(f, out, pl))) return rc;
  if (sync) GKB_CUDA(cudaStreamSynchronize(f->stream));
  return 0;
}

// ---- orbit-determination inputs on the device ---------------------------------------------------------------
namespace {
int od_params(const gkb_od_config* cfg, OdParams& c) {
  if (!cfg) return fail(GKB_ERR_ARG, "NULL OD configuration");
  if (!(cfg->mu > 0.0) || !(cfg->dt > 0.0) || !(cfg->re >= 0.0)) return fail(GKB_ERR_ARG, "OD configuration: mu, dt must be > 0");
  if (!cfg->station || !cfg->truth_obs) return fail(GKB_ERR_ARG, "OD configuration: station / truth_obs tables are NULL");
  c.mu = cfg->mu;
  c.kj2 = 1.5 * cfg->j2 * cfg->mu * cfg->re * cfg->re;
  c.h = cfg->dt;
  c.sigma[0] = cfg->sigma_range;
  c.sigma[1] = cfg->sigma_rate;
  c.seed = cfg->seed;
  c.filter_offset = cfg->filter_offset;
  return 0;
}
}  // namespace

int gkb_od_synthesize(const gkb_od_config* cfg, int steps, int64_t n_filters, int device, double* Phi, double* Htilde,
                      double* real_obs, double* computed_obs, int mem, double* orbit_out) {
  OdParams c;
  int rc = od_params(cfg, c);
  if (rc) return rc;
  if (!cfg->orbit0 || !Phi || !Htilde || !real_obs || !computed_obs) return fail(GKB_ERR_ARG, "NULL argument");
  if (steps < 1 || n_filters < 1) return fail(GKB_ERR_ARG, "steps and n_filters must be >= 1");
  if ((rc = check_device(device))) return rc;
  cudaStream_t s = cudaStreamLegacy;
  const size_t ob = sizeof(double) * 6 * (size_t)n_filters, tb = sizeof(double) * 8 * (size_t)steps;
  const size_t per = sizeof(double) * (size_t)steps * n_filters;
  DevBuf dorb, dtab, dphi, dh, dr, dc;
  auto cleanup = [&]() { dorb.release(); dtab.release(); dphi.release(); dh.release(); dr.release(); dc.release(); };
  if ((rc = dorb.ensure(ob)) || (rc = dtab.ensure(tb))) { cleanup(); return rc; }
  cudaMemcpyAsync(dorb.p, cfg->orbit0, ob, cfg->orbit_mem == GKB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dtab.p, cfg->station, sizeof(double) * 6 * (size_t)steps, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dtab.as<double>() + 6 * (size_t)steps, cfg->truth_obs, sizeof(double) * 2 * (size_t)steps, cudaMemcpyHostToDevice, s);
  double *pphi = Phi, *ph = Htilde, *pr = real_obs, *pc = computed_obs;
  if (mem != GKB_DEVICE) {
    if ((rc = dphi.ensure(per * 36)) || (rc = dh.ensure(per * 12)) || (rc = dr.ensure(per * 2)) || (rc = dc.ensure(per * 2))) { cleanup(); return rc; }
    pphi = dphi.as<double>(); ph = dh.as<double>(); pr = dr.as<double>(); pc = dc.as<double>();
  }
  Timer tm(s);
  launch_od_synth(c, n_filters, steps, dorb.as<double>(), dtab.as<double>(), dtab.as<double>() + 6 * (size_t)steps, pphi, ph, pr, pc, s);
  tm.stop(1, true);
  if (mem != GKB_DEVICE) {
    cudaMemcpyAsync(Phi, pphi, per * 36, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(Htilde, ph, per * 12, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(real_obs, pr, per * 2, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(computed_obs, pc, per * 2, cudaMemcpyDeviceToHost, s);
  }
  if (orbit_out) cudaMemcpyAsync(orbit_out, dorb.p, ob, mem == GKB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);  // the scratch tables are freed below: always complete the work first
  cleanup();
  if (e != cudaSuccess) return fail(GKB_ERR_CUDA, "OD synthesis failed: %s", cudaGetErrorString(e));
  return 0;
}

int gkb_od_run(gkb_filter* f, const gkb_od_config* cfg, int steps, const uint8_t* flags, const gkb_outputs* out) {
  if (!f) return fail(GKB_ERR_ARG, "NULL handle");
  const HostModel& hm = f->hm;
  if (hm.kind != GKB_HYBRID || hm.n != 6 || hm.m != 2)
    return fail(GKB_ERR_UNSUPPORTED, "gkb_od_run needs a GKB_HYBRID handle with n = 6, m = 2 (position / velocity state, range + range-rate)");
  OdParams c;
  int rc = od_params(cfg, c);
  if (rc) return rc;
  if (steps < 1) return fail(GKB_ERR_ARG, "steps must be >= 1");
  if (!cfg->orbit0 && !f->has_orbit) return fail(GKB_ERR_ARG, "gkb_od_run: no reference orbits yet (orbit0 is NULL on the first call)");
  if (out && (out->meas || out->innov || out->pred_covar || out->gain || out->obs_dev))
    return fail(GKB_ERR_UNSUPPORTED, "gkb_od_run writes state / covar / status only");
  if (flags)
    for (int k = 0; k < steps; ++k)
      if (flags[k] & GKB_F_SNC) return fail(GKB_ERR_UNSUPPORTED, "gkb_od_run: SNC epochs are not supported");
  if (cudaSetDevice(f->device) != cudaSuccess) return fail(GKB_ERR_CUDA, "cudaSetDevice failed");
  const size_t ob = sizeof(double) * 6 * (size_t)f->nf;
  if ((rc = f->orbit.ensure(ob))) return rc;
  if (cfg->orbit0) {
    GKB_CUDA(cudaMemcpyAsync(f->orbit.p, cfg->orbit0, ob, cfg->orbit_mem == GKB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, f->stream));
    f->has_orbit = true;
  }
  // per-epoch tables: station [steps][6], truth observation [steps][2], flags [steps] (after them, 8-byte aligned)
  const size_t tb = sizeof(double) * 8 * (size_t)steps;
  if ((rc = f->od_tab.ensure(tb + (size_t)steps + 8))) return rc;
  double* dst = f->od_tab.as<double>();
  GKB_CUDA(cudaMemcpyAsync(dst, cfg->station, sizeof(double) * 6 * (size_t)steps, cudaMemcpyHostToDevice, f->stream));
  GKB_CUDA(cudaMemcpyAsync(dst + 6 * (size_t)steps, cfg->truth_obs, sizeof(double) * 2 * (size_t)steps, cudaMemcpyHostToDevice, f->stream));
  uint8_t* dfl = nullptr;
  if (flags) {
    dfl = reinterpret_cast<uint8_t*>(dst + 8 * (size_t)steps);
    GKB_CUDA(cudaMemcpyAsync(dfl, flags, (size_t)steps, cudaMemcpyHostToDevice, f->stream));
  }
  NlIo io;
  memset(&io, 0, sizeof io);
  io.nf = f->nf;
  io.steps = steps;
  io.vec = f->vec.as<double>();
  io.mat = f->mat.as<double>();
  io.flags = dfl;
  io.strict = f->strict ? 1 : 0;
  // Final-estimate outputs into PINNED host buffers are written by the kernel itself (mapped host memory: every group's
  // rows leave over PCIe as its last chunk ends, under the compute of the other groups) instead of being staged in HBM
  // and copied after the kernel.  Pageable buffers, every-step outputs and GKB_OD_STAGED_OUTPUTS=1 take the staged path.
  gkb_outputs out_staged;
  double *direct_state = nullptr, *direct_covar = nullptr;
  if (out && out->mem == GKB_HOST && !out->every_step && getenv("GKB_OD_STAGED_OUTPUTS") == nullptr) {
    auto mapped = [](void* p) -> double* {
      cudaPointerAttributes a;
      if (p == nullptr || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
      return (a.type == cudaMemoryTypeHost && a.devicePointer != nullptr) ? static_cast<double*>(a.devicePointer) : nullptr;
    };
    direct_state = mapped(out->state);
    direct_covar = mapped(out->covar);
    if (direct_state || direct_covar) {
      out_staged = *out;
      if (direct_state) out_staged.state = nullptr;
      if (direct_covar) out_staged.covar = nullptr;
      out = &out_staged;
    }
  }
  OutPlan pl;
  if ((rc = plan_outputs(f, out, steps, hm.m, pl))) return rc;
  io.every_step = out ? out->every_step : 0;
  io.o_state = direct_state ? direct_state : pl.state;
  io.o_covar = direct_covar ? direct_covar : pl.covar;
  io.status = f->status.as<int32_t>();
  if ((rc = f->sched.ensure(sizeof(int) * (size_t)((f->nf + 31) / 32 + 1)))) return rc;
  io.sched = f->sched.as<int>();
  const bool sync = !(out && out->mem == GKB_DEVICE);
  Timer tm(f->stream);
  rc = launch_od_run(hm, c, io, f->orbit.as<double>(), dst, dst + 6 * (size_t)steps, f->stream);
  if (rc) return fail(rc, "no fused OD kernel for this handle");
  tm.stop(1, sync);
  GKB_CUDA(cudaGetLastError());
  f->step += steps;
  if ((rc = copy_back(f, out, pl))) return rc;
  // the pageable-host table copies above are staged synchronously by the runtime, so `cfg`'s arrays may be reused
  if (sync) GKB_CUDA(cudaStreamSynchronize(f->stream));
  return 0;
}

// ---- SmoothAll -----------------------------------------------------------------------------------------
int gkb_smooth_all(int n, int steps, int64_t n_filters, int device, const double* Phi, int phi_shared, double* state,
                   double* covar, int mem, int32_t* status) {
  if (!Phi || !state || !covar) return fail(GKB_ERR_ARG, "NULL argument");
  if (n < 1 || n > GKB_MAX_N) return fail(GKB_ERR_UNSUPPORTED, "no compiled smoothing kernel for n=%d", n);
  if (steps < 1 || n_filters < 1) return fail(GKB_ERR_ARG, "steps and n_filters must be >= 1");
  int rc = check_device(device);
  if (rc) return rc;
  cudaStream_t s = cudaStreamLegacy;
  const size_t pb = sizeof(double) * (size_t)steps * n * n * (phi_shared ? 1 : n_filters);
  const size_t xb = sizeof(double) * (size_t)steps * n * n_filters, cb = xb * n, sb = sizeof(int32_t) * n_filters;
  Timer tm(s);
  if (mem == GKB_DEVICE) {
    if (status) GKB_CUDA(cudaMemsetAsync(status, 0, sb, s));
    rc = launch_smooth_all(n, n_filters, steps, Phi, phi_shared, state, covar, status, s);
    tm.stop(1, false);
    GKB_CUDA(cudaGetLastError());
    return rc ? fail(rc, "no smoothing kernel for n=%d", n) : 0;
  }
  DevBuf dphi, dx, dP, dst;
  auto cleanup = [&]() { dphi.release(); dx.release(); dP.release(); dst.release(); };
  if ((rc = dphi.ensure(pb)) || (rc = dx.ensure(xb)) || (rc = dP.ensure(cb)) || (rc = dst.ensure(sb))) { cleanup(); return rc; }
  cudaMemcpyAsync(dphi.p, Phi, pb, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dx.p, state, xb, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dP.p, covar, cb, cudaMemcpyHostToDevice, s);
  cudaMemsetAsync(dst.p, 0, sb, s);
  rc = launch_smooth_all(n, n_filters, steps, dphi.as<double>(), phi_shared, dx.as<double>(), dP.as<double>(), dst.as<int32_t>(), s);
  tm.stop(1, true);
  cudaMemcpyAsync(state, dx.p, xb, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(covar, dP.p, cb, cudaMemcpyDeviceToHost, s);
  if (status) cudaMemcpyAsync(status, dst.p, sb, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cleanup();
  if (rc) return fail(rc, "no smoothing kernel for n=%d", n);
  if (e != cudaSuccess) return fail(GKB_ERR_CUDA, "smoothing failed: %s", cudaGetErrorString(e));
  return 0;
}

// ---- HouseholderTransf ---------------------------------------------------------------------------------
int gkb_householder_transf(int n, int m, int64_t count, int device, double* A, int mem) {
  if (!A) return fail(GKB_ERR_ARG, "NULL argument");
  if (!gkb_shape_supported(GKB_SRIF, n, m) && !(n == 2 && m == 3))
    return fail(GKB_ERR_UNSUPPORTED, "no compiled kernel for n=%d m=%d", n, m);
  if (count < 1) return fail(GKB_ERR_ARG, "count must be >= 1");
  int rc = check_device(device);
  if (rc) return rc;
  cudaStream_t s = cudaStreamLegacy;
  const size_t bytes = sizeof(double) * (size_t)(n + m) * (n + 1) * count;
  if (mem == GKB_DEVICE) {
    rc = launch_householder(n, m, count, A, s);
    GKB_CUDA(cudaGetLastError());
    return rc ? fail(rc, "no Householder kernel for n=%d m=%d", n, m) : 0;
  }
  DevBuf d;
  if ((rc = d.ensure(bytes))) return rc;
  cudaMemcpyAsync(d.p, A, bytes, cudaMemcpyHostToDevice, s);
  rc = launch_householder(n, m, count, d.as<double>(), s);
  cudaMemcpyAsync(A, d.p, bytes, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  d.release();
  if (rc) return fail(rc, "no Householder kernel for n=%d m=%d", n, m);
  if (e != cudaSuccess) return fail(GKB_ERR_CUDA, "HouseholderTransf failed: %s", cudaGetErrorString(e));
  return 0;
}

// ---- VanLoan (c2d.go) --------------------------------------------------------------------------------------
int gkb_van_loan(int n, int q, int64_t count, int device, const double* A, int a_shared, const double* Gamma, int g_shared,
                 const double* W, const double* dt, int dt_shared, int mem, double* F, double* Q, int32_t* status) {
  if (!A || !Gamma || !W || !dt || !F || !Q) return fail(GKB_ERR_ARG, "NULL argument");
  if (n < 1 || n > GKB_MAX_N || q < 1 || q > GKB_MAX_N) return fail(GKB_ERR_UNSUPPORTED, "n=%d q=%d outside 1..%d", n, q, GKB_MAX_N);
  if (count < 1) return fail(GKB_ERR_ARG, "count must be >= 1");
  int rc = check_device(device);
  if (rc) return rc;
  cudaStream_t s = cudaStreamLegacy;
  if (mem == GKB_DEVICE) {
    rc = launch_van_loan(n, q, count, A, a_shared, Gamma, g_shared, W, dt, dt_shared, F, Q, status, s);
    GKB_CUDA(cudaGetLastError());
    return rc ? fail(rc, "no Van Loan kernel for n=%d q=%d", n, q) : 0;
  }
  const size_t ab = sizeof(double) * n * n * (a_shared ? 1 : (size_t)count), gb = sizeof(double) * n * q * (g_shared ? 1 : (size_t)count);
  const size_t wb = sizeof(double) * q * q, tb = sizeof(double) * (dt_shared ? 1 : (size_t)count);
  const size_t ob = sizeof(double) * n * n * (size_t)count, sb = sizeof(int32_t) * (size_t)count;
  DevBuf dA, dG, dW, dT, dF, dQ, dS;
  auto cleanup = [&]() { dA.release(); dG.release(); dW.release(); dT.release(); dF.release(); dQ.release(); dS.release(); };
  if ((rc = dA.ensure(ab)) || (rc = dG.ensure(gb)) || (rc = dW.ensure(wb)) || (rc = dT.ensure(tb)) || (rc = dF.ensure(ob)) ||
      (rc = dQ.ensure(ob)) || (rc = dS.ensure(sb))) { cleanup(); return rc; }
  cudaMemcpyAsync(dA.p, A, ab, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dG.p, Gamma, gb, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dW.p, W, wb, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dT.p, dt, tb, cudaMemcpyHostToDevice, s);
  rc = launch_van_loan(n, q, count, dA.as<double>(), a_shared, dG.as<double>(), g_shared, dW.as<double>(), dT.as<double>(),
                       dt_shared, dF.as<double>(), dQ.as<double>(), dS.as<int32_t>(), s);
  cudaMemcpyAsync(F, dF.p, ob, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(Q, dQ.p, ob, cudaMemcpyDeviceToHost, s);
  if (status) cudaMemcpyAsync(status, dS.p, sb, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cleanup();
  if (rc) return fail(rc, "no Van Loan kernel for n=%d q=%d", n, q);
  if (e != cudaSuccess) return fail(GKB_ERR_CUDA, "VanLoan failed: %s", cudaGetErrorString(e));
  return 0;
}

// ---- BatchKF -------------------------------------------------------------------------------------------
int gkb_batch_solve(int n, int m, int steps, int64_t n_filters, int device, const double* R, const double* H, int h_shared,
                    const double* real_obs, const double* computed_obs, int mem, double* xhat0, double* P0,
                    int32_t* status) {
  if (!R || !H || !real_obs || !computed_obs || !xhat0 || !P0) return fail(GKB_ERR_ARG, "NULL argument");
  if (!gkb_shape_supported(GKB_HYBRID, n, m)) return fail(GKB_ERR_UNSUPPORTED, "no compiled kernel for n=%d m=%d", n, m);
  if (steps < 1 || n_filters < 1) return fail(GKB_ERR_ARG, "steps and n_filters must be >= 1");
  int rc = check_device(device);
  if (rc) return rc;
  cudaStream_t s = cudaStreamLegacy;
  double Rs[GKB_MAX_M * GKB_MAX_M];
  sym_from_upper(Rs, R, m);
  const size_t hb = sizeof(double) * (size_t)steps * m * n * (h_shared ? 1 : n_filters);
  const size_t ob = sizeof(double) * (size_t)steps * m * n_filters;
  const size_t xb = sizeof(double) * (size_t)n * n_filters, pb = xb * n, sb = sizeof(int32_t) * n_filters;
  Timer tm(s);
  if (mem == GKB_DEVICE) {
    rc = launch_batch_solve(n, m, Rs, n_filters, steps, H, h_shared, real_obs, computed_obs, xhat0, P0, status, s);
    tm.stop(1, false);
    GKB_CUDA(cudaGetLastError());
    return rc ? fail(rc, "no BatchKF kernel for n=%d m=%d", n, m) : 0;
  }
  DevBuf dh, dr, dc, dx, dP, dst;
  auto cleanup = [&]() { dh.release(); dr.release(); dc.release(); dx.release(); dP.release(); dst.release(); };
  if ((rc = dh.ensure(hb)) || (rc = dr.ensure(ob)) || (rc = dc.ensure(ob)) || (rc = dx.ensure(xb)) || (rc = dP.ensure(pb)) ||
      (rc = dst.ensure(sb))) { cleanup(); return rc; }
  cudaMemcpyAsync(dh.p, H, hb, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dr.p, real_obs, ob, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dc.p, computed_obs, ob, cudaMemcpyHostToDevice, s);
  rc = launch_batch_solve(n, m, Rs, n_filters, steps, dh.as<double>(), h_shared, dr.as<double>(), dc.as<double>(),
                          dx.as<double>(), dP.as<double>(), dst.as<int32_t>(), s);
  tm.stop(1, true);
  cudaMemcpyAsync(xhat0, dx.p, xb, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(P0, dP.p, pb, cudaMemcpyDeviceToHost, s);
  if (status) cudaMemcpyAsync(status, dst.p, sb, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cleanup();
  if (rc) return fail(rc, "no BatchKF kernel for n=%d m=%d", n, m);
  if (e != cudaSuccess) return fail(GKB_ERR_CUDA, "BatchKF solve failed: %s", cudaGetErrorString(e));
  return 0;
}

// ---- Monte Carlo + chi-square ------------------------------------------------------------------------

namespace {
// The derived matrices of a model depend only on the model: the one-thread setup kernel is rerun only when its
// inputs change (a benchmark loop or a parameter sweep calls gkb_mc_chisquare many times with the same model).
struct McSetupCache {
  bool valid = false;
  int ops = 0;
  HostModel in_hm, out_hm;
  double in_x0[GKB_MAX_N], out_x0[GKB_MAX_N];
  double in_A0[GKB_MAX_N * GKB_MAX_N], out_A0[GKB_MAX_N * GKB_MAX_N];
  // runs (or replays) launch_model_setup(hm, ops, x0, A0); x0 / A0 are [GKB_MAX_N] / [GKB_MAX_N^2] arrays
  int run(HostModel& hm, int ops_, double* x0, double* A0, cudaStream_t s) {
    if (ops_ == 0) return 0;
    const bool same = valid && ops == ops_ && memcmp(&in_hm, &hm, sizeof hm) == 0 &&
                      memcmp(in_x0, x0, sizeof in_x0) == 0 && memcmp(in_A0, A0, sizeof in_A0) == 0;
    if (same) {
      hm = out_hm;
      memcpy(x0, out_x0, sizeof out_x0);
      memcpy(A0, out_A0, sizeof out_A0);
      return 0;
    }
    valid = false;
    ops = ops_;
    in_hm = hm;
    memcpy(in_x0, x0, sizeof in_x0);
    memcpy(in_A0, A0, sizeof in_A0);
    const int rc = launch_model_setup(hm, ops_, x0, A0, s);
    if (rc) return rc;
    out_hm = hm;
    memcpy(out_x0, x0, sizeof out_x0);
    memcpy(out_A0, A0, sizeof out_A0);
    valid = true;
    return 0;
  }
};
struct McScratch {
  DevBuf partial, out, u, gu, gu_f, w, v, err;
  McSetupCache setup_truth, setup_filter;
  int device = -1;
  cudaEvent_t done = nullptr;  // multi-device runs: "this shard's sums are in `out`"
  void release() {  // the buffers live on `device`
    if (device >= 0) cudaSetDevice(device);
    DevBuf* bufs[] = {&partial, &out, &u, &gu, &gu_f, &w, &v, &err};
    for (DevBuf* b : bufs) b->release();
    if (done) cudaEventDestroy(done);
    done = nullptr;
    setup_truth.valid = setup_filter.valid = false;
    device = -1;
  }
};
constexpr int kMcMaxShards = 16;
// slot 0 serves gkb_mc_chisquare; gkb_mc_chisquare_multi uses one slot per shard (two shards may share a device)
thread_local McScratch g_mc_slots[kMcMaxShards];
}  // namespace

// The body of gkb_mc_chisquare on one scratch slot.  raw_xstats < 0: the normal call (outputs copied to `out`);
// raw_xstats = 0 / 1: a shard of a multi-device run -- the per-step SUMS (and, with 1, the Mean / StdDev columns) are
// left in g_mc.out [cols][steps] on the device, the error word in g_mc.err, nothing is copied out, no host sync.
static int mc_run(const gkb_mc_config* cfg, const gkb_mc_outputs* out, McScratch& g_mc, int raw_xstats) {
  if (!cfg || !out) return fail(GKB_ERR_ARG, "NULL argument");
  if (!cfg->with_nees && !cfg->with_nis)
    return fail(GKB_ERR_ARG, "Chi Square requires either NEES or NIS or both");  // chisquare.go:17-19
  if (cfg->kind != GKB_VANILLA && cfg->kind != GKB_INFORMATION && cfg->kind != GKB_SQRT)
    return fail(GKB_ERR_ARG, "tested filter kind %d is not an LDKF kind", cfg->kind);
  const int n = cfg->n, m = cfg->m, c = cfg->c, steps = cfg->steps;
  if (!mc_shape_supported(cfg->kind, n, m))
    return fail(GKB_ERR_UNSUPPORTED, "no compiled Monte Carlo kernel for n=%d m=%d", n, m);
  if (c < 0 || c > GKB_MAX_C) return fail(GKB_ERR_UNSUPPORTED, "control size %d outside 0..%d", c, GKB_MAX_C);
  if (cfg->trials < 1 || steps < 1) return fail(GKB_ERR_ARG, "trials and steps must be >= 1");
  if (!cfg->F || !cfg->H || !cfg->Q || !cfg->R || !cfg->x0_truth || !cfg->x0_filter || !cfg->P0)
    return fail(GKB_ERR_ARG, "NULL model array");
  if (cfg->kind == GKB_INFORMATION && cfg->with_nis && n != m)
    return fail(GKB_ERR_DIMS, "NIS of an information filter multiplies an m x m matrix by the n-vector Innovation() (information.go:272-274)");
  if (cfg->noise_mode == GKB_NOISE_REPLAY && (!cfg->w || !cfg->v)) return fail(GKB_ERR_ARG, "replay noise needs w and v");
  int rc = check_device(cfg->device);
  if (rc) return rc;
  cudaStream_t s = cudaStreamLegacy;
  if (g_mc.device != cfg->device) {
    g_mc.release();
    cudaSetDevice(cfg->device);
    g_mc.device = cfg->device;
  }
  // tm: the truth generator (the pure predictor of NewMonteCarloRuns, montecarlo.go:92, with its AWGN's Q, R);
  // hm: the tested filter's own model (chisquare.go:16 takes any LDKF) -- the truth's unless filter_* say otherwise
  HostModel tm, hm;
  memset(&tm, 0, sizeof tm);
  tm.kind = GKB_PREDICTOR; tm.n = n; tm.m = m; tm.c = c; tm.m_r = m; tm.rinv_dim = m;
  memcpy(tm.F, cfg->F, sizeof(double) * n * n);
  if (cfg->G && c > 0) memcpy(tm.G, cfg->G, sizeof(double) * n * c);
  tm.need_ctrl = !(cfg->G == nullptr || c == 0 || is_nil(cfg->G, n * c));
  memcpy(tm.H, cfg->H, sizeof(double) * m * n);
  sym_from_upper(tm.Q, cfg->Q, n);
  sym_from_upper(tm.R, cfg->R, m);
  hm = tm;
  hm.kind = cfg->kind;
  if (cfg->filter_F) memcpy(hm.F, cfg->filter_F, sizeof(double) * n * n);
  if (cfg->filter_G && c > 0) {
    memcpy(hm.G, cfg->filter_G, sizeof(double) * n * c);
    hm.need_ctrl = !is_nil(cfg->filter_G, n * c);
  }
  if (cfg->filter_H) memcpy(hm.H, cfg->filter_H, sizeof(double) * m * n);
  if (cfg->filter_Q) sym_from_upper(hm.Q, cfg->filter_Q, n);
  if (cfg->filter_R) sym_from_upper(hm.R, cfg->filter_R, m);
  McIo io;
  memset(&io, 0, sizeof io);
  sym_from_upper(io.P0, cfg->P0, n);
  memcpy(io.x0_truth, cfg->x0_truth, sizeof(double) * n);
  memcpy(io.x0_filter, cfg->x0_filter, sizeof(double) * n);
  {
    // truth: AWGN colouring chol(Q), chol(R) (distmv.NewNormal, noise.go:146-153)
    double x0s[GKB_MAX_N] = {0}, A0s[GKB_MAX_N * GKB_MAX_N] = {0};
    if ((rc = g_mc.setup_truth.run(tm, kOpSqrtQ | kOpSqrtR, x0s, A0s, s))) return fail(rc, "truth model setup failed (%d)", rc);
    // tested filter: its constructor's derived matrices, from ITS model
    int ops = 0;
    if (cfg->kind == GKB_INFORMATION) ops = kOpFinv | kOpQinv | kOpRinv | (cfg->info_raw_init ? 0 : kOpFromState);
    if (cfg->kind == GKB_SQRT) ops = kOpSqrtQ | kOpSqrtR | kOpCholA0;
    if ((rc = g_mc.setup_filter.run(hm, ops, io.x0_filter, io.P0, s))) return fail(rc, "tested-filter model setup failed (%d)", rc);
  }
  memcpy(io.LQ, tm.sqrtQ, sizeof(double) * n * n);
  memcpy(io.LR, tm.sqrtR, sizeof(double) * m * m);
  if (cfg->noise_mode == GKB_NOISE_PHILOX) {
    for (int i = 0; i < n * n; ++i)
      if (!std::isfinite(io.LQ[i])) return fail(GKB_ERR_ARG, "process noise invalid: Q is not positive definite (noise.go:149-151)");
    for (int i = 0; i < m * m; ++i)
      if (!std::isfinite(io.LR[i])) return fail(GKB_ERR_ARG, "measurement noise invalid: R is not positive definite (noise.go:154-156)");
  }
  io.trials = cfg->trials;
  io.trial_offset = cfg->trial_offset;
  io.steps = steps;
  io.noise_mode = cfg->noise_mode;
  io.seed = cfg->seed;
  io.with_nees = cfg->with_nees;
  io.with_nis = cfg->with_nis;
  io.want_xstats = raw_xstats >= 0 ? raw_xstats : ((out->sum_d || out->sum_dd || out->x_ref) ? 1 : 0);
  const int cols = mc_cols(n, io.want_xstats);
  if (cfg->controls && c > 0 && (tm.need_ctrl || hm.need_ctrl)) {
    // montecarlo.go:98-104 replaces a single control vector by zeros: an all-zero control stream adds
    // +0.0 to every prediction, so it is elided; otherwise G u is formed once per step for all trials --
    // with the truth's G and, when the tested filter carries another G, with that one too.
    bool any = false;
    for (size_t i = 0; i < (size_t)steps * c && !any; ++i) any = cfg->controls[i] != 0.0;
    if (any) {
      if ((rc = g_mc.u.ensure(sizeof(double) * (size_t)steps * c))) return rc;
      GKB_CUDA(cudaMemcpyAsync(g_mc.u.p, cfg->controls, sizeof(double) * (size_t)steps * c, cudaMemcpyHostToDevice, s));
      if (tm.need_ctrl) {
        if ((rc = g_mc.gu.ensure(sizeof(double) * (size_t)steps * n))) return rc;
        launch_gu(tm.G, n, c, g_mc.u.as<double>(), steps, g_mc.gu.as<double>(), s);
        io.gu = g_mc.gu.as<double>();
      }
      const bool same_g = tm.need_ctrl == hm.need_ctrl && memcmp(tm.G, hm.G, sizeof tm.G) == 0;
      if (same_g) {
        io.gu_f = io.gu;
      } else if (hm.need_ctrl) {
        if ((rc = g_mc.gu_f.ensure(sizeof(double) * (size_t)steps * n))) return rc;
        launch_gu(hm.G, n, c, g_mc.u.as<double>(), steps, g_mc.gu_f.as<double>(), s);
        io.gu_f = g_mc.gu_f.as<double>();
      }
    }
  }
  if (cfg->noise_mode == GKB_NOISE_REPLAY) {
    if (cfg->noise_mem == GKB_DEVICE) {
      io.w = cfg->w;
      io.v = cfg->v;
    } else {
      const size_t wb = sizeof(double) * (size_t)steps * n * cfg->trials, vb = sizeof(double) * (size_t)steps * m * cfg->trials;
      if ((rc = g_mc.w.ensure(wb))) return rc;
      if ((rc = g_mc.v.ensure(vb))) return rc;
      GKB_CUDA(cudaMemcpyAsync(g_mc.w.p, cfg->w, wb, cudaMemcpyHostToDevice, s));
      GKB_CUDA(cudaMemcpyAsync(g_mc.v.p, cfg->v, vb, cudaMemcpyHostToDevice, s));
      io.w = g_mc.w.as<double>();
      io.v = g_mc.v.as<double>();
    }
  }
  // optional dumps: device pointers pass through, host pointers are staged
  DevBuf d_tx, d_ty, d_nw, d_nv, d_st;
  const size_t txb = sizeof(double) * (size_t)steps * n * cfg->trials, tyb = sizeof(double) * (size_t)steps * m * cfg->trials;
  auto stage_out = [&](double* user, DevBuf& buf, size_t bytes, double** dev) -> int {
    *dev = nullptr;
    if (!user) return 0;
    if (out->mem == GKB_DEVICE) { *dev = user; return 0; }
    int r = buf.ensure(bytes);
    if (r) return r;
    *dev = buf.as<double>();
    return 0;
  };
  if ((rc = stage_out(out->truth_x, d_tx, txb, &io.truth_x))) return rc;
  if ((rc = stage_out(out->truth_y, d_ty, tyb, &io.truth_y))) return rc;
  if ((rc = stage_out(out->noise_w, d_nw, txb, &io.noise_w))) return rc;
  if ((rc = stage_out(out->noise_v, d_nv, tyb, &io.noise_v))) return rc;
  if (out->status) {
    if (out->mem == GKB_DEVICE) io.status = out->status;
    else {
      if ((rc = d_st.ensure(sizeof(int32_t) * cfg->trials))) return rc;
      io.status = d_st.as<int32_t>();
    }
    GKB_CUDA(cudaMemsetAsync(io.status, 0, sizeof(int32_t) * cfg->trials, s));
  }
  if ((rc = g_mc.err.ensure(sizeof(int32_t)))) return rc;
  io.first_error = g_mc.err.as<int32_t>();
  GKB_CUDA(cudaMemsetAsync(io.first_error, 0, sizeof(int32_t), s));
  const int max_grid = mc_max_grid(cfg->device);
  const size_t pbytes = sizeof(double) * (size_t)max_grid * steps * cols;
  if ((rc = g_mc.partial.ensure(pbytes))) return rc;
  if ((rc = g_mc.out.ensure(sizeof(double) * (size_t)steps * cols))) return rc;
  io.partial = g_mc.partial.as<double>();
  const bool sync = out->mem != GKB_DEVICE;
  Timer timer(s);
  if (steps > kMcChunk) GKB_CUDA(cudaMemsetAsync(io.partial, 0, pbytes, s));  // chunked flush accumulates into the rows
  int grid = 0;
  timer.main_begin();
  rc = launch_mc(tm, hm, io, cfg->device, &grid, s);
  timer.main_end();
  if (rc) return fail(rc, "no Monte Carlo kernel for kind=%d n=%d m=%d", hm.kind, n, m);
  const double scale = (out->sums_only || raw_xstats >= 0) ? 1.0 : 1.0 / (double)cfg->trials;  // stat.Mean, chisquare.go:85-92
  launch_mc_finish(io.partial, grid, steps, cols, scale, g_mc.out.as<double>(), s);
  timer.stop(2, sync);
  GKB_CUDA(cudaGetLastError());
  if (raw_xstats >= 0) return 0;  // a shard of gkb_mc_chisquare_multi: the caller reduces g_mc.out across devices
  const cudaMemcpyKind back = out->mem == GKB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  const double* o = g_mc.out.as<double>();
  const size_t sb = sizeof(double) * (size_t)steps;
  if (out->first_error)  // chisquare.go:40-42 panics when the tested filter's Update fails: the caller must see it
    GKB_CUDA(cudaMemcpyAsync(out->first_error, io.first_error, sizeof(int32_t), back, s));
  if (out->nis) GKB_CUDA(cudaMemcpyAsync(out->nis, o, sb, back, s));
  if (out->nees) GKB_CUDA(cudaMemcpyAsync(out->nees, o + steps, sb, back, s));
  if (io.want_xstats) {
    // device layout [col][steps] -> user layout [steps][n] (small: transposed on the host)
    if (out->mem == GKB_DEVICE) return fail(GKB_ERR_UNSUPPORTED, "sum_d / sum_dd / x_ref are host-only outputs");
    std::vector<double> tmp((size_t)steps * 3 * n);
    GKB_CUDA(cudaMemcpyAsync(tmp.data(), o + (size_t)kMcBaseCols * steps, sizeof(double) * steps * 3 * n, cudaMemcpyDeviceToHost, s));
    GKB_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < steps; ++k)
      for (int i = 0; i < n; ++i) {
        if (out->sum_d) out->sum_d[(size_t)k * n + i] = tmp[(size_t)i * steps + k];
        if (out->sum_dd) out->sum_dd[(size_t)k * n + i] = tmp[(size_t)(n + i) * steps + k];
        if (out->x_ref) out->x_ref[(size_t)k * n + i] = tmp[(size_t)(2 * n + i) * steps + k];
      }
  }
  if (out->mem == GKB_HOST) {
    if (out->truth_x) GKB_CUDA(cudaMemcpyAsync(out->truth_x, io.truth_x, txb, cudaMemcpyDeviceToHost, s));
    if (out->truth_y) GKB_CUDA(cudaMemcpyAsync(out->truth_y, io.truth_y, tyb, cudaMemcpyDeviceToHost, s));
    if (out->noise_w) GKB_CUDA(cudaMemcpyAsync(out->noise_w, io.noise_w, txb, cudaMemcpyDeviceToHost, s));
    if (out->noise_v) GKB_CUDA(cudaMemcpyAsync(out->noise_v, io.noise_v, tyb, cudaMemcpyDeviceToHost, s));
    if (out->status) GKB_CUDA(cudaMemcpyAsync(out->status, io.status, sizeof(int32_t) * cfg->trials, cudaMemcpyDeviceToHost, s));
    GKB_CUDA(cudaStreamSynchronize(s));
  }
  d_tx.release(); d_ty.release(); d_nw.release(); d_nv.release(); d_st.release();
  return 0;
}

int gkb_mc_chisquare(const gkb_mc_config* cfg, const gkb_mc_outputs* out) { return mc_run(cfg, out, g_mc_slots[0], -1); }

// ---- the same over several GPUs of this process, with the collective inside the ABI -----------------------
namespace {
// NCCL is bound at run time (dlopen): the library has no link-time dependency on it, a Python process that imported
// torch already carries torch's bundled libnccl, and a cgo caller gets the system one.
struct NcclApi {
  void* lib = nullptr;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi& nccl_api() {
  static NcclApi api = []() {
    NcclApi a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) return a;
    a.CommInitAll = reinterpret_cast<int (*)(void**, int, const int*)>(dlsym(a.lib, "ncclCommInitAll"));
    a.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(a.lib, "ncclCommDestroy"));
    a.GroupStart = reinterpret_cast<int (*)()>(dlsym(a.lib, "ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<int (*)()>(dlsym(a.lib, "ncclGroupEnd"));
    a.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(a.lib, "ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(a.lib, "ncclGetErrorString"));
    a.ok = a.CommInitAll && a.CommDestroy && a.GroupStart && a.GroupEnd && a.AllReduce;
    return a;
  }();
  return api;
}
// one communicator set per device list, kept for the life of the process (ncclCommInitAll costs ~100 ms)
struct NcclComms {
  std::vector<int> devices;
  std::vector<void*> comms;
};
NcclComms* nccl_comms_for(const int* devices, int n) {
  static std::mutex mu;  // handles are per-thread, this cache is per-process
  std::lock_guard<std::mutex> lock(mu);
  static std::vector<NcclComms*> cache;
  for (NcclComms* c : cache)
    if ((int)c->devices.size() == n && std::equal(devices, devices + n, c->devices.begin())) return c;
  NcclApi& api = nccl_api();
  if (!api.ok) return nullptr;
  NcclComms* c = new NcclComms();
  c->devices.assign(devices, devices + n);
  c->comms.assign(n, nullptr);
  if (api.CommInitAll(c->comms.data(), n, devices) != 0) {
    delete c;
    return nullptr;
  }
  cache.push_back(c);
  return c;
}

// out0[j] += sum_i peer[i][j] (i = 1 .. n-1, in shard order: a fixed, rank-ordered sum) for j < count, through
// peer-memory loads (NVLink / NVSwitch P2P) -- the whole "collective" of this path is 2 x steps doubles.
struct PeerPtrs { const double* p[kMcMaxShards]; int n; };
__global__ void mc_peer_reduce_kernel(double* __restrict__ out0, const __grid_constant__ PeerPtrs peers, int count) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  double s = out0[j];
  for (int i = 1; i < peers.n; ++i) s += peers.p[i][j];
  out0[j] = s;
}
}  // namespace

int gkb_mc_chisquare_multi(const gkb_mc_config* cfg, const int* devices, int n_devices, int reduce, const gkb_mc_outputs* out) {
  if (!cfg || !out || !devices) return fail(GKB_ERR_ARG, "NULL argument");
  if (n_devices < 1 || n_devices > kMcMaxShards) return fail(GKB_ERR_ARG, "n_devices %d outside 1..%d", n_devices, kMcMaxShards);
  if (reduce != GKB_REDUCE_NCCL && reduce != GKB_REDUCE_PEER) return fail(GKB_ERR_ARG, "unknown reduction mode %d", reduce);
  if (out->mem != GKB_HOST) return fail(GKB_ERR_UNSUPPORTED, "gkb_mc_chisquare_multi writes host outputs");
  if (out->truth_x || out->truth_y || out->noise_w || out->noise_v || out->status)
    return fail(GKB_ERR_UNSUPPORTED, "per-trial dumps are single-device outputs (gkb_mc_chisquare)");
  if (cfg->noise_mode != GKB_NOISE_PHILOX) return fail(GKB_ERR_UNSUPPORTED, "multi-device runs draw their noise from Philox (keyed by the global trial index)");
  if (cfg->trials < n_devices) return fail(GKB_ERR_ARG, "fewer trials (%lld) than devices (%d)", (long long)cfg->trials, n_devices);
  if (reduce == GKB_REDUCE_NCCL)
    for (int i = 0; i < n_devices; ++i)
      for (int j = 0; j < i; ++j)
        if (devices[i] == devices[j]) return fail(GKB_ERR_ARG, "NCCL needs distinct devices (device %d is listed twice)", devices[i]);
  const int n = cfg->n, steps = cfg->steps;
  const int xstats = (out->sum_d || out->sum_dd || out->x_ref) ? 1 : 0;
  const int cols = mc_cols(n, xstats);
  const int sum_count = (kMcBaseCols + (xstats ? 2 * n : 0)) * steps;  // the x_ref columns are identical on every device: not summed
  int rc;
  // 1. every device runs its contiguous trial range; nothing here waits for a kernel
  gkb_mc_outputs raw;
  memset(&raw, 0, sizeof raw);
  raw.mem = GKB_DEVICE;
  raw.sums_only = 1;
  const int64_t base = cfg->trials / n_devices, rem = cfg->trials % n_devices;
  for (int i = 0; i < n_devices; ++i) {
    gkb_mc_config ci = *cfg;
    const int64_t lo = (int64_t)i * base + std::min<int64_t>(i, rem);
    ci.device = devices[i];
    ci.trials = base + (i < rem ? 1 : 0);
    ci.trial_offset = cfg->trial_offset + lo;
    McScratch& sc = g_mc_slots[i];
    if ((rc = mc_run(&ci, &raw, sc, xstats))) return rc;
    if (!sc.done) GKB_CUDA(cudaEventCreateWithFlags(&sc.done, cudaEventDisableTiming));
    GKB_CUDA(cudaEventRecord(sc.done, cudaStreamLegacy));
  }
  // 2. the one collective of the path: the per-step sums
  McScratch& s0 = g_mc_slots[0];
  if (n_devices > 1 && reduce == GKB_REDUCE_NCCL) {
    NcclApi& api = nccl_api();
    NcclComms* cm = nccl_comms_for(devices, n_devices);
    if (!api.ok || !cm) return fail(GKB_ERR_CUDA, "NCCL is not available in this process (dlopen libnccl.so.2 / ncclCommInitAll failed): use GKB_REDUCE_PEER");
    int nrc = api.GroupStart();
    for (int i = 0; i < n_devices && nrc == 0; ++i) {
      cudaSetDevice(devices[i]);
      double* b = g_mc_slots[i].out.as<double>();
      nrc = api.AllReduce(b, b, (size_t)sum_count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, cm->comms[i], cudaStreamLegacy);
    }
    const int erc = api.GroupEnd();
    if (nrc != 0 || erc != 0) return fail(GKB_ERR_CUDA, "ncclAllReduce failed: %s", api.GetErrorString ? api.GetErrorString(nrc ? nrc : erc) : "?");
  } else if (n_devices > 1) {
    GKB_CUDA(cudaSetDevice(devices[0]));
    PeerPtrs pp;
    pp.n = n_devices;
    DevBuf staged[kMcMaxShards];
    for (int i = 1; i < n_devices; ++i) {
      GKB_CUDA(cudaStreamWaitEvent(cudaStreamLegacy, g_mc_slots[i].done, 0));
      int can = devices[i] == devices[0] ? 1 : 0;
      if (!can) {
        cudaDeviceCanAccessPeer(&can, devices[0], devices[i]);
        if (can) {
          cudaError_t e = cudaDeviceEnablePeerAccess(devices[i], 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
          else if (e != cudaSuccess) { can = 0; (void)cudaGetLastError(); }
        }
      }
      if (can) {
        pp.p[i] = g_mc_slots[i].out.as<double>();
      } else {  // no P2P mapping between the two devices: bring the 16 KB over with a peer copy
        if ((rc = staged[i].ensure(sizeof(double) * (size_t)sum_count))) return rc;
        GKB_CUDA(cudaMemcpyPeerAsync(staged[i].p, devices[0], g_mc_slots[i].out.p, devices[i], sizeof(double) * (size_t)sum_count, cudaStreamLegacy));
        pp.p[i] = staged[i].as<double>();
      }
    }
    pp.p[0] = s0.out.as<double>();
    mc_peer_reduce_kernel<<<(sum_count + 255) / 256, 256, 0, cudaStreamLegacy>>>(s0.out.as<double>(), pp, sum_count);
    GKB_CUDA(cudaGetLastError());
    GKB_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    for (int i = 1; i < n_devices; ++i) staged[i].release();
  }
  // 3. results from device 0; the error words from every device
  GKB_CUDA(cudaSetDevice(devices[0]));
  std::vector<double> h((size_t)cols * steps);
  GKB_CUDA(cudaMemcpy(h.data(), s0.out.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
  int32_t first = 0;
  for (int i = 0; i < n_devices; ++i) {
    GKB_CUDA(cudaSetDevice(devices[i]));
    int32_t e = 0;
    GKB_CUDA(cudaMemcpy(&e, g_mc_slots[i].err.p, sizeof e, cudaMemcpyDeviceToHost));  // also waits for device i
    if (e != 0 && first == 0) first = e;
  }
  const double scale = out->sums_only ? 1.0 : 1.0 / (double)cfg->trials;  // stat.Mean over ALL trials, chisquare.go:85-92
  for (int k = 0; k < steps; ++k) {
    if (out->nis) out->nis[k] = h[k] * scale;
    if (out->nees) out->nees[k] = h[(size_t)steps + k] * scale;
    for (int i = 0; i < n && xstats; ++i) {
      if (out->sum_d) out->sum_d[(size_t)k * n + i] = h[(size_t)(kMcBaseCols + i) * steps + k];
      if (out->sum_dd) out->sum_dd[(size_t)k * n + i] = h[(size_t)(kMcBaseCols + n + i) * steps + k];
      if (out->x_ref) out->x_ref[(size_t)k * n + i] = h[(size_t)(kMcBaseCols + 2 * n + i) * steps + k];
    }
  }
  if (out->first_error) *out->first_error = first;
  return 0;
}

}  // extern "C"
