// engine_internal.h -- host runtime <-> kernel launchers (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gokalman_b200.h"

namespace gkb {

// (n, m) shapes with compiled one-filter-per-thread kernels.
// (in four groups, so that the heaviest templates -- the fused Monte Carlo kernels -- compile as parallel parts)
#define GKB_FOR_EACH_SHAPE_G0(X) X(1, 1) X(2, 1) X(2, 2) X(3, 1) X(3, 2) X(3, 3) X(4, 1) X(4, 2) X(4, 3)
#define GKB_FOR_EACH_SHAPE_G1(X) X(5, 1) X(5, 2) X(6, 1) X(6, 2) X(6, 3)
#define GKB_FOR_EACH_SHAPE_G2(X) X(7, 1) X(7, 2) X(7, 3)
#define GKB_FOR_EACH_SHAPE_G3(X) X(8, 1) X(8, 2) X(8, 3)
#define GKB_FOR_EACH_SHAPE(X) GKB_FOR_EACH_SHAPE_G0(X) GKB_FOR_EACH_SHAPE_G1(X)

// LDKF kinds (Vanilla / pure predictor / Information / Square root) also get n = 7 and 8 (the north star's "n <= 8"): slower
// (the 8 x 8 intermediates no longer fit the registers: ptxas reports local memory) but correct, same parity bar.
#define GKB_FOR_EACH_BIG_SHAPE(X) GKB_FOR_EACH_SHAPE_G2(X) GKB_FOR_EACH_SHAPE_G3(X)
#define GKB_FOR_EACH_LTI_SHAPE(X) GKB_FOR_EACH_SHAPE(X) GKB_FOR_EACH_BIG_SHAPE(X)

constexpr int kThreads = 128;  // threads per CTA for the register kernels

// Host copy of everything a filter handle knows about its model (row-major, full matrices).
struct HostModel {
  int kind, n, m, c, q;
  int m_r;        // dimension of R as last set (SetNoise may change it independently of H)
  int need_ctrl;  // !IsNil(G) at construction
  int rinv_dim;   // information: dimension of Rinv fixed at construction
  int non_tri_r;
  double F[GKB_MAX_N * GKB_MAX_N];
  double G[GKB_MAX_N * GKB_MAX_C];
  double H[GKB_MAX_M * GKB_MAX_N];
  double Q[GKB_MAX_N * GKB_MAX_N];
  double R[GKB_MAX_M * GKB_MAX_M];
  double Finv[GKB_MAX_N * GKB_MAX_N];
  double Qinv[GKB_MAX_N * GKB_MAX_N];
  double Rinv[GKB_MAX_M * GKB_MAX_M];
  double sqrtQ[GKB_MAX_N * GKB_MAX_N];
  double sqrtR[GKB_MAX_M * GKB_MAX_M];
  double L[GKB_MAX_M * GKB_MAX_M];
};

// Device-side I/O of a batched LDKF update launch (all pointers are device pointers).
struct LtiIo {
  int64_t nf;
  int steps;
  int step0;  // filter step counter at entry: index into the replay noise
  double* vec;  // [n][nf]
  double* mat;  // [n*n][nf]
  const double* y;
  int y_shared;
  const double* gu;  // [steps][n]: G u per step (gu_kernel), nullptr = no control term
  const double* w;  // replay noise [replay_steps][n][nf] or nullptr
  const double* v;  // [replay_steps][m_v][nf] or nullptr
  const double* w2; // Vanilla only: what the SECOND Process(k) call returns (AWGN draws afresh), nullptr = w again
  int replay_steps;
  int every_step;
  double *o_state, *o_meas, *o_innov, *o_covar, *o_pred, *o_gain, *o_obsdev;
  int32_t* status;
};

struct NlIo {
  int64_t nf;
  int steps;
  double* vec;
  double* mat;
  const uint8_t* flags;  // [steps]
  const double* Phi;
  int phi_shared;
  const double* Htilde;
  int h_shared;
  const double* real_obs;
  const double* computed_obs;
  const double* Gamma;  // [steps][n*q] shared or nullptr
  int* sched;           // TMA production path: [1 + ceil(nf / 32)] ints of scheduler scratch (task counter + per-group flags)
  int chunks, chunk_len;  // filled by the launcher: epochs are run in `chunks` chunks of `chunk_len`
  int every_step;
  int strict;  // gkb_set_strict: reference-order arithmetic (filters_strict.cuh), hybrid only
  double *o_state, *o_meas, *o_innov, *o_covar, *o_pred, *o_gain, *o_obsdev;
  int32_t* status;
};

// Large-state (warp-per-filter, tensor-core) Vanilla update: FILTER-MAJOR arrays, model in device memory.
struct TileIo {
  int64_t nf;
  int steps;
  int m;            // true measurement size (<= 8); the kernel pads to one 8-wide tile
  double* x;        // [nf][n]
  double* P;        // [nf][n*n]
  const double* y;  // [steps][nf][m], or [steps][m] when y_shared
  int y_shared;
  const double* F;  // [n*n]
  const double* Q;  // [n*n]
  const double* H;  // [8][n], rows >= m zero
  const double* R;  // [8][8], unit diagonal beyond m
  const double* gu; // [steps][n]: G u per step (tile_gu_kernel), nullptr = no control term
  int every_step;
  double *o_state, *o_meas, *o_innov, *o_covar, *o_pred, *o_gain;  // [rows][nf][C]
  int32_t* status;
};

struct McIo {
  int64_t trials;
  int64_t trial_offset;
  int steps;
  const double* gu;    // device [steps][n]: G u per step of the TRUTH generator, nullptr = no (or all-zero) control
  const double* gu_f;  // the same for the tested filter's own G (== gu when both carry the same G)
  int noise_mode;
  unsigned long long seed;
  const double* w;  // replay [steps][n][trials]
  const double* v;  // replay [steps][m][trials]
  int with_nees, with_nis;
  double* partial;  // [grid][steps][kMcCols] per-CTA partial sums
  int want_xstats;
  double *truth_x, *truth_y, *noise_w, *noise_v;
  int32_t* status;
  int32_t* first_error;  // one word: the most negative status any trial hit, 0 if none
  double x0_truth[GKB_MAX_N];
  double x0_filter[GKB_MAX_N];
  double P0[GKB_MAX_N * GKB_MAX_N];
  double LQ[GKB_MAX_N * GKB_MAX_N];  // chol_lower(Q): colours the process noise
  double LR[GKB_MAX_M * GKB_MAX_M];
};
// columns of McIo::partial per step: NIS, NEES, then sum d_i (n), sum d_i^2 (n), xref_i (n), d = x - xref
constexpr int kMcBaseCols = 2;
constexpr int kMcChunk = 64;   // steps accumulated in shared memory between flushes to McIo::partial (4 KB per CTA at 2 columns)
inline int mc_cols(int n, int want_xstats) { return kMcBaseCols + (want_xstats ? 3 * n : 0); }

// ---- launchers (each returns a gkb_status; GKB_ERR_UNSUPPORTED when the shape is not compiled) ----
// One-thread device kernel for the constructor-time algebra (the same templates the filters use).
enum SetupOps {
  kOpFinv = 1,        // Finv = inv(F)                       information.go:38-41,117-123
  kOpQinv = 2,        // Qinv = inv(Q)                       information.go:43-46
  kOpRinv = 4,        // Rinv = inv(R)                       information.go:47-50
  kOpSqrtQ = 8,       // sqrtQ = chol_lower(Q)               squareroot.go:102-105, distmv.NewNormal
  kOpSqrtR = 16,      // sqrtR = chol_lower(R)               squareroot.go:107-110, srif.go:39-41
  kOpFromState = 32,  // A0 = inv(P0) or 0, x0 = A0 x0       information.go:65-81
  kOpCholA0 = 64,     // A0 = chol_lower(P0)                 squareroot.go:35-40
  kOpSrifInit = 128   // A0 = chol(diag(1/P0_ii)), x0 = A0 x0, check inv(L)   srif.go:20-45
};
// x0 [n] and A0 [n*n] are host arrays, updated in place where an op says so.  Dispatches on
// (hm.n, hm.m_r).  Returns a gkb_status (GKB_ERR_SINGULAR_R when NewSRIF's inverse of L fails).
int launch_model_setup(HostModel& hm, int ops, double* x0, double* A0, cudaStream_t s);
// gu[k][i] = sum_j G[i][j] u[k][j]: the control term, identical for every filter of a batch.
int launch_gu(const double* G_host, int n, int c, const double* u_dev, int steps, double* gu_dev, cudaStream_t s);
int launch_lti_update(const HostModel& hm, const LtiIo& io, cudaStream_t s);
int launch_lti_update_big(const HostModel& hm, const LtiIo& io, cudaStream_t s);  // kernels_lti.cu part 1: n = 7, 8
int launch_nl_run(const HostModel& hm, const NlIo& io, cudaStream_t s);
int launch_nl_run_big(const HostModel& hm, const NlIo& io, cudaStream_t s);  // kernels_nl_big.cu: n = 7, 8
// kernels_nl_tma.cu: the TMA production path of the NLDKF kinds.  0 = launched, 1 = not applicable to this call.
int launch_nl_tma(const HostModel& hm, const NlIo& io, cudaStream_t s);
// SmoothAll (hybrid.go:209-238, srif.go:165-192) over stored [steps][C][nf] histories, in place.
int launch_smooth_all(int n, int64_t nf, int steps, const double* Phi, int phi_shared, double* xs, double* Ps,
                      int32_t* status, cudaStream_t s);
// HouseholderTransf (helper.go:142-172) on `count` matrices [(n+m)*(n+1)][count], in place.
int launch_householder(int n, int m, int64_t count, double* A, cudaStream_t s);
// BatchKF (batch.go:34-79): accumulation over the measurement streams + Solve(), one thread per batch filter.
int launch_batch_solve(int n, int m, const double* R_host, int64_t nf, int steps, const double* H, int h_shared,
                       const double* real_obs, const double* computed_obs, double* xhat0, double* P0, int32_t* status,
                       cudaStream_t s);
// Orbit-determination inputs on the device (kernels_od.cu; OdParams in od_synth.cuh).  orbit [6][nf] is read and
// advanced; station [steps][6] and tobs [steps][2] are device tables shared by the batch.
struct OdParams {
  double mu;        // km^3 / s^2
  double kj2;       // 1.5 J2 mu Re^2
  double h;         // seconds per epoch (one RK4 step)
  double sigma[2];  // standard deviation of the range [km] / range-rate [km/s] measurement noise
  unsigned long long seed;
  long long filter_offset;  // global index of this batch's filter 0 (Philox is keyed by the global index)
};
// rows of the per-(filter, epoch) block od_step produces, in stream order: Phi, Htilde, real, computed
constexpr int kOdPhi = 0, kOdH = 36, kOdReal = 48, kOdComp = 50, kOdRows = 52;
int launch_od_synth(const OdParams& c, int64_t nf, int steps, double* orbit, const double* station, const double* tobs,
                    double* Phi, double* Ht, double* real_obs, double* comp_obs, cudaStream_t s);
// The fused run: synthesis + hybrid CKF / EKF step per epoch, no streams in HBM (n = 6, m = 2 only).
int launch_od_run(const HostModel& hm, const OdParams& c, const NlIo& io, double* orbit, const double* station,
                  const double* tobs, cudaStream_t s);
// AWGN samples on the device (kernels_noise.cu): for filters [0, nf) and steps [step0, step0 + steps), the coloured
// draws w = LQ z[0:n], v = LR z[n:n+m], w2 = LQ z[n+m:2n+m] of z = Philox normals of (seed, filter_offset + filter, step),
// written SoA [steps][component][nf].  LQ [n*n], LR [m*m] are host arrays (lower Cholesky factors).
int launch_awgn_fill(int n, int m, const double* LQ, const double* LR, unsigned long long seed, long long filter_offset,
                     int64_t nf, int steps, int step0, double* w, double* v, double* w2, cudaStream_t s);
// Van Loan c2d (c2d.go:13-75) for `count` systems, one thread each (kernels_c2d.cu).  Device pointers.
int launch_van_loan(int n, int q, int64_t count, const double* A, int a_shared, const double* Gamma, int g_shared,
                    const double* W, const double* dt, int dt_shared, double* F, double* Q, int32_t* status, cudaStream_t s);
// Large-state Vanilla (kernels_tile.cu): n in {16, 24, 32}, m <= 8.
int tile_shape_supported(int n, int m);
int launch_tile_update(const TileIo& io, int n, int device, cudaStream_t s);
// gu[k][i] = sum_j G[i][j] u[k][j] for a large-state handle (G [n][c] and u [steps][c] on the device).
int launch_tile_gu(const double* G_dev, int n, int c, const double* u_dev, int steps, double* gu_dev, cudaStream_t s);
// Upper bound on the CTAs launch_mc will use (rows of McIo::partial to allocate, zero-filled).
int mc_max_grid(int device);
// Picks a persistent grid (SM count x resident CTAs per SM, capped by the work) and launches.
// tm: the truth generator's model (F, G, H; its chol(Q), chol(R) are in io.LQ / io.LR); hm: the tested filter's.
int launch_mc(const HostModel& tm, const HostModel& hm, const McIo& io, int device, int* grid_out, cudaStream_t s);
int mc_shape_supported(int kind, int n, int m);
int launch_mc_finish(const double* partial, int grid, int steps, int cols, double scale, double* out_cols,
                     cudaStream_t s);

}  // namespace gkb
