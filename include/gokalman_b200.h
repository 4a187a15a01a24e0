/* gokalman_b200.h -- C-ABI of the B200-native batched Kalman-filter engine.
 *
 * The reference (ChristopherRabotin/gokalman) is a pure-Go library with no FFI; its extension
 * points are the Go interfaces LDKF / NLDKF (kalman.go:35-60), Estimate (kalman.go:64-72) and
 * Noise (noise.go:13-20).  This header is the boundary a cgo shim binds to put hand-written
 * sm_100a kernels behind those interfaces (INTEGRATION.md shows the shim).  Each entry point cites
 * the reference method it replaces.
 *
 * One handle = a batch of `n_filters` independent filters that share one model, advanced in
 * lockstep; n_filters = 1 is the drop-in for a single Go filter object.  Plain pointers and sizes
 * only.  Every call returns 0 on success or a negative gkb_status; gkb_last_error() gives the
 * message of the calling thread's last failure.  A handle must be used by one host thread at a
 * time (the Go filters are not goroutine-safe either: vanilla.go:217-218).  There is no CPU
 * fallback: every call fails with GKB_ERR_CUDA when no sm_100 device is usable.
 *
 * Array layout ("SoA"): a per-filter quantity with C components over S steps is
 * double[S][C][n_filters] (filter index fastest), so that a warp's loads are coalesced; with
 * n_filters = 1 this is the ordinary row-major [S][C].  Matrices are row-major (n x n -> n*n
 * components).  All arithmetic is FP64.
 */
#ifndef GOKALMAN_B200_H
#define GOKALMAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GKB_MAX_N 8  /* one-filter-per-thread register kernels */
#define GKB_TILE_MAX_N 64 /* large-state (warp-per-filter, FP64 tensor-core) Vanilla: n in {16, 24, 32, 40, 48, 56, 64}, m <= 8 */
#define GKB_TILE_MAX_M 8
#define GKB_MAX_M 3
#define GKB_MAX_C 4
#define GKB_MAX_Q 3

typedef enum gkb_kind {
  GKB_VANILLA = 0,     /* NewVanilla                vanilla.go:21-40    */
  GKB_PREDICTOR = 1,   /* NewPurePredictorVanilla   vanilla.go:43-62    */
  GKB_INFORMATION = 2, /* NewInformation[FromState] information.go:20-81 */
  GKB_SQRT = 3,        /* NewSquareRoot             squareroot.go:21-50 */
  GKB_HYBRID = 4,      /* NewHybridKF               hybrid.go:23-34     */
  GKB_SRIF = 5         /* NewSRIF                   srif.go:14-49       */
} gkb_kind;

typedef enum gkb_status {
  GKB_OK = 0,
  GKB_ERR_DIMS = -1,         /* checkMatDims failures (helper.go:100-130)                  */
  GKB_ERR_SINGULAR_S = -2,   /* vanilla.go:164-167, hybrid.go:150-152                      */
  GKB_ERR_ASYMMETRIC = -3,   /* vanilla.go:207-215 (never raised: P is kept symmetric)     */
  GKB_ERR_LOCKED = -4,       /* hybrid.go:105-107, srif.go:102-104                         */
  GKB_ERR_SINGULAR_PHI = -5, /* srif.go:112-114                                            */
  GKB_ERR_SINGULAR_R = -6,   /* srif.go:228-230 (panic in the reference)                   */
  GKB_ERR_NOISE_RANGE = -7,  /* noise.go:74-76,82-84 (panic in the reference)              */
  GKB_ERR_UNSUPPORTED = -8,  /* shape outside the compiled kernel table                    */
  GKB_ERR_CUDA = -9,         /* CUDA runtime failure / no device                           */
  GKB_ERR_ARG = -10,
  GKB_ERR_NONFINITE = -11    /* a filter produced a non-finite state or covariance         */
} gkb_status;

typedef struct gkb_filter gkb_filter;

/* Where the arrays passed to a call live. */
typedef enum gkb_mem { GKB_HOST = 0, GKB_DEVICE = 1 } gkb_mem;

const char* gkb_version(void);
const char* gkb_last_error(void);
int gkb_device_count(void);
/* 1 if a filter of (kind, n, m) can be created: every kind n <= 8, m <= 3 (n = 5: m <= 2) plus the large-state Vanilla
 * shapes.  gkb_mc_chisquare, gkb_smooth_all, gkb_batch_solve and gkb_householder_transf cover the same n <= 8 table.
 * n = 7, 8 run the same one-filter-per-thread templates with L1-resident spills (correct to the same bar, slower);
 * the TMA production path of the NLDKF kinds is n <= 6. */
int gkb_shape_supported(int kind, int n, int m);

/* ---- construction: replaces NewVanilla / NewPurePredictorVanilla / NewInformation /
 *      NewSquareRoot (LDKF, kalman.go:35-48).  x0 is [n] (shared by all filters) or, when
 *      x0_per_filter != 0, [n][n_filters].  P0 is n x n (for GKB_INFORMATION x0/P0 are the
 *      information state i0 and matrix I0).  G may be NULL (c = 0).  Arrays are host pointers and
 *      are copied.  Only the upper triangle of P0, Q, R is read (mat64.SymDense semantics). */
/* Large-state handles: GKB_VANILLA with n in {16, 24, 32, 40, 48, 56, 64} and m <= 8 (Noiseless noise: no replay samples)
 * runs one WARP per filter on the FP64 tensor-core path with the covariance resident in
 * shared memory.  Such a handle uses FILTER-MAJOR arrays everywhere -- x0 (per filter) [N][n],
 * state [N][n], covar [N][n*n], y [steps][N][m], outputs [steps][N][C] -- so that one filter's
 * matrix is contiguous; gkb_filter_major() tells the caller which layout a handle uses. */
int gkb_create_lti(int kind, int n, int m, int c, int64_t n_filters, int device,
                   const double* x0, int x0_per_filter, const double* P0,
                   const double* F, const double* G, const double* H, const double* Q, const double* R,
                   gkb_filter** out);
/* NewInformationFromState (information.go:65-81): I0 = inv(P0) (zeros if singular), i0 = I0 x0. */
int gkb_create_information_from_state(int n, int m, int c, int64_t n_filters, int device,
                                      const double* x0, const double* P0, const double* F, const double* G,
                                      const double* H, const double* Q, const double* R, gkb_filter** out);
/* NewHybridKF (hybrid.go:23-34).  Q is q x q (used as Gamma Q Gamma^T), may be NULL when q = 0. */
int gkb_create_hybrid(int n, int m, int q, int64_t n_filters, int device, const double* x0,
                      int x0_per_filter, const double* P0, const double* Q, const double* R, gkb_filter** out);
/* NewSRIF (srif.go:14-49).  P0 is assumed diagonal, as in the reference. */
int gkb_create_srif(int n, int m, int64_t n_filters, int device, const double* x0, int x0_per_filter,
                    const double* P0, const double* R, int non_tri_r, gkb_filter** out);
void gkb_destroy(gkb_filter* f);

/* ---- LDKF setters (kalman.go:41-46).  Same quirks as the reference: SetInputControl does not
 *      re-evaluate needCtrl; on GKB_INFORMATION SetNoise does NOT refresh inv(Q)/inv(R)
 *      (information.go:136-138); on GKB_SQRT it re-factors chol(Q), chol(R) (squareroot.go:100-114). */
int gkb_set_state_transition(gkb_filter* f, const double* F);
int gkb_set_input_control(gkb_filter* f, int c, const double* G);
int gkb_set_measurement_matrix(gkb_filter* f, int m, const double* H);
int gkb_set_noise(gkb_filter* f, const double* Q, int m_r, const double* R);
/* Replay noise: Noise.Process(k)/Measurement(k) return the uploaded vectors, indexed by the
 * filter's step counter like noise.go BatchNoise (73-86) but keeping Q and R.  w is
 * [steps][n][n_filters], v is [steps][m][n_filters] (either may be NULL = zeros), `mem` says
 * where they live; they are copied.  Cleared by gkb_set_noise. */
int gkb_set_replay_noise(gkb_filter* f, int steps, const double* w, const double* v, int mem);
/* AWGN noise (noise.go:109-159) on an ordinary LDKF handle: from now on Noise.Process(k) / Noise.Measurement(k) are
 * drawn on the device from Philox4x32-10 keyed by (seed, filter_offset + filter, step k) and coloured with chol(Q),
 * chol(R) of the handle's current noise matrices -- normals [0, n) for the first Process(k) call of a step,
 * [n, n + m) for Measurement(k), [n + m, 2n + m) for Vanilla.Update's second Process(k) call (vanilla.go:195; the
 * reference's AWGN also draws afresh there).  It is the stream gkb_mc_chisquare uses for (trial, step): a pure
 * predictor with this noise reproduces NewMonteCarloRuns' truth sample for sample.  GKB_ERR_ARG at the next
 * gkb_update when Q or R is not positive definite (NewAWGN panics, noise.go:149-156).  Cleared by gkb_set_noise
 * and gkb_set_replay_noise. */
int gkb_set_philox_noise(gkb_filter* f, uint64_t seed, int64_t filter_offset);
/* The same samples on the host (AWGN.Process(k) / AWGN.Measurement(k) of the Go interface, noise.go:127-137), for ONE
 * (filter, step): w [n] first Process draw, v [m] Measurement draw, w2 [n] second Process draw; any may be NULL. */
int gkb_awgn_sample(int n, int m, const double* Q, const double* R, uint64_t seed, int64_t filter, int step, int device,
                    double* w, double* v, double* w2);
/* Reset() (vanilla.go:121-125): state <- initial estimate, step <- 0. */
int gkb_reset(gkb_filter* f);
/* Stream the handle's kernels and copies are enqueued on (a cudaStream_t; NULL = the legacy
 * default stream, which is the default). */
int gkb_set_stream(gkb_filter* f, void* stream);
/* Reference-order ("strict") arithmetic for a GKB_HYBRID handle: every product of hybrid.go:114-182 as a full
 * dense product in the written order, no fused multiply-adds (Go on amd64 never fuses), IEEE divisions, the
 * dense Joseph form and AsSymDense (GKB_ERR_ASYMMETRIC can be raised in this mode) -- the arithmetic the
 * reference itself executes (the `0 +` that starts each of gonum's sums is elided: exact up to the sign of a zero),
 * at about 0.4 of the production kernels' speed (5.1e9 against 1.2-1.4e10 updates/s at n = 6, m = 2 on one B200).
 * WHEN TO USE IT: whenever the results have to be the reference's.  The fast kernels restructure the Joseph update
 * and use FMAs; that is invisible (<= 1e-10) on well-conditioned runs, but the conventional covariance form of an
 * orbit-determination run (statOD: R = 1e-6 against P0 = 10, cond(P) ~ 1e13) amplifies ANY change of rounding to
 * percent level after a few dozen epochs -- the reference's own formulas move that much when a C compiler merely
 * contracts a*b+c (bench.py: production_vs_strict, cpu_baseline.fma_spread).  bench.py's headline runs in this mode.
 * On a GKB_SRIF handle it selects the literal epoch of srif.go:101-160 (the general
 * kernel: x-bar = Phi inv(R) b, b-bar = R-bar x-bar formed explicitly, full mat64.Inverse tests) instead of the
 * production epoch, which takes b-bar = b and differs from it at rounding level (1.5e-13 on the full-size run).
 * on = 0 selects the production kernels.  DEFAULT: on = 1 for a GKB_HYBRID handle created with n_filters = 1 (the
 * reference-shaped, one-filter use: its estimates are the reference's bit for bit), on = 0 for batched handles and SRIF. */
int gkb_set_strict(gkb_filter* f, int on);
int64_t gkb_n_filters(const gkb_filter* f);
/* 1 when the handle's per-filter arrays are filter-major [N][C] (large-state handles), 0 for SoA [C][N]. */
int gkb_filter_major(const gkb_filter* f);
int gkb_step(const gkb_filter* f);

/* ---- outputs of an update call: the fields of Estimate (kalman.go:64-72), any pointer may be
 *      NULL.  With every_step = 0 only the estimate after the last step is written ([C][N]);
 *      with every_step = 1 all steps are ([steps][C][N]).  `mem` = GKB_HOST or GKB_DEVICE for all
 *      pointers in the struct. */
typedef struct gkb_outputs {
  int mem;
  int every_step;
  double* state;      /* Estimate.State()           [n]                                        */
  double* meas;       /* Estimate.Measurement()     [m]                                        */
  double* innov;      /* Estimate.Innovation()      [m]  (information / SRIF: [n], see below)   */
  double* covar;      /* Estimate.Covariance()      [n*n]                                      */
  double* pred_covar; /* Estimate.PredCovariance()  [n*n]                                      */
  double* gain;       /* Gain()                     [n*m]                                      */
  double* obs_dev;    /* ObservationDev() (hybrid, SRIF) [m]                                   */
  int32_t* status;    /* [n_filters]: 0 or the first gkb_status the filter hit                  */
} gkb_outputs;
/* A filter whose Update / Predict fails at some step (the reference returns (nil, err): singular S, singular Phi,
 * a non-finite estimate) keeps its PREVIOUS estimate, records the first error in status[filter], and the rows of
 * the failed (filter, step) in every requested output array are filled with NaN -- never left holding data of an
 * earlier call.  (Large-state handles: the failing filter stops at that step and is restored to its state at the
 * start of the call; its output rows from the failing step on are not written.)
 * innov has n components for GKB_INFORMATION (returns i+, information.go:272-274) and GKB_SRIF
 * (returns b, srif.go:238-240). */

/* ---- LDKF.Update(measurement, control), `steps` times (vanilla.go:128-220,
 *      information.go:153-227, squareroot.go:129-274).  y is [steps][m][n_filters], or
 *      [steps][m] when y_shared != 0; u is [steps][c] (shared) or NULL.  `in_mem` says where y/u
 *      live.  The time loop runs inside one kernel launch with the state in registers. */
int gkb_update(gkb_filter* f, int steps, const double* y, int y_shared, const double* u, int in_mem,
               const gkb_outputs* out);

/* ---- NLDKF (kalman.go:51-60): one call runs `steps` epochs of
 *        Prepare(Phi, Htilde) [+ PreparePNT(Gamma)] + Update(real, computed)  or  Predict()
 *      (hybrid.go:78-204, srif.go:82-160).  flags[k] (shared by all filters) selects per epoch:
 *        GKB_F_MEAS  Update (else Predict);  GKB_F_EKF  EKF mode on (EnableEKF/DisableEKF);
 *        GKB_F_SNC   PreparePNT(Gamma[k]) was called.
 *      Phi [steps][n*n][N] (or [steps][n*n] if phi_shared), Htilde [steps][m*n][N] (or shared),
 *      real_obs/computed_obs [steps][m][N], Gamma [steps][n*q] shared or NULL. */
#define GKB_F_MEAS 1
#define GKB_F_EKF 2
#define GKB_F_SNC 4
int gkb_nl_run(gkb_filter* f, int steps, const uint8_t* flags, const double* Phi, int phi_shared,
               const double* Htilde, int h_shared, const double* real_obs, const double* computed_obs,
               const double* Gamma, int in_mem, const gkb_outputs* out);

/* ---- Orbit-determination inputs ON THE DEVICE: what the reference's callers compute with the `smd` propagator
 *      before every Prepare(Phi, Htilde) + Update(real, computed) (hybrid_test.go:159-294: reference orbit, its
 *      state-transition matrix, range / range-rate partials and computed observations), for a whole batch.
 *      Per filter: a reference orbit (ECI r [km], v [km/s]) propagated with two-body + J2 dynamics, ONE classical
 *      RK4 step of `dt` seconds per epoch on the state and on the variational equations (Phi = STM over the epoch).
 *      Per epoch, shared by the batch: the tracking station's ECI position / velocity at the END of the epoch and
 *      the truth's noise-free (range, range-rate) there; every filter's real observation is that plus
 *      sigma * N(0, 1) from Philox keyed by (seed, filter_offset + filter, epoch).  Outputs per (filter, epoch):
 *      Phi [36], Htilde [12] = d(range, range-rate)/d(r, v), real [2], computed [2] -- the inputs of gkb_nl_run. */
typedef struct gkb_od_config {
  double mu;              /* gravitational parameter, km^3/s^2                                            */
  double j2;              /* J2 zonal coefficient (0 = pure two-body)                                      */
  double re;              /* equatorial radius, km                                                         */
  double dt;              /* seconds per epoch                                                             */
  const double* orbit0;   /* [6][n_filters] initial reference orbits; gkb_od_run: NULL = continue          */
  int orbit_mem;          /* GKB_HOST or GKB_DEVICE (orbit0)                                               */
  const double* station;  /* host [steps][6]                                                               */
  const double* truth_obs;/* host [steps][2]                                                               */
  double sigma_range, sigma_rate;
  uint64_t seed;
  int64_t filter_offset;  /* global index of filter 0 (multi-GPU shards keep one noise stream per filter)  */
} gkb_od_config;
/* Writes the streams ([steps][C][n_filters], C = 36 / 12 / 2 / 2; `mem` = where the four arrays live) and, when
 * orbit_out != NULL ([6][n_filters], same `mem`), the reference orbits after the last epoch. */
int gkb_od_synthesize(const gkb_od_config* cfg, int steps, int64_t n_filters, int device, double* Phi, double* Htilde,
                      double* real_obs, double* computed_obs, int mem, double* orbit_out);
/* The fused run on a GKB_HYBRID handle with n = 6, m = 2: per epoch the synthesis above and then
 * Prepare + Update / Predict (flags as in gkb_nl_run, host array or NULL = Update every epoch; no SNC), in one
 * kernel -- Phi / Htilde / observations live in registers and never reach HBM.  Bit-identical to gkb_od_synthesize
 * followed by gkb_nl_run on its streams.  Outputs: state / covar (+ status), final or every step.  Honours
 * gkb_set_strict.  The handle keeps the advanced reference orbits: cfg->orbit0 = NULL continues from them.
 * Final-estimate outputs with out->mem = GKB_HOST: when out->state / out->covar are page-locked (cudaHostAlloc /
 * cudaHostRegister) the kernel writes them directly over PCIe while it runs; pageable buffers are staged and copied
 * after the kernel.  Either way the arrays are complete when the call returns. */
int gkb_od_run(gkb_filter* f, const gkb_od_config* cfg, int steps, const uint8_t* flags, const gkb_outputs* out);

/* ---- SmoothAll (hybrid.go:209-238, srif.go:165-192): backward sweep over the stored estimates of a
 *      run, in place.  For k = steps-2 .. 0:  S = inv(Phi[k+1]);  state[k] = S state[k+1];
 *      covar[k] = S covar[k+1] S^T -- each step reads the values the previous one wrote, like the
 *      reference's loop.  Phi [steps][n*n][N] (or [steps][n*n] shared), state [steps][n][N], covar
 *      [steps][n*n][N]; `mem` says where all of them live.  status [N] (optional) gets
 *      GKB_ERR_SINGULAR_PHI where an STM cannot be inverted (the reference returns an error there).
 *      Epochs that used SNC (Gamma) are "not yet implemented" in the reference (hybrid.go:234 panics):
 *      the caller must not pass such histories. */
int gkb_smooth_all(int n, int steps, int64_t n_filters, int device, const double* Phi, int phi_shared,
                   double* state, double* covar, int mem, int32_t* status);

/* ---- HouseholderTransf(A, n, m) (helper.go:142-172), the reference's exported helper and the core of the
 *      SRIF measurement update (srif.go:298-340), on `count` independent (n+m) x (n+1) matrices in place.
 *      A is [(n+m)*(n+1)][count] (row-major components, matrix index fastest; count = 1 is the plain
 *      row-major matrix).  Sign(|v| < 1e-12) = +1 as in helper.go:133-138. */
int gkb_householder_transf(int n, int m, int64_t count, int device, double* A, int mem);

/* ---- BatchKF (batch.go:34-79): `steps` SetNextMeasurement(realObs, computedObs, Phi, H) calls followed
 *      by Solve(), for N independent batch filters.  Per filter: Lambda = sum (H^T R) H,
 *      N = sum (H^T R)(real - computed) -- the reference multiplies by R, not inv(R) (batch.go:50), kept;
 *      P0 = inv(Lambda) (upper triangle mirrored, AsSymDense), xHat0 = P0 N.  Phi is stored but never
 *      used by the reference's accumulation, so it is not an argument.  R is m x m (host, upper triangle
 *      read); H [steps][m*n][N] (or [steps][m*n] shared), observations [steps][m][N]; outputs xhat0 [n][N],
 *      P0 [n*n][N]; `mem` says where H / observations / outputs live.  status [N] (optional):
 *      GKB_ERR_SINGULAR_S where Lambda cannot be inverted (Solve returns that error; outputs are zero). */
int gkb_batch_solve(int n, int m, int steps, int64_t n_filters, int device, const double* R, const double* H,
                    int h_shared, const double* real_obs, const double* computed_obs, int mem, double* xhat0,
                    double* P0, int32_t* status);

/* ---- VanLoan (c2d.go:13-75) for a batch of continuous-time systems: per system M = [[-A dt, Gamma W Gamma^T dt],
 *      [0, A^T dt]], E = expm(M) (Higham's scaling and squaring, Pade 3/5/7/9/13: the algorithm of mat64.Dense.Exp),
 *      F = (E_22)^T, Q = AsSymDense(F E_12).  A [n*n][count] (or [n*n] when a_shared), Gamma [n*q][count] (or shared),
 *      W [q*q] shared, dt [count] (or one value when dt_shared); outputs F, Q [n*n][count]; status [count] optional.
 *      `mem` says where all arrays live.  n, q <= 8.  The Nyquist warning of c2d.go:15-28 (eigenvalues of A) is the
 *      host wrapper's job: this call never fails on it, like the reference still returns F and Q. */
int gkb_van_loan(int n, int q, int64_t count, int device, const double* A, int a_shared, const double* Gamma, int g_shared,
                 const double* W, const double* dt, int dt_shared, int mem, double* F, double* Q, int32_t* status);

/* Raw filter state: x-like vector [n][N] and matrix [n*n][N] (vanilla/hybrid: x, P; information:
 * i, I; sqrt: x, S; SRIF: b, R).  Host pointers. */
int gkb_get_state(const gkb_filter* f, double* vec, double* mat);
int gkb_set_state(gkb_filter* f, const double* vec, const double* mat);

/* ---- Monte Carlo + chi-square: NewMonteCarloRuns (montecarlo.go:92-119) fused with
 *      NewChiSquare (chisquare.go:16-95).  Per trial: a pure-predictor Vanilla with AWGN noise
 *      generates the truth (state, measurement) and a tested filter of `kind` consumes the
 *      measurements; NEES and NIS are reduced over trials per step.  Nothing is stored unless an
 *      optional dump pointer is given. */
typedef enum gkb_noise_mode {
  GKB_NOISE_PHILOX = 0, /* Philox4x32-10, key = seed, counter = (trial, step, block); inverse normal CDF */
  GKB_NOISE_REPLAY = 1  /* uploaded, already coloured w [steps][n][trials], v [steps][m][trials]  */
} gkb_noise_mode;

typedef struct gkb_mc_config {
  int kind;               /* tested filter: GKB_VANILLA, GKB_INFORMATION (from state) or GKB_SQRT */
  int n, m, c;
  const double *F, *G, *H, *Q, *R; /* host; the TRUTH generator's model (and the tested filter's too
                                      unless filter_* below are given)                            */
  const double* x0_truth; /* [n]                                                                */
  const double* x0_filter;/* [n]                                                                */
  const double* P0;       /* [n*n]                                                              */
  int64_t trials;         /* trials run by THIS call (this GPU's shard)                         */
  int64_t trial_offset;   /* global index of the first trial: Philox is keyed by global trial   */
  int steps;
  const double* controls; /* [steps][c] host, or NULL = zero controls (montecarlo.go:98-104)    */
  int noise_mode;
  uint64_t seed;
  const double *w, *v;    /* GKB_NOISE_REPLAY only; location given by noise_mem                 */
  int noise_mem;
  int with_nees, with_nis;
  int info_raw_init;      /* GKB_INFORMATION only: x0_filter / P0 already are (i0, I0) as given to
                             NewInformation, instead of (x0, P0) as given to NewInformationFromState */
  int device;
  /* The tested filter's OWN model -- NewChiSquare takes any LDKF (chisquare.go:16,39) with its own F / G / H and
   * its own Noise (Q, R), e.g. a deliberately mis-tuned Q against a fixed truth.  Each pointer: host, same
   * shape as the truth generator's matrix above, or NULL = identical to it.  F, G, H, Q, R above are the pure
   * predictor's (montecarlo.go:92) and colour the truth's AWGN; NIS uses the tested H and R (chisquare.go:64-66). */
  const double *filter_F, *filter_G, *filter_H, *filter_Q, *filter_R;
} gkb_mc_config;

typedef struct gkb_mc_outputs {
  int mem;              /* location of every pointer below                                        */
  int sums_only;        /* 1: write per-step SUMS over this call's trials (for a multi-GPU
                           all-reduce); 0: write means = sum / trials (chisquare.go:85-92)        */
  double* nis;          /* [steps]                                                               */
  double* nees;         /* [steps]                                                               */
  /* optional, host only: per-step statistics of the truth states over this call's trials for
   * MonteCarloRuns.Mean / StdDev (montecarlo.go:18-59), always raw sums.  d = x - x_ref where
   * x_ref is the noise-free trajectory:  mean = x_ref + sum_d / N,
   * var = (sum_dd - sum_d^2 / N) / (N - 1).  Each is [steps][n]. */
  double* sum_d;
  double* sum_dd;
  double* x_ref;
  double* truth_x;      /* optional dump [steps][n][trials]                                      */
  double* truth_y;      /* optional dump [steps][m][trials]                                      */
  double* noise_w;      /* optional dump of the coloured noise actually used [steps][n][trials]  */
  double* noise_v;      /* optional dump [steps][m][trials]                                      */
  int32_t* status;      /* optional [trials]: per-trial first error                               */
  int32_t* first_error; /* optional, ONE word: 0, or the status of a failed trial (the reference
                           panics when the tested filter's Update fails, chisquare.go:40-42)      */
} gkb_mc_outputs;

int gkb_mc_chisquare(const gkb_mc_config* cfg, const gkb_mc_outputs* out);

/* ---- the same run sharded over several GPUs of THIS process, collective included: device
 *      devices[i] runs the contiguous trial range i of cfg->trials (TOTAL trials; Philox is keyed by the global trial
 *      index cfg->trial_offset + trial, so the trajectories do not depend on the device count), then the per-step
 *      NEES / NIS sums -- and the Mean / StdDev sums when out->sum_d / sum_dd are given -- are reduced:
 *        GKB_REDUCE_NCCL  one ncclAllReduce (SUM, double, 2 x steps [+ 2 n steps]) over a communicator set created with
 *                         ncclCommInitAll and cached per device list (libnccl.so.2 is bound with dlopen at first use);
 *        GKB_REDUCE_PEER  one kernel on devices[0] that adds the peers' sums through NVLink peer-memory loads in
 *                         shard order -- a fixed, rank-ordered sum: bit-reproducible for a given device list, and the
 *                         only mode that accepts the same device twice.
 *      cfg->device is ignored; outputs are host arrays (nis / nees means or sums, sum_d / sum_dd / x_ref, first_error);
 *      per-trial dumps and replay noise are single-device features.  n_devices = 1 is gkb_mc_chisquare. */
typedef enum gkb_reduce { GKB_REDUCE_NCCL = 0, GKB_REDUCE_PEER = 1 } gkb_reduce;
int gkb_mc_chisquare_multi(const gkb_mc_config* cfg, const int* devices, int n_devices, int reduce, const gkb_mc_outputs* out);

/* Device time (ms, CUDA events on the launch stream) of the kernels launched by the last
 * gkb_update / gkb_nl_run / gkb_mc_chisquare call on this thread, and how many kernels that was. */
float gkb_last_kernel_ms(void);
/* Same, for the dominant kernel alone (the fused Monte Carlo kernel without its memset/finish). */
float gkb_last_main_kernel_ms(void);
int gkb_last_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif
