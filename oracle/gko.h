/* gko.h -- CPU oracle for the gokalman hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (gokalman_b200/) never links, imports or calls it.
 *
 * It is a line-by-line CPU restatement (plain C, FP64, no FMA contraction) of the reference's
 * filters, one function per reference function, all quirks included:
 *   vanilla.go:128-220     -> gko_update (GKO_VANILLA / GKO_PREDICTOR)
 *   information.go:153-227 -> gko_update (GKO_INFORMATION), 231-316 -> estimate accessors
 *   squareroot.go:129-274  -> gko_update (GKO_SQRT)
 *   hybrid.go:104-204      -> gko_nl_update / gko_nl_predict (GKO_HYBRID), 209-238 -> gko_hybrid_smooth
 *   srif.go:14-49,101-160,223-340 -> GKO_SRIF
 *   noise.go:23-106        -> noiseless / replay (BatchNoise-style index-by-k) noise
 *   montecarlo.go:92-119, chisquare.go:16-95 -> gko_mc_chisquare
 *
 * PARITY PINNING: vanilla/information/sqrt are pinned by the reference's golden CSVs
 * (examples/jerkcar/{vanilla,information,sqrt}.csv, 6 decimals); HouseholderTransf and the SRIF
 * measurement update by the reference's known-answer tests (helper_test.go:108-117,
 * srif_test.go:15-56), VanLoan by c2d_test.go:9-33.  HybridKF, SmoothAll, BatchKF, Monte Carlo and chi-square
 * numerics have NO runnable reference pin (their reference tests need the absent `smd` package / are time-seeded):
 * for those this oracle is "parity unpinned" by reference vectors and is pinned INDEPENDENTLY instead
 * (tests/test_oracle_crosscheck.py): hybrid CKF == the golden-pinned vanilla on an LTI model, the SmoothAll identity,
 * BatchKF against numpy's normal equations, and the Monte Carlo + chi-square means against the exact moments of an
 * analytic linear-Gaussian recursion (tests/chi2_moments.py).  The OD-input synthesis (gko_od.c) restates the engine's
 * own documented algorithm (the reference's callers use the external `smd` propagator): parity unpinned.
 * The reference itself (Go + gonum) cannot be built here: no Go toolchain, no gonum sources.
 */
#ifndef GKO_H
#define GKO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum gko_kind { GKO_VANILLA = 0, GKO_PREDICTOR = 1, GKO_INFORMATION = 2, GKO_SQRT = 3, GKO_HYBRID = 4, GKO_SRIF = 5 };

/* error codes returned by update calls (0 = OK).  "panic" marks where the reference panics. */
enum gko_err {
  GKO_OK = 0,
  GKO_ERR_DIMS = -1,
  GKO_ERR_SINGULAR_S = -2,   /* vanilla.go:164-167, hybrid.go:150-152 */
  GKO_ERR_ASYMMETRIC = -3,   /* vanilla.go:207-215, hybrid.go:184-192; information.go:214-222 panics */
  GKO_ERR_LOCKED = -4,       /* hybrid.go:105-107, srif.go:102-104 */
  GKO_ERR_SINGULAR_PHI = -5, /* srif.go:112-114 */
  GKO_ERR_SINGULAR_R = -6,   /* srif.go:228-230 panics */
  GKO_ERR_NOISE_RANGE = -7   /* noise.go:74-76,82-84 panics */
};

typedef struct gko_filter gko_filter;

/* One Estimate (kalman.go:64-72), flattened.  Every array is caller-provided storage inside the
 * struct; sizes follow the filter that produced it. */
#define GKO_MAXN 64
#define GKO_MAXM 16
typedef struct gko_estimate {
  int n, m;
  double state[GKO_MAXN];              /* Estimate.State()                                    */
  double meas[GKO_MAXM];               /* Estimate.Measurement()                              */
  double innov[GKO_MAXN];              /* Estimate.Innovation() (information/SRIF: n entries) */
  int innov_len;
  double covar[GKO_MAXN * GKO_MAXN];   /* Estimate.Covariance()                               */
  double pred_covar[GKO_MAXN * GKO_MAXN];
  double gain[GKO_MAXN * GKO_MAXM];    /* Gain() where the estimate type has one              */
  double obs_dev[GKO_MAXM];            /* hybrid/SRIF ObservationDev()                        */
  /* raw internal representation: information (i, I, I-), sqrt (S+, S-), SRIF (b, R, Rbar) */
  double raw_vec[GKO_MAXN];
  double raw_mat[GKO_MAXN * GKO_MAXN];
  double raw_pred_mat[GKO_MAXN * GKO_MAXN];
  int covar_ok, pred_covar_ok;         /* 0 when the lazy inverse failed (zeros returned)     */
} gko_estimate;

/* ---- LDKF constructors (kalman.go:35-48).  G may be NULL / all-zero => needCtrl=false. -------- */
gko_filter* gko_new_vanilla(int n, int m, int c, const double* x0, const double* P0, const double* F,
                            const double* G, const double* H, const double* Q, const double* R,
                            int pure_predictor);
gko_filter* gko_new_information(int n, int m, int c, const double* i0, const double* I0, const double* F,
                                const double* G, const double* H, const double* Q, const double* R);
gko_filter* gko_new_information_from_state(int n, int m, int c, const double* x0, const double* P0,
                                           const double* F, const double* G, const double* H,
                                           const double* Q, const double* R);
gko_filter* gko_new_sqrt(int n, int m, int c, const double* x0, const double* P0, const double* F,
                         const double* G, const double* H, const double* Q, const double* R);
/* ---- NLDKF constructors (kalman.go:51-60) ------------------------------------------------------ */
gko_filter* gko_new_hybrid(int n, int m, int q, const double* x0, const double* P0, const double* Q,
                           const double* R);
gko_filter* gko_new_srif(int n, int m, const double* x0, const double* P0, const double* R, int non_tri_r);
void gko_free(gko_filter* f);

/* setters (mirror SetStateTransition / SetInputControl / SetMeasurementMatrix / SetNoise) */
void gko_set_state_transition(gko_filter* f, const double* F);
void gko_set_input_control(gko_filter* f, int c, const double* G);
void gko_set_measurement_matrix(gko_filter* f, int m, const double* H);
/* Noiseless(Q,R): zero samples. m_r is the dimension of R. */
void gko_set_noise(gko_filter* f, const double* Q, int m_r, const double* R);
/* Replay noise: like noise.go BatchNoise (index by k) but carrying Q and R.  w is [steps][n],
 * v is [steps][m]; either may be NULL (zeros).  The arrays are copied. */
void gko_set_replay(gko_filter* f, int steps, const double* w, const double* v, int m_v);
/* AWGN semantics (noise.go:127-131: a fresh draw on every call): w2 [steps][n] is what the SECOND Process(k) call of
 * Vanilla.Update (vanilla.go:195) returns; NULL = BatchNoise semantics (the same vector k twice). */
void gko_set_replay_second_draw(gko_filter* f, const double* w2);
void gko_reset(gko_filter* f);
void gko_initial_estimate(const gko_filter* f, gko_estimate* est);

/* LDKF.Update(measurement, control).  u may be NULL when the filter has no control. */
int gko_update(gko_filter* f, const double* y, const double* u, gko_estimate* est);

/* NLDKF */
void gko_prepare(gko_filter* f, const double* Phi, const double* Htilde);
void gko_prepare_pnt(gko_filter* f, const double* Gamma);
void gko_enable_ekf(gko_filter* f, int on);
int gko_nl_predict(gko_filter* f, gko_estimate* est);
int gko_nl_update(gko_filter* f, const double* real_obs, const double* computed_obs, gko_estimate* est);
void gko_srif_set_non_tri_r(gko_filter* f, int non_tri_r);

/* srif.go:298-340 measurementSRIFUpdate: R[n x n], H[m x n], b[n], y[m] -> Rk, bk, ek */
void gko_measurement_srif_update(int n, int m, const double* R, const double* H, const double* b,
                                 const double* y, double* Rk, double* bk, double* ek);

/* hybrid.go:209-238 / srif.go:165-192 SmoothAll over a stored history: Phi[steps][n*n],
 * x[steps][n], P[steps][n*n] are overwritten in place for k = steps-2 .. 0. Returns 0 or
 * GKO_ERR_SINGULAR_PHI / GKO_ERR_ASYMMETRIC. */
int gko_smooth_all(int n, int steps, const double* Phi, double* x, double* P);

/* Batch runners (bench.py CPU baseline; batch-sized parity checks): one filter object per filter, OpenMP
 * over filters.  NLDKF: SoA streams [step][component][filter], flags bit 0 = Update (else Predict), bit 1 =
 * EKF; outputs [component][filter] of the last estimate.  Vanilla: y [steps][nf][m], outputs filter-major. */
int gko_run_nl_batch(int kind, int n, int m, int64_t nf, int steps, const uint8_t* flags, const double* x0,
                     const double* P0, const double* R, const double* Phi, const double* Htilde,
                     const double* real_obs, const double* computed_obs, int threads, double* out_state,
                     double* out_covar);
int gko_run_vanilla_batch(int n, int m, int64_t nf, int steps, const double* x0, const double* P0, const double* F,
                          const double* H, const double* Q, const double* R, const double* y, int threads,
                          double* out_state, double* out_covar);

/* batch.go:34-79 BatchKF: `count` SetNextMeasurement calls followed by Solve().  R[m x m],
 * H[count][m*n], real_obs / computed_obs [count][m] -> xhat0[n], P0[n*n].  Returns 0,
 * GKO_ERR_SINGULAR_S when Lambda cannot be inverted, GKO_ERR_ASYMMETRIC from AsSymDense. */
int gko_batch_solve(int n, int m, int count, const double* R, const double* H, const double* real_obs,
                    const double* computed_obs, double* xhat0, double* P0);

/* ---- Monte Carlo + chi-square (montecarlo.go:92-119 + chisquare.go:16-95) -------------------- */
typedef struct gko_mc_config {
  int n, m, c;
  int kind;                   /* tested filter: GKO_VANILLA, GKO_INFORMATION or GKO_SQRT      */
  const double *F, *G, *H, *Q, *R;
  const double* x0_truth;     /* initial state of the truth generator (pure predictor)        */
  const double* x0_filter;    /* initial state of the tested filter                           */
  const double* P0;
  int trials, steps;
  const double* controls;     /* [steps][c], or NULL => zero controls (montecarlo.go:98-104)   */
  /* noise source for the truth generator: replay arrays if non-NULL, else Philox4x32-10 keyed by
   * `seed` with counter (trial_offset+trial, step) -> inverse normal CDF (gko_icdf_normal), coloured by chol(Q), chol(R). */
  const double* w;            /* [trials][steps][n], already coloured                          */
  const double* v;            /* [trials][steps][m]                                            */
  uint64_t seed;
  int64_t trial_offset;
  int with_nees, with_nis;
  int threads;                /* OpenMP threads over trials; <= 1 = serial like the reference  */
  /* The tested filter's OWN model (chisquare.go:16 takes any LDKF: its F/G/H and its Noise's Q/R), each NULL =
   * the same matrix as the truth generator's above.  NIS uses the tested filter's H and R (chisquare.go:64-66). */
  const double *tF, *tG, *tH, *tQ, *tR;
} gko_mc_config;

/* Returns 0 or an error code.  nis_means / nees_means are [steps] (chisquare.go:94 returns them
 * in this order).  Optional: mean_state / std_state [steps][n] = MonteCarloRuns.Mean/StdDev
 * (montecarlo.go:18-59) of the truth states; truth_x [trials][steps][n] and truth_y
 * [trials][steps][m] receive the generated truth when non-NULL. */
int gko_mc_chisquare(const gko_mc_config* cfg, double* nis_means, double* nees_means,
                     double* mean_state, double* std_state, double* truth_x, double* truth_y);

/* Philox4x32-10 (Salmon et al., SC'11; Random123 v1.09), one block: ctr[4], key[2] -> out[4]. */
void gko_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* One standard normal from one 32-bit word: the engine's piecewise-quintic inverse CDF (same table as the kernels). */
double gko_icdf_normal(uint32_t k);
/* The oracle's standard-normal stream: normal #j (j = 0..n+m-1) of (seed, trial, step). */
void gko_philox_normals(uint64_t seed, uint64_t trial, uint32_t step, int count, double* z);

/* gko_od.c: the engine's OD-input synthesis restated as the generic RK4 on the 6 + 36 state / STM equations
 * (two-body + J2), range / range-rate partials and observations.  Streams are SoA [step][component][filter]. */
int gko_od_synth(double mu, double j2, double re, double dt, int64_t nf, int steps, const double* orbit0,
                 const double* station, const double* truth_obs, double sigma_range, double sigma_rate, uint64_t seed,
                 int64_t filter_offset, double* Phi, double* Ht, double* real_obs, double* comp_obs, double* orbit_out);

/* gko_c2d.c: c2d.go:13-75 VanLoan (without the Nyquist warning) and the Higham-2005 matrix exponential it rests on. */
int gko_expm(double* E, const double* A, int d);
int gko_van_loan(int n, int q, const double* A, const double* Gamma, const double* W, double dt, double* F, double* Q);

#ifdef __cplusplus
}
#endif
#endif
