/* gko_linalg.h -- dense FP64 helpers for the CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * The reference (ChristopherRabotin/gokalman) does all arithmetic through gonum
 * (github.com/gonum/matrix/mat64, un-vendored, un-pinned: `go get` HEAD circa Dec 2016 - 2017,
 * .travis.yml:2-7).  gonum's mat64 is a row-major pure-Go restatement of reference BLAS/LAPACK, so
 * this file restates the *published* unblocked LAPACK algorithms the gonum calls resolve to for
 * n <= 64 (block size 64 is never reached):
 *   Dense.Mul / MulVec      -> dgemm / dgemv, naive sequential inner sum          (gko_mul*)
 *   Dense.Inverse           -> dgetf2 (partial pivot) + dtrti2 + dgetri, then a
 *                              condition test against 1e16                        (gko_inverse)
 *   Cholesky.Factorize      -> dpotf2                                             (gko_chol_lower)
 *   QR.Factorize + RFromQR  -> dgeqr2 / dlarfg / dlarf, R_kk = -sign(a_kk)|a_k:|   (gko_qr_r)
 * All matrices are row-major double arrays with explicit dimensions.
 */
#ifndef GKO_LINALG_H
#define GKO_LINALG_H

#ifdef __cplusplus
extern "C" {
#endif

/* C[r x c] = A[r x k] * B[k x c].  C must not alias A or B. */
void gko_mul(double* C, const double* A, const double* B, int r, int k, int c);
/* C[r x c] = A[r x k] * B^T, B is [c x k]. */
void gko_mul_nt(double* C, const double* A, const double* B, int r, int k, int c);
/* C[r x c] = A^T * B, A is [k x r], B is [k x c]. */
void gko_mul_tn(double* C, const double* A, const double* B, int r, int k, int c);
/* y[r] = A[r x c] * x[c] */
void gko_mulvec(double* y, const double* A, const double* x, int r, int c);
/* y[c] = A^T x, A is [r x c] */
void gko_mulvec_t(double* y, const double* A, const double* x, int r, int c);
void gko_transpose(double* At, const double* A, int r, int c);

/* Explicit inverse by LU with partial pivoting (mat64.Dense.Inverse).
 * Returns 0 on success, 1 if exactly singular (zero pivot: output undefined, filled with what
 * the factorisation left), 2 if the condition number exceeds 1e16 (output IS the computed
 * inverse, as gonum still fills the receiver).  *cond_out (may be NULL) receives
 * ||A||_inf * ||A^-1||_inf (gonum uses dgecon's estimate of the same quantity; unverified here). */
int gko_inverse(double* Ainv, const double* A, int n, double* cond_out);

/* Lower Cholesky factor L (A = L L^T) from the upper triangle of A (mat64.Cholesky.Factorize +
 * TriDense.LFromCholesky).  Returns 1 if A is positive definite, 0 otherwise (the reference
 * ignores this flag everywhere). */
int gko_chol_lower(double* L, const double* A, int n);

/* R factor ([rows x cols], zeros below the diagonal) of the Householder QR of A[rows x cols]
 * (mat64.QR.Factorize + Dense.RFromQR), LAPACK dgeqr2 sign convention. */
void gko_qr_r(double* R, const double* A, int rows, int cols);

/* helper.go:142-172 HouseholderTransf: in-place on A[(n+m) x (n+1)]. */
void gko_householder_transf(double* A, int n, int m);

/* helper.go:65-84 AsSymDense: returns 0 if every off-diagonal pair agrees within 1e-6 abs OR
 * 1e-2 rel (floats.EqualWithinAbsOrRel), else 1.  On success mirrors the upper triangle into the
 * lower one in place (mat64.SymDense only ever reads the upper triangle). */
int gko_as_sym(double* A, int n);

/* helper.go:133-138 */
double gko_sign(double v);

#ifdef __cplusplus
}
#endif
#endif
