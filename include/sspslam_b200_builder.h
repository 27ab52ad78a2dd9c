/* C-ABI of libssb_builder.so - the device side of the BUILDER, a separate shared library so that the
 * stepping library (libssb.so, include/sspslam_b200.h) does not depend on cuBLAS / cuSOLVER.
 *
 * Replaces the regularised least-squares decoder solves nengo's builder performs for every decoded connection
 * (`nengo.solvers.LstsqL2`, used by the reference at sspslam/networks/pathintegration.py:180-182, binding.py:316-317,
 * slam.py:298-303, associativememory.py:38-54: 509 solves per model), batched over the ensembles of one
 * shape (and over the models of trials that have their own network seed).  Plain pointers, float64 like nengo. */
#ifndef SSPSLAM_B200_BUILDER_H
#define SSPSLAM_B200_BUILDER_H
#ifdef __cplusplus
extern "C" {
#endif

/* For s < n_sys:  sigma = reg * max(A_s),  X_s = (A_s^T A_s + m sigma^2 I)^-1 A_s^T Y_s.
 * A [n_sys][m][n] activities (row-major), Y [n_sys][m][k] targets, X [n_sys][n][k] decoders (transposed: nengo's
 * decoders are X^T).  Needs m >= n (nengo's default evaluation-point count is >= 2 n).  Returns 0, or < 0 with
 * ssb_builder_last_error(). */
int ssb_solve_decoders(int device, int n_sys, int m, int n, int k, const double* A, const double* Y, double reg, double* X);
const char* ssb_builder_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SSPSLAM_B200_BUILDER_H */
