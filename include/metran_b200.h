/*
 * metran_b200.h -- C ABI of libmetran_b200.so: batched sequential-processing Kalman
 * filter, innovation log-likelihood and RTS smoother for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the Kalman hot path of pastas/metran.  Every entry
 * point replaces one reference interface (file:line relative to the reference tree):
 *
 *   mkf_filter_history   <- seqkalmanfilter(...)            metran/kalmanfilter.py:236-400
 *                            (the 9-in / 7-out `filtermethod` plug-in point selected by
 *                             SPKalmanFilter.__init__/run_filter, :494-504, :729-771)
 *   mkf_filter_loglik    <- run_filter() + get_mle(warmup)   metran/kalmanfilter.py:696-778, 550-567
 *                            (what Metran.get_mle / BaseSolver.objfunction consume,
 *                             metran/metran.py:605-622, metran/solver.py:42-63)
 *   mkf_dfm_loglik       <- same, for Metran-shaped models   metran/metran.py:265-416
 *                            (A, Q diagonal; H = [I | loadings]; quirk Q13 of SURVEY.md)
 *   mkf_dfm_filter_history <- seqkalmanfilter for Metran-shaped batches (thread per model)
 *   mkf_dfm_loglik_alpha <- Metran._get_matrices + the above metran/metran.py:246-416, 605-622
 *                            (parameter vector alpha -> phi, q on the device)
 *   mkf_smoother         <- kalmansmoother(...)              metran/kalmanfilter.py:403-476
 *   mkf_project          <- SPKalmanFilter.simulate()        metran/kalmanfilter.py:569-603
 *   mkf_reduce_sum       <- the pooled objective handed to NCCL all-reduce (new; the
 *                            reference has no multi-device path)
 *
 * Conventions
 *   - All data pointers are DEVICE pointers to IEEE fp64 (double) unless noted; layouts
 *     are model-major and C-contiguous: y[B][T][m], A[B][n][n], H[B][m][n], fc[B][T][n][n].
 *   - A `*_stride` argument is the distance IN ELEMENTS between consecutive models;
 *     0 shares one array between all models of the batch (e.g. one observation matrix
 *     for 8192 parameter draws).
 *   - Observations are NaN-coded: a slot is "missing" when it is not finite or equals
 *     -1e10 exactly, the rule of SPKalmanFilter.set_observations (kalmanfilter.py:657,
 *     666-667).  Observed series of a time step are processed in ascending index order,
 *     the order set_observations produces (kalmanfilter.py:667-674).
 *   - Calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream) and return an int code; nothing throws across the boundary.
 *     Inputs are borrowed for the duration of the enqueued work, outputs are caller-owned.
 *   - Numerical failure follows the reference: NaN/inf propagate into the outputs
 *     (kalmanfilter.py:378 takes log of whatever f is); additionally status[b] gets
 *     MKF_STATUS_* bits so a caller can tell.
 *   - There is no CPU fallback: without a CUDA device every compute entry point
 *     returns MKF_ERR_CUDA.
 */
#ifndef METRAN_B200_H
#define METRAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MKF_VERSION 100 /* 0.1.0 */

/* return codes */
#define MKF_OK 0
#define MKF_ERR_BAD_ARG 1     /* NULL pointer, non-positive size, bad stride */
#define MKF_ERR_UNSUPPORTED 2 /* shape outside the compiled kernels (see mkf_limits) */
#define MKF_ERR_CUDA 3        /* CUDA runtime error; see mkf_last_cuda_error() */

/* per-model status bits (int32) */
#define MKF_STATUS_NONPOSITIVE_F 1 /* an innovation variance f <= 0 or non-finite */
#define MKF_STATUS_CHOL_FAILED 2   /* smoother: a predicted covariance was not positive definite; the
                                      singular directions got zero gain (the reference's pinv does the same) */

/* flags */
#define MKF_FLAG_A_DIAGONAL 1 /* caller guarantees A is diagonal (Metran: always)  */
#define MKF_FLAG_Q_DIAGONAL 2 /* caller guarantees Q is diagonal (Metran: always)  */

int mkf_version(void);
const char *mkf_error_string(int code);
const char *mkf_last_cuda_error(void);
/* n_max: largest state dimension of the generic kernels; dfm_m_max / dfm_k_max: largest
 * series / common-factor counts of the thread-per-model DFM kernels. */
void mkf_limits(int32_t *n_max, int32_t *dfm_m_max, int32_t *dfm_k_max);

/*
 * Forward filter with the full reference result contract (kalmanfilter.py:392-400).
 *   y[B][T][m]; A,Q[B][n][n]; H[B][m][n]; R[B][m]; x0[B][n] (NULL -> 0); P0[B][n][n]
 *   (NULL -> I), the defaults of run_filter (kalmanfilter.py:747-750).
 * The covariance predict uses diag(A) only, the mean predict the full A
 * (kalmanfilter.py:318-331).
 * Outputs (any history pointer may be NULL to skip it):
 *   sigma[B][T], detf[B][T]: per-step sum v^2/f and sum log f, 0 on steps without
 *     observation; cnt[B][T] (int32): observed-series count.  The reference's compacted
 *     `sigmas[:sigmacount]` is sigma[cnt > 0] (kalmanfilter.py:380-382).
 *   fm[B][T][n], fc[B][T][n][n], pm, pc: filtered / predicted means and covariances.
 *   mle[B], nobs[B] (int64), status[B] (int32): -2 log L with `warmup` semantics of
 *     get_mle (kalmanfilter.py:563-566): the first `warmup` OBSERVED steps are dropped
 *     from the sums, nobs counts observations of raw time steps t >= warmup.
 */
int mkf_filter_history(int64_t B, int64_t T, int32_t n, int32_t m,
                       const double *y, int64_t y_stride,
                       const double *A, int64_t A_stride,
                       const double *Q, int64_t Q_stride,
                       const double *H, int64_t H_stride,
                       const double *R, int64_t R_stride,
                       const double *x0, int64_t x0_stride,
                       const double *P0, int64_t P0_stride,
                       int32_t warmup, uint32_t flags,
                       double *sigma, double *detf, int32_t *cnt,
                       double *fm, double *fc, double *pm, double *pc,
                       double *mle, int64_t *nobs, int32_t *status, void *stream);

/* Same recursion, -2 log L only (no history traffic). */
int mkf_filter_loglik(int64_t B, int64_t T, int32_t n, int32_t m,
                      const double *y, int64_t y_stride,
                      const double *A, int64_t A_stride,
                      const double *Q, int64_t Q_stride,
                      const double *H, int64_t H_stride,
                      const double *R, int64_t R_stride,
                      const double *x0, int64_t x0_stride,
                      const double *P0, int64_t P0_stride,
                      int32_t warmup, uint32_t flags,
                      double *mle, int64_t *nobs, int32_t *status, void *stream);

/*
 * Metran-shaped dynamic factor model, one thread per model: n = m + k states,
 * A = diag(phi[n]), Q = diag(q[n]), H = [I_m | gamma[m][k]], R = r[m] (NULL -> 0),
 * x0 = 0, P0 = I (metran.py:283-384, kalmanfilter.py:747-750).
 */
int mkf_dfm_loglik(int64_t B, int64_t T, int32_t m, int32_t k,
                   const double *y, int64_t y_stride,
                   const double *phi, int64_t phi_stride,
                   const double *q, int64_t q_stride,
                   const double *gamma, int64_t gamma_stride,
                   const double *r, int64_t r_stride,
                   int32_t warmup,
                   double *mle, int64_t *nobs, int32_t *status, void *stream);

/*
 * mkf_filter_history for Metran-shaped models, one thread per model (same arithmetic as
 * mkf_dfm_loglik): the warp transposes its 32 per-step records through shared memory so the
 * history leaves in contiguous 256-byte pieces; bound by HBM writes (8 [2(n + n^2) + 2.5]
 * bytes per time step and model).  Output layout and semantics as mkf_filter_history.
 */
int mkf_dfm_filter_history(int64_t B, int64_t T, int32_t m, int32_t k,
                           const double *y, int64_t y_stride,
                           const double *phi, int64_t phi_stride,
                           const double *q, int64_t q_stride,
                           const double *gamma, int64_t gamma_stride,
                           const double *r, int64_t r_stride,
                           int32_t warmup,
                           double *sigma, double *detf, int32_t *cnt,
                           double *fm, double *fc, double *pm, double *pc,
                           double *mle, int64_t *nobs, int32_t *status, void *stream);

/*
 * As mkf_dfm_loglik with the parameter -> matrix map of Metran fused in front:
 * phi_i = exp(-dt / alpha_i) (metran.py:262-263); q_i = (1 - phi_i^2)(1 - comm_i) for
 * the m specific factors, comm_i = sum_j gamma[i][j]^2, and 1 - phi_i^2 for the k common
 * factors (metran.py:310-322).  alpha[B][m+k] in state order (series, then common).
 */
int mkf_dfm_loglik_alpha(int64_t B, int64_t T, int32_t m, int32_t k,
                         const double *y, int64_t y_stride,
                         const double *alpha, int64_t alpha_stride, double dt,
                         const double *gamma, int64_t gamma_stride,
                         int32_t warmup,
                         double *mle, int64_t *nobs, int32_t *status, void *stream);

/*
 * -2 log L on a COMPACTED time grid.  Metran forces a daily grid
 * (metran.py:571), so irregularly sampled series are mostly steps without any observation (94.5 % of
 * the reference's example data), for which the filter only predicts (kalmanfilter.py:335 not taken).
 * yc[B][J][m] holds the steps with at least one observation, tidx[B][J] (int32, ascending) their
 * original time indices; the k = tidx[j] - tidx[j-1] predicts between two of them are applied in
 * closed form (A, Q diagonal): phi^k and q (1 - phi^2k) / (1 - phi^2).  Models with fewer than J
 * observed steps pad with NaN rows and repeat their last index (k = 0: a no-op).  warmup / nobs
 * semantics as mkf_filter_history, with the ORIGINAL time index.  Either alpha (with dt) or phi
 * and q are given (the other pointers NULL); r may be NULL.  Strides of 0 share an array.
 * Same result as mkf_dfm_loglik on the full grid up to the rounding of the closed form
 * (<= 1e-12 relative on -2 log L; tests/test_gpu_compact.py).
 */
int mkf_dfm_loglik_compact(int64_t B, int64_t J, int32_t m, int32_t k,
                           const double *yc, int64_t yc_stride,
                           const int32_t *tidx, int64_t tidx_stride,
                           const double *alpha, int64_t alpha_stride, double dt,
                           const double *phi, int64_t phi_stride,
                           const double *q, int64_t q_stride,
                           const double *gamma, int64_t gamma_stride,
                           const double *r, int64_t r_stride,
                           int32_t warmup,
                           double *mle, int64_t *nobs, int32_t *status, void *stream);

/*
 * RTS smoother (kalmanfilter.py:453-474): G_t = P_f[t] A^T P_p[t+1]^-1 (solved by
 * Cholesky; the reference uses pinv), x_s[t] = x_f[t] + G_t (x_s[t+1] - x_p[t+1]),
 * P_s[t] = P_f[t] + G_t (P_s[t+1] - P_p[t+1]) G_t^T.  sm[B][T][n], sc[B][T][n][n].
 */
int mkf_smoother(int64_t B, int64_t T, int32_t n,
                 const double *fm, const double *fc, const double *pm, const double *pc,
                 const double *A, int64_t A_stride, uint32_t flags,
                 double *sm, double *sc, int32_t *status, void *stream);

/*
 * mkf_smoother with the projections of SPKalmanFilter.simulate (kalmanfilter.py:596-602) fused into
 * its epilogue: while the smoothed covariance of a step is still on chip,
 *   proj_mean[B][T][p] = Hp x_s[t],  proj_var[B][T][p] = max(diag(Hp P_s[t] Hp^T), 0),
 *   state_var[B][T][n] = diag(P_s[t])  (may be NULL)
 * are written, so no consumer has to read sc[B][T][n][n] back; sc itself may be NULL (then the
 * smoothed covariances are not materialised at all).  Hp[B][p][n], stride 0 = shared.
 */
int mkf_smoother_project(int64_t B, int64_t T, int32_t n,
                         const double *fm, const double *fc, const double *pm, const double *pc,
                         const double *A, int64_t A_stride, uint32_t flags,
                         int32_t p, const double *Hp, int64_t Hp_stride,
                         double *sm, double *sc, double *proj_mean, double *proj_var, double *state_var,
                         int32_t *status, void *stream);

/*
 * mkf_smoother runs batches of at least `min_batch` diagonal-A models with n = 4 or n = 8 on the
 * thread-per-model kernel (TMA-staged inputs and outputs; needs 16-byte aligned arrays), everything else
 * on the lane-per-row / four-warp kernels.  mkf_smoother_set_thread_min_batch sets that threshold (default
 * 6144; negative = default) and returns the previous one; mkf_dfm_set_hist_tma(0) disables the kernel
 * altogether.  Same results to rounding (used by the parity tests).  No reference counterpart.
 */
int64_t mkf_smoother_set_thread_min_batch(int64_t min_batch);

/*
 * Projection of state moments (kalmanfilter.py:596-602): means_out[B][T][p] = Hp x_t,
 * var_out[B][T][p] = max(diag(Hp P_t Hp^T), 0); Hp[B][p][n].
 */
int mkf_project(int64_t B, int64_t T, int32_t n, int32_t p,
                const double *means, const double *covs,
                const double *Hp, int64_t Hp_stride,
                double *means_out, double *var_out, void *stream);

/*
 * mkf_dfm_loglik / mkf_dfm_loglik_alpha run long, fully observed series (R = 0) as three launches:
 * transient kernel -> steady-state kernel -> resume kernel, exchanging one record per model through
 * a library-owned device scratch buffer (grow-only, one per device and stream; no reference
 * counterpart).  mkf_dfm_set_handoff(0) forces the single-kernel path (same results, bit for bit;
 * used by the parity tests); returns the previous setting.  mkf_release_workspaces() frees the
 * scratch buffers (call with the device idle).
 */
int mkf_dfm_set_handoff(int enabled);
void mkf_release_workspaces(void);
/*
 * mkf_dfm_filter_history writes the state arrays of n = 4 and n = 8 models with TMA tensor stores
 * (one cp.async.bulk.tensor per array and step for a warp's 32 models) when the arrays are 16-byte
 * aligned; mkf_dfm_set_hist_tma(0) forces the shared-memory transposition used for every other shape
 * (same values; used by the parity tests).  Returns the previous setting.  No reference counterpart.
 */
int mkf_dfm_set_hist_tma(int enabled);
/*
 * Measurement aid for the three-launch path (no reference counterpart).  With
 * mkf_dfm_set_profile(1) the library records CUDA events around each launch; afterwards
 * mkf_dfm_last_profile(stream, ms, stats) waits for `stream` and returns for the most recent
 * three-launch call on it: ms[3] = device time of the transient, steady-state and resume kernels
 * (-1 when not recorded); stats[3] = {warps, warps handed to the steady-state kernel, warp-steps done
 * after the hand-off}.  Either output pointer may be NULL.  MKF_ERR_BAD_ARG when no such call exists.
 */
int mkf_dfm_set_profile(int enabled);
int mkf_dfm_last_profile(void *stream, float *ms, int64_t *stats);

/* Deterministic (fixed-tree) sum of v[0..count) into out[0]; feeds the all-reduce. */
int mkf_reduce_sum(const double *v, int64_t count, double *out, void *stream);

/*
 * Measurement aid (no reference counterpart): saturates the fp64 FMA pipe with
 * blocks x 256 threads x iters x 64 independent DFMAs so bench.py can time the
 * denominator of the fp64 roofline on the same device and clocks.
 */
int mkf_probe_fp64(int32_t blocks, int32_t iters, double *scratch, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* METRAN_B200_H */
