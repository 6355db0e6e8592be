/* nmgp_b200.h -- C ABI of the B200-native NMGP log-posterior + gradient library (libnmgp_b200.so).
 *
 * The reference (Corleno/Nonstationary_Multivariate_Gaussian_Process) is pure Python and has no FFI;
 * its plugin surface is the module attribute `Utility.logpos.nlogpos_obj{_S,,_SVC}` called from the
 * drivers' MAP/HMC loops.  Each entry point below names the reference interface it stands in for.
 * Conventions: plain pointers and sizes, FP64 everywhere, no C++/torch types; every function returns
 * 0 on success or a negative NMGP_E* code and never throws; `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream); all work is stream-ordered on it; the caller owns every
 * buffer it passes, the plan owns its caches and workspace; a plan is not thread-safe.
 *
 * Layouts (row-major, contiguous):
 *   x      [S,N]      sorted time stamps per subject            (Utility/logpos.py: argument `x`)
 *   Y      [S,N,M]    observations                              (argument `Y`, N by M)
 *   pars   [S,P]      flat parameter vectors, reference order:
 *            stationary   [tilde_l, tilde_sigma, uL(T), tilde_sigma2_err]           P = T+3     (logpos.py:46-57)
 *            separable    [tilde_l(N), tilde_sigma(N), uL(T), tilde_sigma2_err]     P = 2N+T+1  (logpos.py:17-29)
 *            nonseparable [tilde_l(N), uL(N*T) time-major, tilde_sigma2_err]        P = N+N*T+1 (logpos.py:32-43)
 *          with T = M(M+1)/2 and uL a row-major lower triangle whose diagonal slots are log-scale
 *          (Utility/utils.py:10-22).
 *   vals   [S,NMGP_NVALS]  vals[0] = -logpost (what nlogpos_obj* returns), vals[1] = loglik, then the
 *          prior components in the order of the reference's verbose tuple, zero padded:
 *            stationary   [., ., lp_tilde_l, lp_uL, lp_sigma2, 0]                   (logpos.py:399-400)
 *            separable    [., ., lp_tilde_l, lp_tilde_sigma, lp_uL, lp_sigma2]      (logpos.py:231-232)
 *            nonseparable [., ., lp_tilde_l, lp_uL, lp_sigma2, 0]                   (logpos.py:313-320)
 *   grad   [S,P]      d(-logpost)/d pars  (what `NegLog.backward()` leaves in the leaves' .grad)
 *   info   [S]        0 = ok; k>0 = the Cholesky of that subject's covariance failed at pivot k
 *                     (value and gradient of that subject are NaN; other subjects are unaffected)
 */
#ifndef NMGP_B200_H
#define NMGP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMGP_STATIONARY 0   /* Utility/logpos.py:383-462  nlogpos_obj_S   */
#define NMGP_SEPARABLE 1    /* Utility/logpos.py:216-296  nlogpos_obj     */
#define NMGP_NONSEPARABLE 2 /* Utility/logpos.py:299-380  nlogpos_obj_SVC */

/* irregularly sampled data, one observation (x_n, indx_n, y_n) per row -- the "Hadamard" objectives */
#define NMGP_HADAMARD 3     /* Utility/logpos.py:465-558  nlogpos_obj_hadamard     (separable)  */
#define NMGP_HADAMARD_SVC 4 /* Utility/logpos.py:561-637  nlogpos_obj_hadamard_SVC (nonseparable) */
#define NMGP_HADAMARD_S 5   /* Utility/logpos.py:640-716  nlogpos_obj_hadamard_S   (stationary) */

#define NMGP_NVALS 6
#define NMGP_NHYPER 9

#define NMGP_OK 0
#define NMGP_EINVAL (-1)
#define NMGP_ECUDA (-2)
#define NMGP_ENOMEM (-3)

/* hyper[NMGP_NHYPER], host memory, keyword arguments of the objective in signature order:
 *   stationary   {mu_tilde_l, sigma_tilde_l, a, b, c}                                            (logpos.py:383)
 *   separable    {mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma,
 *                 beta_tilde_sigma, a, b, c}                                                     (logpos.py:216)
 *   nonseparable {mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, a, b}          (logpos.py:299)
 */

typedef struct nmgp_plan nmgp_plan;

/* Number of parameters P of one subject for (model, N, M); negative on bad arguments.
 * Mirrors the slicing of vec2pars_S / vec2pars / vec2pars_SVC (logpos.py:17-57). */
int nmgp_n_params(int model, int N, int M);

/* Create the evaluation plan for S independent subjects of equal shape (N time points, M outputs).
 * Caches everything that is loop-invariant in the drivers' MAP/HMC loops: copies of x and Y, the
 * GP-prior covariances of logpos.py:271-281 / :357-365 factored once (the reference re-factors them
 * 1+T times per call), and the workspace.  x_dev / Y_dev are device pointers.  prior_flag mirrors the
 * `Prior=` keyword.  workspace_limit_bytes bounds the factorisation workspace (0 = choose from free
 * memory); subjects are processed in chunks that fit. */
int nmgp_plan_create(nmgp_plan** out, int model, int S, int N, int M, const double* x_dev, const double* Y_dev,
                     const double* hyper, int prior_flag, size_t workspace_limit_bytes, void* stream);

/* Plan for the Hadamard objectives (model = NMGP_HADAMARD / _SVC / _S): S subjects with N observations each,
 * x [S,N] inputs, indx [S,N] (int32) the output index 0..M-1 of every observation, y [S,N] the observations, all device
 * pointers.  Parameter vectors, hyper-parameters and value tuples follow the regularly sampled model of the same family
 * (pars: logpos.py:479 / :60-72 / :657; the cross-output factor entries are RAW, no exp on the diagonal: logpos.py:518,
 * 582-583, 682).  The plan is evaluated with nmgp_logpost_grad / _host like any other; the covariance is one dense
 * N x N matrix per subject, K_x[n,n'] <r_n, r_n'> + sigma2_err I with r_n the row indx_n of the factor at observation n. */
int nmgp_plan_create_hadamard(nmgp_plan** out, int model, int S, int N, int M, const double* x_dev, const int* indx_dev,
                              const double* y_dev, const double* hyper, int prior_flag, size_t workspace_limit_bytes,
                              void* stream);

int nmgp_plan_destroy(nmgp_plan* plan);

/* One batched evaluation: for every subject the value tuple and (if grad_dev != NULL) the gradient.
 * Device pointers; stream-ordered; no host synchronisation.  Stands in for
 * `nlogpos_obj*(pars, Y, x, **hyper, verbose=True)` followed by `.backward()`
 * (e.g. Nonseparable_Model/Nonseparable_model_mpisim.py:183-186) over S subjects at once. */
int nmgp_logpost_grad(nmgp_plan* plan, const double* pars_dev, double* vals_dev, double* grad_dev, int* info_dev,
                      void* stream);

/* Gradient of -log posterior of every subject with respect to the plan's hyper-parameters, hgrad_dev [S][NMGP_NHYPER] in
 * the order documented above (unused slots 0).  In the reference every subject carries its own fixed hyper-parameter
 * dictionary (Nonseparable_model_mpisim.py:311-312); this is the quantity a caller that TIES the hyper-priors across
 * subjects needs: sum it over the local subjects and all-reduce it across ranks (sharding.py).  No reference
 * implementation exists; validated against finite differences of the objective.  Uses the cached prior factors and
 * needs no likelihood factorisation.  Zero when the plan was created with prior_flag = 0. */
int nmgp_hyper_grad(nmgp_plan* plan, const double* pars_dev, double* hgrad_dev, void* stream);

/* nmgp_logpost_grad and nmgp_hyper_grad in one pass: the hyper-parameter gradient reuses the evaluation's own prior solves
 * (three extra small launches per chunk).  This is the sweep of a sharded run with tied hyper-priors: vals / hgrad are
 * summed over the rank's subjects and all-reduced in one 17-double vector (sharding.py). */
int nmgp_logpost_grad_hyper(nmgp_plan* plan, const double* pars_dev, double* vals_dev, double* grad_dev, double* hgrad_dev,
                            int* info_dev, void* stream);

/* One rank's contribution to the sweep's single all-reduce, in one launch with a fixed summation order:
 * out17[0..5] = sum over the successful subjects (info == 0) of vals[s][0..5], out17[6] = number of failed subjects,
 * out17[7] = S, out17[8..16] = sum over the successful subjects of hgrad[s][0..8] (zeros if hgrad_dev == NULL).
 * Device pointers.  The reference has no counterpart (its MPI ranks exchange nothing, Nonseparable_model_mpisim.py:39-43). */
int nmgp_sweep_reduce(const double* vals_dev, const double* hgrad_dev, const int* info_dev, long S, double* out17_dev,
                      void* stream);

/* Replace the plan's hyper-parameters (hyper[NMGP_NHYPER], host memory, same order as at creation) without rebuilding the
 * plan: the step a tied-hyper-prior optimiser takes between sweeps.  The GP-prior covariances whose (alpha, beta) changed
 * are factored again on `stream` (the reference re-factors them in every call, logpos.py:271-281 / :357-365); x, Y and the
 * workspace stay.  Must not run concurrently with an evaluation of the same plan. */
int nmgp_plan_set_hyper(nmgp_plan* plan, const double* hyper, void* stream);

/* CUDA-graph replay of launch-bound evaluations.  Single-chunk plans whose evaluation is at most 256 small kernels (the one-subject-per-process MAP / HMC loops of Stationary_model.py:106-131, Separable_model.py:149-231,
 * Nonseparable_model_mpisim.py:163-207) capture the kernel sequence of nmgp_logpost_grad / nmgp_logpost_grad_host once per
 * buffer tuple and replay it with one cudaGraphLaunch; same kernels, same results bit for bit.  mode 0 = automatic (default),
 * 1 = never (A/B timing, debugging).  nmgp_plan_graph_replays counts the evaluations served by a replay. */
int nmgp_plan_set_graph(nmgp_plan* plan, int mode);
long nmgp_plan_graph_replays(const nmgp_plan* plan);

/* Same call with HOST buffers (pageable or pinned): copies pars host->device, evaluates, copies
 * vals / grad / info back and synchronises the stream.  This is the call a reference maintainer binds
 *  and what bench.py times as `e2e`. */
int nmgp_logpost_grad_host(nmgp_plan* plan, const double* pars_host, double* vals_host, double* grad_host,
                           int* info_host, void* stream);

/* Same evaluation with CUDA events between its phases (synchronises once per chunk): phase_ms[NMGP_NPHASES] (host)
 * receives the milliseconds spent in {0: parameter transform + covariance build, 1: Cholesky factorisation (potrf),
 * 2: inverse (trtri + lauum + symmetrise), 3: GP-prior triangular solves, 4: alpha = Sigma^-1 y, gradient contraction
 * and assembly}.  bench.py uses it for the live roofline figures ("FP64 Cholesky TFLOP/s"). */
#define NMGP_NPHASES 5
int nmgp_logpost_grad_profile(nmgp_plan* plan, const double* pars_dev, double* vals_dev, double* grad_dev,
                              int* info_dev, float* phase_ms, void* stream);

/* Factorisation engine selection: 0 = automatic (left-looking accumulate-in-registers path for large batches,
 * right-looking tile tasks otherwise), 1 = force right-looking, 2 = force left-looking, 3 = left-looking with the
 * W^T W inverse also where the Takahashi sweep would be used (<= 16 block columns), 4 = automatic potrf with the
 * level-synchronous recursive triangular inverse + W^T W (the path for a few large matrices).  For tests and A/B timing. */
int nmgp_plan_set_engine(nmgp_plan* plan, int mode);

/* Number of kernels the plan's last evaluation launched (bench.py's `gpu_launches`). */
long nmgp_plan_last_launches(const nmgp_plan* plan);
/* Bytes of device memory owned by the plan. */
size_t nmgp_plan_device_bytes(const nmgp_plan* plan);
/* Subjects per chunk and the factorisation block size the plan chose. */
int nmgp_plan_chunk(const nmgp_plan* plan);
int nmgp_plan_block(const nmgp_plan* plan);

/* Device-resident optimiser step for the drivers' MAP loops (`optimizer = torch.optim.Adam(...)`; `optimizer.step()`
 * after `NegLog.backward()`: Stationary_model.py:112-126, Separable_model.py:158-166, Nonseparable_model_mpisim.py:177-190):
 * one Adam update (torch.optim.Adam semantics: no weight decay, no amsgrad) of all S parameter vectors in place, so that
 * `pars` and `grad` never leave the GPU between evaluations.  m, v: first / second moment estimates [S,P] (zero before
 * step 1); step: 1-based iteration count; frozen: optional [P] bytes, non-zero = parameter is not optimised (the
 * stationary drivers hold tilde_sigma fixed, Stationary_model.py:88,116); info: optional [S], subjects with
 * info != 0 (failed factorisation, NaN gradient) are left untouched.  All pointers are device pointers. */
int nmgp_adam_step(double* pars_dev, const double* grad_dev, double* m_dev, double* v_dev, const int* info_dev,
                   const unsigned char* frozen_dev, long S, long P, double lr, double beta1, double beta2, double eps,
                   long step, void* stream);

/* Device-resident pieces of Hamiltonian Monte Carlo for all S subjects at once.  The drivers' HMC loops
 * (Separable_model.py:209-210, Stationary_model_mpiKAISER.py:205-206, Nonseparable_model_mpiKAISER.py:267-270) hand
 * `potential_func = logpos.nlogpos_obj*` to an external sampler (HMC_Sampler, not part of the reference repository); what a
 * batched sampler needs from this library besides nmgp_logpost_grad is below, so that positions, momenta and gradients never
 * leave the GPU (LogPosteriorPlan.hmc_sample composes them).  All pointers are device pointers, [S,P] unless noted.
 *   nmgp_hmc_kick    p -= step * grad          for subjects with info[s] == 0 (info may be NULL)
 *   nmgp_hmc_drift   q += eps * p
 *   nmgp_hmc_accept  per subject: dH = (U[s] + |p0|^2/2) - (vals_prop[s,0] + |p1|^2/2); the proposal is accepted iff
 *                    failed[s] == 0 (may be NULL), dH is finite and log_u[s] < dH; then q <- q_prop, grad <- grad_prop,
 *                    U[s] <- vals_prop[s,0].  vals_prop is the [S,NMGP_NVALS] array of nmgp_logpost_grad; accepted [S] out. */
int nmgp_hmc_kick(double* p_dev, const double* grad_dev, const int* info_dev, long S, long P, double step, void* stream);
int nmgp_hmc_drift(double* q_dev, const double* p_dev, long S, long P, double eps, void* stream);
int nmgp_hmc_accept(double* q_dev, const double* q_prop_dev, double* grad_dev, const double* grad_prop_dev, double* U_dev,
                    const double* vals_prop_dev, const double* p0_dev, const double* p1_dev, const int* failed_dev,
                    const double* log_u_dev, int* accepted_dev, long S, long P, void* stream);

/* ---------------------------------------------------------------- posterior prediction (nonseparable model) ----
 * Stand-ins for the device work inside `point_/pointwise_/test_predmap_inhomogeneous_sampling`
 * (Utility/prediction.py:1038-1262; callers Nonseparable_Model/Nonseparable_model.py:377,387,399).  The reference draws,
 * per new input x* and per sample: tilde_l* ~ N(mu_l, sigma2_l), uL* ~ N(mu_uL, sigma2_uL) (the GP priors conditioned on the
 * MAP values), then y ~ N(mu_f, sigma2_y) given (tilde_l*, uL*).  The draws themselves stay on the host (torch's generator,
 * so the reference's random stream is reproduced, see prediction.py of this package); the two calls below produce every
 * quantity the draws need.  All pointers are device pointers; S subjects of the plan, G new inputs per subject.
 *
 * nmgp_predict_prior_moments  (prediction.py:1060-1092): xstar [S,G] ->
 *     mu_l [S,G], s2_l [S,G]   conditional mean / variance of tilde_l at x*   (variance < 0 -> 1e-6, prediction.py:1066)
 *     mu_uL [S,G,T], s2_uL [S,G]   same for the T columns of uL (one shared variance, prediction.py:1079-1081)
 *   computed with the plan's cached Cholesky factors of the prior covariances (the reference LU-solves per call).
 *
 * nmgp_predict_moments  (prediction.py:1130-1165): tl_star [S,G,n_sample], uL_star [S,G,n_sample,T] ->
 *     (uL_star is the unconstrained row-major triangle, exp() on its diagonal slots as utils.py:10-22; with
 *      flags & NMGP_PRED_RAW_FACTOR it is used as the factor itself, which is what the posterior-sample variant
 *      point_predsample_inhomogeneous does, prediction.py:1310-1311)
 *     mu_f [S,G,n_sample,M]   k_f^T Sigma^-1 y
 *     s2_y [S,G,n_sample,M]   diag(A - k_f^T Sigma^-1 k_f) + sigma2_err   (values <= 0 -> 1e-6, prediction.py:1163)
 *     info [S] (may be NULL)  as in nmgp_logpost_grad
 *   Sigma is built, factored and inverted ONCE per subject by the hot-path engine (the reference repeats an n x n
 *   `symeig` + `cholesky` per x* and per sample); see csrc/predict.cu for the sample-dependent part. */
/* The separable model's predictors (prediction.py:34-460) condition tilde_l and tilde_sigma the same way:
 * nmgp_predict_prior_moments on a separable plan returns, in the same four buffers, the moments of tilde_l ([S,G], [S,G]) and
 * of tilde_sigma ([S,G,1], [S,G]).
 *
 * nmgp_predict_moments_sep  (separable: prediction.py:82-118, 226-266, 372-398; stationary: prediction.py:1587-1596):
 *     tl_star, ts_star [S,G,n_sample]   tilde_l* and tilde_sigma* at every (new input, sample); the stationary model and the
 *                                       plug-in predictors pass their fixed values
 *     mu_f [S,G,n_sample,M]             k_f^T Sigma^-1 y with k_f = B (x) k_x
 *     quad [S,G,n_sample,M]             diag(k_f^T Sigma^-1 k_f); the caller forms sigma2_y = a2 - quad + sigma2_err with the
 *                                       variant's own a2 (prediction.py:103 vs :254) and clipping rule
 *   Sigma = B (x) K_x + sigma2 I is never formed: B = V diag(lam) V^T gives M independent N x N problems
 *   S_m = lam_m K_x + sigma2 I, factored and inverted once per subject by the hot-path engine. */
int nmgp_predict_moments_sep(nmgp_plan* plan, const double* pars_dev, const double* xstar_dev, int G, int n_sample,
                             const double* tl_star_dev, const double* ts_star_dev, double* mu_f_dev, double* quad_dev,
                             int* info_dev, void* stream);
int nmgp_predict_prior_moments(nmgp_plan* plan, const double* pars_dev, const double* xstar_dev, int G, double* mu_l_dev,
                               double* s2_l_dev, double* mu_uL_dev, double* s2_uL_dev, void* stream);
#define NMGP_PRED_RAW_FACTOR 1
int nmgp_predict_moments(nmgp_plan* plan, const double* pars_dev, const double* xstar_dev, int G, int n_sample,
                         const double* tl_star_dev, const double* uL_star_dev, int flags, double* mu_f_dev,
                         double* s2_y_dev, int* info_dev, void* stream);

/* Last error message of the calling thread ("" if none). */
const char* nmgp_last_error(void);

/* ---------------------------------------------------------------- unit entry points (tests, ncu) ----
 * Elementwise covariance kernels with the reference's semantics, device pointers:
 *   nmgp_rbf_cov       Utility/kernels.py:24-43   alpha^2 exp(-0.5 |(x1_i - x2_j)/beta|^2)   (+1e-6 I if x2 == NULL)
 *   nmgp_gibbs_cov     Utility/kernels.py:46-73   sigma1_i sigma2_j sqrt(2 l1_i l2_j/(l1_i^2+l2_j^2)) exp(-d/(..)) (+1e-6 I)
 * out is [N1,N2] row-major.  sigma pointers may be NULL (= ones). */
int nmgp_rbf_cov(const double* x1, int N1, const double* x2, int N2, double alpha, double beta, double* out,
                 void* stream);
int nmgp_gibbs_cov(const double* x1, const double* sigma1, const double* ell1, int N1, const double* x2,
                   const double* sigma2, const double* ell2, int N2, double* out, void* stream);

/* Dense nonseparable covariance K + sigma2_err I of logpos.py:339-352 for `batch` subjects, written in the
 * REFERENCE's output-major ordering (row (m,i) -> m*N+i), full symmetric [batch, NM, NM].
 * (The plan itself builds the same matrix time-major and blocked; this entry exists for parity tests.) */
int nmgp_nonseparable_cov(const double* x, const double* pars, int batch, int N, int M, double* out, void* stream);

/* Live FP64 tensor-pipe (DMMA.8x8x4) peak of the current device: register-only mma.sync.m8n8k4.f64 for at least
 * `min_seconds` (<= 10), CUDA-event timed on `stream`.  bench.py's roofline denominator (no reference counterpart). */
int nmgp_fp64_dmma_probe(double min_seconds, double* tflops_out, double* seconds_out, void* stream);

/* ---- stand-alone helpers behind the reference-signature mirrors (cold paths; all device pointers, FP64, row-major) ----
 * nmgp_pairwise_sqdist: out[i,j] = (x1_i^2 + x2_j^2) - 2 x1_i x2_j for N x 1 inputs, x2 NULL = x1
 *                       (Utility/kernels.py:5-21 `pairwise_distances`, same operation order).
 * nmgp_kron:            out [h1 h2, w1 w2] = t1 [h1,w1] (x) t2 [h2,w2]
 *                       (Utility/kronecker_operation.py:5-22 `kronecker_product`; :25-33 `kronecker_product_diag` is w1 = w2 = 1).
 * nmgp_kron_mv:         out [M1 N1] = (B [M1,M2] (x) K [N1,N2]) y [M2 N2] as vec((K Y) B^T), y output-major; scratch [M2 N1]
 *                       (Utility/kronecker_operation.py:72-85 `kron_mv`).
 * nmgp_gram:            out [R,R] = L L^T for the stacked per-time-point factors L [R,M]
 *                       (Utility/logpos.py:111-118 `generate_K_index_SVC`).
 * nmgp_sym_eig:         eigenvalues (ascending) and eigenvectors (columns of V [M,M]) of a symmetric M x M matrix, M <= 16,
 *                       one-warp cyclic Jacobi (what `torch.symeig(B, eigenvectors=True)` returns at
 *                       Utility/kronecker_operation.py:45, distributions.py:37).
 * nmgp_kron_eig_solve:  for Sigma = sigma2 I + B (x) K (B [M,M], K [N,N] symmetric, M <= 16): out2[0] = log det Sigma,
 *                       out2[1] = r^T Sigma^-1 r (0 if r is NULL), inv_out (or NULL) = Sigma^-1 dense [MN,MN], info (or NULL) =
 *                       0 or the first failing Cholesky pivot.  Block formulation: eig(B) + M Cholesky factorisations of
 *                       lam_m K + sigma2 I on the batched engine -- no N x N eigensolver.  Stands in for
 *                       `kron_inv` / `kron_logdet` (Utility/kronecker_operation.py:36-69) and
 *                       `multivariate_normal_logpdf0/1/2` (Utility/distributions.py:26-113).  Synchronises `stream`. */
int nmgp_pairwise_sqdist(const double* x1, int N1, const double* x2, int N2, double* out, void* stream);
int nmgp_kron(const double* t1, int h1, int w1, const double* t2, int h2, int w2, double* out, void* stream);
int nmgp_kron_mv(const double* B, int M1, int M2, const double* K, int N1, int N2, const double* y, double* out,
                 double* scratch, void* stream);
int nmgp_gram(const double* L, int R, int M, double* out, void* stream);
int nmgp_sym_eig(const double* B, int M, double* lam, double* V, void* stream);
int nmgp_kron_eig_solve(const double* B, int M, const double* K, int N, double sigma2, double* inv_out, const double* r,
                        double* out2, int* info, void* stream);

/* Batched blocked Cholesky engine on `batch` symmetric positive-definite matrices [batch, n, n] (row-major,
 * lower triangle referenced).  In place:
 *   potrf: lower triangle <- L (strict upper untouched), logdet[b] = log det A_b, info[b] as above.
 *   potri: after potrf, full matrix <- inverse (both triangles).
 * These wrap the same tile kernels the plan uses (copy into the padded block layout, run, copy out). */
int nmgp_potrf_batched(double* A, int n, int batch, double* logdet, int* info, void* stream);
int nmgp_potrf_potri_batched(double* A, int n, int batch, double* logdet, int* info, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NMGP_B200_H */
