// Model-specific kernels around the factorisation engine: parameter transforms, covariance construction,
// gradient contraction, priors.  See models.cu.
#pragma once
#include "common.cuh"
#include "engine.cuh"

namespace nmgp {

// Hyper-parameter constants precomputed on the host once per plan.
struct HyperConst {
  // GP priors (nonseparable: [tilde_l, uL]; separable: [tilde_l, tilde_sigma])
  double mu0, mu1;
  // torch.distributions.Normal(loc, scale) built from Python numbers rounds loc/scale to float32 and evaluates
  // scale^2 and log(scale) in float32 (Utility/logpos.py:283,446,450); these are those values, widened.
  double n_loc, n_var, n_logscale;   // Normal(0,c) on uL (stationary / separable)
  double s_loc, s_var, s_logscale;   // Normal(mu_tilde_l, sigma_tilde_l) on tilde_l (stationary)
  double ig_a, ig_b, ig_alogb, ig_lgamma;  // inverse gamma: a log b, lgamma(a)  (distributions.py:126-134)
  double half_log2pi;                // math.log(math.sqrt(2*math.pi)) as torch's Normal.log_prob forms it
  int prior;                         // `Prior=` keyword
};

// Per-chunk scratch (device pointers); sizes in comments are per chunk of `cs` subjects.
struct Scratch {
  double* ell;    // [cs][N]
  double* sig;    // [cs][N]            (separable / stationary)
  double* s2;     // [cs]
  double* Lst;    // nonseparable: [cs][n][MT] rows (i,m) of L_i, zero padded; separable: [cs][M*M] L
  double* Kx;     // [cs][N][N]  Gibbs kernel incl. jitter
  double* CK;     // [cs][N][N]  c_ij * K0_ij  (d K0_ij / d tilde_l_i)
  double* alpha;  // [cs][nmat][n]   Sigma^-1 y
  double* yv;     // [cs][nmat][n]   right-hand sides (nonseparable: Y itself is used instead)
  double* Wout;   // nonseparable [cs][n][MT];  separable: rowSK [cs][N][M]
  double* Vout;   // nonseparable [cs][n][MT];  separable: KA    [cs][N][M]
  double* Sa;     // nonseparable [cs][N][MT]  rank-one (alpha alpha^T) part of the gradient sums, Kx-weighted
  double* Ca;     // nonseparable [cs][N][MT]  same, CK-weighted
  double* Ua;     // nonseparable [cs][N][MT]  u_j = L_j^T alpha_j
  double* gl;     // separable [cs][N]   sum_j GK_ij CK_ij
  double* gs;     // separable [cs][N]   sum_j GK_ij K0_ij
  double* lam;    // separable [cs][M]
  double* Vec;    // separable [cs][M][M]  eigenvectors of B (columns)
  double* R0;     // [cs][N]       residual for prior 0
  double* R1;     // [cs][N][nv1]  residual for prior 1 (nonseparable nv1 = T; separable nv1 = 1)
  double* Z0;     // [cs][N]
  double* Z1;     // [cs][N][nv1]
  double* G0;     // [cs][N]
  double* G1;     // [cs][N][nv1]
};

// ---- stand-alone helpers of kronecker_operation.py / kernels.py / distributions.py (kron.cu; cold paths)
int launch_pairwise_sqdist(const double* x1, int N1, const double* x2, int N2, double* out, cudaStream_t st);
int launch_kron(const double* t1, int h1, int w1, const double* t2, int h2, int w2, double* out, cudaStream_t st);
int launch_kron_mv(const double* B, int M1, int M2, const double* K, int N1, int N2, const double* y, double* out,
                   double* scratch, cudaStream_t st);
int launch_gram(const double* L, int R, int M, double* out, cudaStream_t st);
int launch_sym_eig(const double* B, int M, double* lam, double* V, cudaStream_t st);
// sigma2 I + B (x) K through eig(B) + M Cholesky factorisations: out2 = {log det, r^T inverse r}, inv_out the dense inverse
int kron_eig_solve(const double* B, int M, const double* K, int N, double sigma2, double* inv_out, const double* r,
                   double* out2, int* info_out, cudaStream_t st);

#define NMGP_FOR_EACH_M(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)

int padded_M(int M);  // template bucket for the nonseparable contraction (>= M)

// ---- elementwise covariance kernels (unit entry points + plan creation)
int launch_rbf_cov(const double* x1, int N1, const double* x2, int N2, double alpha, double beta, double* out,
                   cudaStream_t st);
int launch_gibbs_cov(const double* x1, const double* s1, const double* l1, int N1, const double* x2, const double* s2,
                     const double* l2, int N2, double* out, cudaStream_t st);
// prior covariance alpha^2 exp(-0.5 d/beta^2) + jitter I for `cs` subjects, written into the engine's padded layout
int launch_prior_cov_blocks(const double* x, int cs, int N, double alpha, double beta, const BlockBatch& b,
                            cudaStream_t st, long* launches);
// after potrf: compact lower factor [cs][N][N] and half log-determinants
int launch_extract_factor(const BlockBatch& b, int N, double* Lp, double* hld, cudaStream_t st, long* launches);

// blocked substitution with the cached prior factor: trans=0 solves L X = rhs, trans=1 solves L^T X = rhs;
// rhs / out are [cs][N][nv]
int launch_prior_solve(const double* Lp, const double* rhs, double* out, int cs, int N, int nv, int trans,
                       cudaStream_t st, long* launches);

// ---- nonseparable model (Utility/logpos.py:326-380)
int svc_forward(int cs, int N, int M, const double* x, const double* pars, int P, const HyperConst& h, const Scratch& w,
                const BlockBatch& b, cudaStream_t st, long* launches);  // prep + Kx + build
int svc_backward(int cs, int N, int M, const double* Y, const double* pars, int P, const HyperConst& h,
                 const Scratch& w, const BlockBatch& b, const double* hld0, const double* hld1, double* vals,
                 double* grad, int* info, cudaStream_t st, long* launches);  // symv + contraction + finish
int launch_nonseparable_cov_reference_order(const double* x, const double* pars, int batch, int N, int M, double* out,
                                            cudaStream_t st);

// pieces of the nonseparable pass used on their own by the prediction path
int launch_svc_prep(int cs, int N, int M, const double* pars, int P, const HyperConst& h, const Scratch& w, cudaStream_t st,
                    long* launches);   // pars -> ell, Lst, s2, prior residuals R0 / R1
// Gibbs kernel (with jitter) and its tilde_l log-derivative for `cs` subjects: Kx, CK [cs][N][N]; sig may be NULL (= ones)
int launch_kx(const double* x, const double* ell, const double* sig, int cs, int N, double* Kx, double* CK, cudaStream_t st,
              long* launches);
int launch_reduce_info(const int* info_mat, int cs, int nmat, int* info_out, cudaStream_t st, long* launches);
int launch_symv(const BlockBatch& b, int n, const double* y, double* alpha, int batch, cudaStream_t st, long* launches);

// ---- posterior prediction, nonseparable model (Utility/prediction.py:1038-1262; predict.cu)
// conditional moments of one GP prior at G new inputs per subject; scratch: predict_prior_scratch_per_subject doubles each
size_t predict_prior_scratch_per_subject(int N, int G);
int launch_predict_prior(const double* x, const double* Lp, const double* Z, int cs, int N, int nv, const double* xstar,
                         int G, double alpha, double beta, double mu, double* scratch, double* mean_out, double* s2_out,
                         cudaStream_t st, long* launches);
// predictive mean / variance of the M outputs for G * ns (grid point, sample) columns per subject, after the chunk's
// covariance has been built, factored and inverted in `b` (A = Sigma^-1) with `w` from svc_forward
size_t predict_scratch_per_subject(int N, int M, long C);
int predict_moments_chunk(int cs, int N, int M, const double* x, const double* Y, const Scratch& w, const BlockBatch& b,
                          const double* xstar, const double* tl_star, const double* uL_star, int G, int ns,
                          int raw_factor, double* scratch, size_t scratch_doubles, double* mu_f, double* s2y,
                          cudaStream_t st, long* launches);

// separable / stationary models (prediction.py:34-460, 1566-1692): after sep_forward + factorisation + inverse of the chunk's
// M matrices per subject; tl_star / ts_star [cs][G*ns] are the (sampled or plugged-in) tilde_l* / tilde_sigma*; returns
// mu_f = k_f^T Sigma^-1 y and quad = diag(k_f^T Sigma^-1 k_f), both [cs][G*ns][M]
size_t predict_sep_scratch_per_subject(int N, int M, long C);
int predict_moments_sep_chunk(int cs, int N, int M, const double* x, const Scratch& w, const BlockBatch& b,
                              const double* xstar, const double* tl_star, const double* ts_star, int G, int ns,
                              double* scratch, size_t scratch_doubles, double* mu_f, double* quad, cudaStream_t st,
                              long* launches);

// ---- separable / stationary models (Utility/logpos.py:237-296, 405-462)
int sep_forward(int model, int cs, int N, int M, const double* x, const double* Y, const double* pars, int P,
                const HyperConst& h, const Scratch& w, const BlockBatch& b, cudaStream_t st, long* launches);
int sep_backward(int model, int cs, int N, int M, const double* pars, int P, const HyperConst& h, const Scratch& w,
                 const BlockBatch& b, const double* hld0, const double* hld1, double* vals, double* grad, int* info,
                 cudaStream_t st, long* launches);

// ---- irregularly sampled ("Hadamard") objectives (Utility/logpos.py:465-716; hadamard.cu).  variant: 0 separable,
// 1 SVC, 2 stationary.  One N x N matrix per subject; w.Lst holds the row factors [cs][N][16], w.Ua their transpose
// [cs][16][N], w.Wout / Vout / Sa / Ca the gradient tables [cs][N][16].
int had_forward(int variant, int cs, int N, int M, const double* x, const int* indx, const double* pars, int P,
                const HyperConst& h, const Scratch& w, const BlockBatch& b, cudaStream_t st, long* launches);
int had_backward(int variant, int cs, int N, int M, const double* y, const int* indx, const double* pars, int P,
                 const HyperConst& h, const Scratch& w, const BlockBatch& b, const double* hld0, const double* hld1,
                 double* vals, double* grad, int* info, cudaStream_t st, long* launches);

// ---- gradient with respect to the hyper-parameters of the priors (hyper.cu; the north star's tied-hyper-prior extension,
// no reference implementation: SURVEY 8e)
struct HyperRaw {
  double hy[9];            // the plan's hyper-parameter vector (include/nmgp_b200.h order), unrounded
  double a, b, digamma_a;  // inverse gamma on sigma2_err
  int prior;
};
int launch_sep_prep(int model, int cs, int N, int M, const double* Y, const double* pars, int P, const HyperConst& h,
                    const Scratch& w, cudaStream_t st, long* launches);
int prior_traces(const double* x, const double* Lp, int cs, int N, double alpha, double beta, double* scratch, double* trI,
                 double* trB, cudaStream_t st, long* launches);
int launch_sweep_reduce(const double* vals, const double* hgrad, const int* info, long S, double* out, cudaStream_t st);
int prior_quad_blocks(int N);   // partial sums per subject written by launch_prior_quad: out is [cs][blocks][4]
int launch_prior_quad(const double* x, const double* Z, const double* G, int cs, int N, int nv, double alpha, double beta,
                      double* out, cudaStream_t st, long* launches);
int launch_hyper_finish(int model, int cs, int N, int M, int P, const double* pars, const HyperRaw& h, const double* s2v,
                        const double* q0, const double* q1, const double* trI0, const double* trB0, const double* trI1,
                        const double* trB1, int nv1, double* hgrad, cudaStream_t st, long* launches);

}  // namespace nmgp
