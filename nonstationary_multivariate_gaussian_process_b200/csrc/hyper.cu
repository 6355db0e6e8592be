// Gradient of -log posterior with respect to the HYPER-parameters of the priors, per subject.
//
// In the reference every subject carries its own fixed hyper-parameter dictionary (Nonseparable_model_mpisim.py:311-312);
// BASELINE.json's north star ties them across the subjects of a sharded run ("many independent subjects sharing
// hyperpriors ... NCCL only to all-reduce the shared-hyperparameter gradient and log-posterior scalars").  The per-subject
// gradients computed here are summed over the local subjects and all-reduced by sharding.py; there is no reference
// implementation of this quantity, so it is validated against finite differences of the oracle's objective
// (tests/test_gpu_hyper_grad.py).
//
// GP prior  v_t ~ N(mu 1, Sigma_p),  Sigma_p = alpha^2 E + jitter I,  E_ij = exp(-0.5 (x_i - x_j)^2 / beta^2),  t = 1..nv:
//   lp = sum_t -0.5 [N log 2pi + log det Sigma_p + r_t^T Sigma_p^-1 r_t],   r_t = v_t - mu 1,   g_t = Sigma_p^-1 r_t
//   d lp / d mu    = sum_t 1^T g_t
//   d lp / d alpha = sum_t [-0.5 tr(Sigma_p^-1 dS_a) + 0.5 g_t^T dS_a g_t],   dS_a = 2 alpha E = (2/alpha)(Sigma_p - jitter I)
//                  = (1/alpha) sum_t [-(N - jitter tr Sigma_p^-1) + (r_t^T g_t - jitter g_t^T g_t)]
//   d lp / d beta  = sum_t [-0.5 tr(Sigma_p^-1 dS_b) + 0.5 g_t^T dS_b g_t],   dS_b = (alpha^2 / beta^3) E o D,  D_ij = (x_i - x_j)^2
// tr Sigma_p^-1 and tr(Sigma_p^-1 dS_b) depend only on (x, alpha, beta): formed once per plan from W = L_p^-1
// (tr Sigma_p^-1 = |W|_F^2, tr(Sigma_p^-1 B) = sum_ij (W B)_ij W_ij).
#include "models.cuh"

namespace nmgp {

namespace {

// identity right-hand sides [cs][N][N]
__global__ void eye_kernel(double* __restrict__ I, int N, long total) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long r = idx % ((long)N * N);
  I[idx] = (r / N == r % N) ? 1.0 : 0.0;
}

// Bd[c][i][j] = (alpha^2 / beta^3) E_ij D_ij
__global__ void dbeta_cov_kernel(const double* __restrict__ x, int N, double alpha2, double beta, double* __restrict__ Bd) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y, c = blockIdx.z;
  if (j >= N) return;
  const double d = x[(long)c * N + i] - x[(long)c * N + j];
  const double d2 = d * d;
  Bd[((long)c * N + i) * N + j] = alpha2 / (beta * beta * beta) * exp(-0.5 * d2 / (beta * beta)) * d2;
}

// per row i of W = L^-1 (lower triangular, [cs][N][N]):  pI[c][i] = sum_j W_ij^2,  pB[c][i] = sum_j (sum_k W_ik Bd_kj) W_ij
__global__ void __launch_bounds__(128) trace_rows_kernel(const double* __restrict__ W, const double* __restrict__ Bd, int N,
                                                         double* __restrict__ pI, double* __restrict__ pB) {
  __shared__ double scratch[40];
  __shared__ double wrow[2048];
  const int i = blockIdx.x, c = blockIdx.y;
  const double* Wi = W + ((long)c * N + i) * N;
  const double* B = Bd + (long)c * N * N;
  double sI = 0.0, sB = 0.0;
  for (int k0 = 0; k0 <= i; k0 += 2048) {             // W_ik, k <= i, staged through shared memory in slices
    const int kn = min(2048, i + 1 - k0);
    __syncthreads();
    for (int k = threadIdx.x; k < kn; k += blockDim.x) wrow[k] = Wi[k0 + k];
    __syncthreads();
    for (int j = threadIdx.x; j <= i; j += blockDim.x) {
      double y = 0.0;
      for (int k = 0; k < kn; ++k) y += wrow[k] * B[(long)(k0 + k) * N + j];
      sB += y * Wi[j];
    }
  }
  for (int j = threadIdx.x; j <= i; j += blockDim.x) sI += Wi[j] * Wi[j];
  const double tI = block_sum(sI, scratch), tB = block_sum(sB, scratch);
  if (threadIdx.x == 0) { pI[(long)c * N + i] = tI; pB[(long)c * N + i] = tB; }
}

__global__ void __launch_bounds__(256) sum_rows_kernel(const double* __restrict__ pI, const double* __restrict__ pB, int N,
                                                       double* __restrict__ trI, double* __restrict__ trB) {
  __shared__ double scratch[40];
  const int c = blockIdx.x;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) { a += pI[(long)c * N + i]; b += pB[(long)c * N + i]; }
  const double ta = block_sum(a, scratch), tb = block_sum(b, scratch);
  if (threadIdx.x == 0) { trI[c] = ta; trB[c] = tb; }
}

// per subject and prior, partial sums over a block of QRB rows i (one CTA each; summed in block order by hyper_finish_kernel,
// so the result does not depend on scheduling):
//   sg = sum_{i,t} g_it,  gg = sum g_it^2,  rg = sum z_it^2 (= r^T g),  gBg = sum_i sum_j dS_b,ij sum_t g_it g_jt
// The QRB rows of G sit in shared memory; every thread walks columns j and reads G_j once for all QRB rows.
constexpr int QRB = 8;
// A CTA walks `rbpc` consecutive row blocks (many subjects: one CTA per subject, one reduction at the end; a single large
// subject: one row block per CTA, so that it still spreads over the GPU).
__global__ void __launch_bounds__(128, 6) prior_quad_kernel(const double* __restrict__ x, const double* __restrict__ Z,
                                                            const double* __restrict__ G, int N, int nv, double alpha2,
                                                            double beta, int rbpc, double* __restrict__ part /*[cs][parts][4]*/) {
  extern __shared__ double gi[];   // [QRB][nv]
  __shared__ double scratch[40];
  __shared__ double xi[QRB];
  const int c = blockIdx.y;
  const int nblk = (N + QRB - 1) / QRB;
  const int rb0 = blockIdx.x * rbpc, rb1 = min(nblk, rb0 + rbpc);
  const double* Gc = G + (long)c * N * nv;
  const double* Zc = Z + (long)c * N * nv;
  const double* xs = x + (long)c * N;
  const double ib2 = 1.0 / (beta * beta);
  double sg = 0.0, gg = 0.0, rg = 0.0, gbg = 0.0;
  for (int rb = rb0; rb < rb1; ++rb) {
    const int i0 = rb * QRB;
    const int nr = min(QRB, N - i0);
    __syncthreads();                                        // the previous row block's gi / xi are no longer read
    for (int idx = threadIdx.x; idx < QRB * nv; idx += blockDim.x) {
      double g = 0.0;
      if (idx < nr * nv) {
        g = Gc[(long)i0 * nv + idx];
        const double z = Zc[(long)i0 * nv + idx];
        sg += g; gg += g * g; rg += z * z;
      }
      gi[idx] = g;                                          // rows past N: zero, their weights drop out
    }
    if (threadIdx.x < QRB) xi[threadIdx.x] = threadIdx.x < nr ? xs[i0 + threadIdx.x] : 0.0;
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      double dot[QRB];
#pragma unroll
      for (int r = 0; r < QRB; ++r) dot[r] = 0.0;
      const double* gj = Gc + (long)j * nv;
#pragma unroll 4
      for (int t = 0; t < nv; ++t) {           // several loads of the G_j row in flight: the kernel is load-latency bound
        const double gjt = gj[t];
#pragma unroll
        for (int r = 0; r < QRB; ++r) dot[r] += gi[r * nv + t] * gjt;
      }
      const double xj = xs[j];
#pragma unroll
      for (int r = 0; r < QRB; ++r) {
        const double d = xi[r] - xj, d2 = d * d;            // D_ii = 0: the diagonal drops out by itself
        gbg += exp(-0.5 * d2 * ib2) * d2 * dot[r];
      }
    }
  }
  gbg *= alpha2 / (beta * beta * beta);
  const double a = block_sum(sg, scratch), b = block_sum(gg, scratch), r = block_sum(rg, scratch), q = block_sum(gbg, scratch);
  if (threadIdx.x == 0) {
    double* o = part + ((long)c * gridDim.x + blockIdx.x) * 4;
    o[0] = a; o[1] = b; o[2] = r; o[3] = q;
  }
}

// hgrad[c][k] = d(-logpost_c)/d hyper_k in the plan's hyper-parameter order (include/nmgp_b200.h)
__global__ void hyper_finish_kernel(int model, int cs, int N, int M, int P, const double* __restrict__ pars, HyperRaw h,
                                    const double* __restrict__ s2v, const double* __restrict__ q0, const double* __restrict__ q1,
                                    const double* __restrict__ trI0, const double* __restrict__ trB0,
                                    const double* __restrict__ trI1, const double* __restrict__ trB1, int nv1, int nblk,
                                    double* __restrict__ hgrad) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cs) return;
  double* o = hgrad + (long)c * 9;
  for (int k = 0; k < 9; ++k) o[k] = 0.0;
  if (!h.prior) return;
  const double s2 = s2v[c];
  const int T = tril_size(M);
  const double* p = pars + (long)c * P;
  // inverse gamma on sigma2_err: (-a-1) log s2 - b/s2 + a log b - lgamma(a)
  const double d_a = -log(s2) + log(h.b) - h.digamma_a, d_b = -1.0 / s2 + h.a / h.b;
  auto gp = [&](const double* q, double trI, double trB, double alpha, int nv, double& d_mu, double& d_alpha, double& d_beta) {
    double sg = 0.0, gg = 0.0, rg = 0.0, gbg = 0.0;
    for (int bk = 0; bk < nblk; ++bk) { sg += q[bk * 4 + 0]; gg += q[bk * 4 + 1]; rg += q[bk * 4 + 2]; gbg += q[bk * 4 + 3]; }
    d_mu = sg;
    d_alpha = (1.0 / alpha) * (-(double)nv * ((double)N - kJitter * trI) + (rg - kJitter * gg));
    d_beta = -0.5 * (double)nv * trB + 0.5 * gbg;
  };
  if (model == 0) {          // stationary {mu_tilde_l, sigma_tilde_l, a, b, c}
    const double d = p[0] - h.hy[0], sd = h.hy[1], cc = h.hy[4];
    double su = 0.0;
    for (int t = 0; t < T; ++t) su += p[2 + t] * p[2 + t];
    o[0] = -(d / (sd * sd));
    o[1] = -(d * d / (sd * sd * sd) - 1.0 / sd);
    o[2] = -d_a; o[3] = -d_b;
    o[4] = -(su / (cc * cc * cc) - (double)T / cc);
    return;
  }
  double m0, a0, b0, m1, a1, b1;
  gp(q0 + (long)c * nblk * 4, trI0[c], trB0[c], h.hy[1], 1, m0, a0, b0);
  gp(q1 + (long)c * nblk * 4, trI1[c], trB1[c], h.hy[4], nv1, m1, a1, b1);
  o[0] = -m0; o[1] = -a0; o[2] = -b0; o[3] = -m1; o[4] = -a1; o[5] = -b1;
  o[6] = -d_a; o[7] = -d_b;
  if (model == 1) {          // separable: Normal(0, c) on uL
    const double cc = h.hy[8];
    double su = 0.0;
    for (int t = 0; t < T; ++t) su += p[2 * N + t] * p[2 * N + t];
    o[8] = -(su / (cc * cc * cc) - (double)T / cc);
  }
}

// One rank's contribution to the sweep's all-reduce, in one launch and in a fixed summation order:
// out[0..5] = sum over the successful subjects (info == 0) of vals[s][0..5], out[6] = #failed, out[7] = #subjects,
// out[8..16] = sum over the successful subjects of hgrad[s][0..8] (zeros when hgrad == NULL).
__global__ void __launch_bounds__(1024) sweep_reduce_kernel(const double* __restrict__ vals, const double* __restrict__ hgrad,
                                                            const int* __restrict__ info, long S, double* __restrict__ out) {
  __shared__ double scratch[40];
  double acc[17];
#pragma unroll
  for (int k = 0; k < 17; ++k) acc[k] = 0.0;
  for (long s = threadIdx.x; s < S; s += blockDim.x) {
    if (info[s] != 0) { acc[6] += 1.0; continue; }
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k] += vals[s * 6 + k];
    if (hgrad) {
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[8 + k] += hgrad[s * 9 + k];
    }
  }
  acc[7] = 0.0;
#pragma unroll
  for (int k = 0; k < 17; ++k) {
    const double t = block_sum(acc[k], scratch);
    if (threadIdx.x == 0) out[k] = (k == 7) ? (double)S : t;
  }
}

}  // namespace

int launch_sweep_reduce(const double* vals, const double* hgrad, const int* info, long S, double* out, cudaStream_t st) {
  sweep_reduce_kernel<<<1, 1024, 0, st>>>(vals, hgrad, info, S, out);
  NMGP_CUDA_TRY(cudaGetLastError());
  return 0;
}

#define NMGP_LAUNCH_CHECK()                 \
  do {                                      \
    NMGP_CUDA_TRY(cudaGetLastError());      \
    if (launches) ++*launches;              \
  } while (0)

// tr(Sigma_p^-1) and tr(Sigma_p^-1 dSigma_p/dbeta) for `cs` subjects from their cached prior factors.
// scratch: 3 * cs * N * N + 2 * cs * N doubles.
int prior_traces(const double* x, const double* Lp, int cs, int N, double alpha, double beta, double* scratch, double* trI,
                 double* trB, cudaStream_t st, long* launches) {
  if (cs <= 0) return 0;
  const size_t NN = (size_t)cs * N * N;
  double* I = scratch;
  double* W = I + NN;
  double* Bd = W + NN;
  double* pI = Bd + NN;
  double* pB = pI + (size_t)cs * N;
  eye_kernel<<<(unsigned)((NN + 255) / 256), 256, 0, st>>>(I, N, (long)NN);
  NMGP_LAUNCH_CHECK();
  NMGP_TRY(launch_prior_solve(Lp, I, W, cs, N, N, 0, st, launches));
  dim3 gb((N + 127) / 128, N, cs);
  dbeta_cov_kernel<<<gb, 128, 0, st>>>(x, N, alpha * alpha, beta, Bd);
  NMGP_LAUNCH_CHECK();
  dim3 gt(N, cs);
  trace_rows_kernel<<<gt, 128, 0, st>>>(W, Bd, N, pI, pB);
  NMGP_LAUNCH_CHECK();
  sum_rows_kernel<<<cs, 256, 0, st>>>(pI, pB, N, trI, trB);
  NMGP_LAUNCH_CHECK();
  return 0;
}

int prior_quad_blocks(int N) { return (N + QRB - 1) / QRB; }

// row blocks per CTA and the resulting number of partial sums per subject for a launch over `cs` subjects
static int quad_rbpc(int N, int cs) {
  const int nblk = prior_quad_blocks(N);
  long rbpc = (long)nblk * cs / 4096;
  if (rbpc < 1) rbpc = 1;
  if (rbpc > nblk) rbpc = nblk;
  return (int)rbpc;
}
static int quad_parts(int N, int cs) { const int r = quad_rbpc(N, cs); return (prior_quad_blocks(N) + r - 1) / r; }

int launch_prior_quad(const double* x, const double* Z, const double* G, int cs, int N, int nv, double alpha, double beta,
                      double* out, cudaStream_t st, long* launches) {
  if (cs <= 0) return 0;
  dim3 grid(quad_parts(N, cs), cs);
  prior_quad_kernel<<<grid, 128, (size_t)QRB * nv * sizeof(double), st>>>(x, Z, G, N, nv, alpha * alpha, beta, quad_rbpc(N, cs),
                                                                          out);
  NMGP_LAUNCH_CHECK();
  return 0;
}

int launch_hyper_finish(int model, int cs, int N, int M, int P, const double* pars, const HyperRaw& h, const double* s2v,
                        const double* q0, const double* q1, const double* trI0, const double* trB0, const double* trI1,
                        const double* trB1, int nv1, double* hgrad, cudaStream_t st, long* launches) {
  if (cs <= 0) return 0;
  hyper_finish_kernel<<<(cs + 127) / 128, 128, 0, st>>>(model, cs, N, M, P, pars, h, s2v, q0, q1, trI0, trB0, trI1, trB1, nv1,
                                                        quad_parts(N, cs), hgrad);
  NMGP_LAUNCH_CHECK();
  return 0;
}

}  // namespace nmgp
