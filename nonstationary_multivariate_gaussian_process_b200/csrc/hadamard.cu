// "Hadamard" objectives: irregularly sampled data, ONE observation per row -- (x_n, indx_n, y_n), indx_n the output the
// observation belongs to (Utility/logpos.py:465-716: nlogpos_obj_hadamard, nlogpos_obj_hadamard_SVC, nlogpos_obj_hadamard_S).
//
// All three share one structure: with r_n the row `indx_n` of the cross-output factor that applies at observation n
// (the shared L for the separable / stationary variants, L_n for the SVC variant; raw entries, no exp on the diagonal:
// logpos.py:518, 582-583, 682),
//     Sigma[n,n'] = K_x[n,n'] * <r_n, r_n'> + sigma2_err [n = n'],      K_x the Gibbs / RBF kernel incl. jitter,
// an N x N dense matrix that goes through the same batched Cholesky + inverse engine as every other model.  With
// G = -0.5 Sigma^-1 + 0.5 alpha alpha^T, alpha = Sigma^-1 y:
//     d/dR[n][k]           = 2 sum_n' G[n,n'] K_x[n,n'] R[n'][k]
//     d/dtilde_l[n]        = 2 <r_n, sum_n' G[n,n'] CK[n,n'] r_n'>              CK = c_nn' K0  (K0 = kernel without jitter)
//     d/dtilde_sigma[n]    = <r_n, dR[n]> - 2 jitter G[n,n] |r_n|^2             (only the jitter separates K_x from K0)
//     d/dtilde_sigma2_err  = sigma2_err tr G
// and the chain rule onto the parameter layout of each variant is a scatter (finish kernel).
#include "models.cuh"

namespace nmgp {

namespace {

constexpr int RM = 16;   // row factors are stored padded to 16 entries (M <= 16)

__device__ __forceinline__ void tril_unrank(int t, int& m, int& k) {
  m = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while (m * (m + 1) / 2 > t) --m;
  while ((m + 1) * (m + 2) / 2 <= t) ++m;
  k = t - m * (m + 1) / 2;
}

// pars -> ell, sig, sigma2, row factors R, prior residuals.  variant: 0 separable (pars = [tilde_l(N), tilde_sigma(N),
// L_vec(T), ts2], logpos.py:479), 1 SVC ([tilde_l(N), L_vecs(N*T), ts2], :60-72), 2 stationary ([tilde_l, tilde_sigma,
// L_vec(T), ts2], :657).
__global__ void had_prep_kernel(int variant, const double* __restrict__ pars, int P, int N, int M,
                                const int* __restrict__ indx, double mu0, double mu1, double* __restrict__ ell,
                                double* __restrict__ sig, double* __restrict__ s2, double* __restrict__ R,
                                double* __restrict__ Rt, double* __restrict__ R0, double* __restrict__ R1) {
  const int c = blockIdx.x;
  const int T = tril_size(M);
  const double* p = pars + (long)c * P;
  if (threadIdx.x == 0) s2[c] = exp(p[P - 1]);
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const double tl = variant == 2 ? p[0] : p[n];
    const double ts = variant == 0 ? p[N + n] : (variant == 2 ? p[1] : 0.0);
    ell[(long)c * N + n] = exp(tl);
    sig[(long)c * N + n] = exp(ts);                      // SVC: exp(0) = 1 (sigma1 defaults to ones, kernels.py:56-57)
    if (variant != 2) R0[(long)c * N + n] = tl - mu0;
    if (variant == 0) R1[(long)c * N + n] = ts - mu1;
    const double* Lv = variant == 0 ? p + 2 * N : (variant == 1 ? p + N + (long)n * T : p + 2);
    const int m = indx[(long)c * N + n];
    for (int k = 0; k < RM; ++k) {
      const double r = (k <= m && m < M) ? Lv[m * (m + 1) / 2 + k] : 0.0;
      R[((long)c * N + n) * RM + k] = r;                 // [n][k]: a row at a time (finish kernel)
      Rt[((long)c * RM + k) * N + n] = r;                // [k][n]: coalesced over the observations (build, contraction)
    }
  }
  if (variant == 1)
    for (int idx = threadIdx.x; idx < N * T; idx += blockDim.x) R1[(long)c * N * T + idx] = p[N + idx] - mu1;
}

// Sigma in the engine's padded layout (both triangles, identity padding).  A CTA writes a patch of HB_ROWS rows x 128
// columns: each thread keeps the factor row of ITS column in registers and walks the patch's rows, whose factor rows sit in
// shared memory -- per element one K_x load, M FMAs and one store (the first version re-read M factor entries from L2 for
// every element: 17 ms for 3000 x (N = 600), a fifth of the HBM rate).
constexpr int HB_ROWS = 16;
template <int MM>
__global__ void __launch_bounds__(128) had_build_kernel(const double* __restrict__ Kx, const double* __restrict__ R,
                                                        const double* __restrict__ Rt, const double* __restrict__ s2v, int N,
                                                        double* __restrict__ A, long strideA, int ld) {
  __shared__ double rs[HB_ROWS][MM];
  const int q = blockIdx.x * 128 + threadIdx.x;
  const int p0 = blockIdx.y * HB_ROWS;
  const int c = blockIdx.z;
  for (int idx = threadIdx.x; idx < HB_ROWS * MM; idx += 128) {
    const int pr = idx / MM, k = idx % MM;
    rs[pr][k] = (p0 + pr < N) ? R[((long)c * N + p0 + pr) * RM + k] : 0.0;
  }
  double rq[MM];
#pragma unroll
  for (int k = 0; k < MM; ++k) rq[k] = (q < N) ? Rt[((long)c * RM + k) * N + q] : 0.0;
  __syncthreads();
  if (q >= ld) return;
  const double s2 = s2v[c];
  double* Ac = A + (long)c * strideA;
#pragma unroll 4
  for (int pr = 0; pr < HB_ROWS; ++pr) {
    const int p = p0 + pr;
    if (p >= ld) break;
    double v;
    if (p < N && q < N) {
      double ki = 0.0;
#pragma unroll
      for (int k = 0; k < MM; ++k) ki += rs[pr][k] * rq[k];
      v = Kx[((long)c * N + p) * N + q] * ki;
      if (p == q) v += s2;
    } else {
      v = (p == q) ? 1.0 : 0.0;
    }
    Ac[(long)p * ld + q] = v;
  }
}

// One warp per row n, eight rows per CTA; the transposed factor table Rt [M][N] of the subject is staged in shared memory
// once per CTA (it was re-read from L2 by every lane for every element: 104 GB of L2 traffic per sweep at N = 600).
// PASS 0 (weights = row n of Sigma^-1): alpha[n], W[n][k] = sum_n' z K_x R[n'][k], V[n][k] = sum z CK R.
// PASS 1 (weights = alpha): Sa[n][k] = sum_n' alpha_n' K_x[n,n'] R[n'][k], Ca[n][k] = sum alpha_n' CK[n,n'] R[n'][k].
template <int PASS, int MM>
__global__ void __launch_bounds__(256) had_contract_kernel(const double* __restrict__ A, long strideA, int ld, int N,
                                                           const double* __restrict__ y, const double* __restrict__ Kx,
                                                           const double* __restrict__ CK, const double* __restrict__ Rt,
                                                           double* __restrict__ alpha, double* __restrict__ Wo,
                                                           double* __restrict__ Vo, int rt_in_smem) {
  extern __shared__ __align__(16) double rts[];      // [MM][N] when rt_in_smem
  const int c = blockIdx.y;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const double* Rg = Rt + (long)c * RM * N;
  if (rt_in_smem) {
    for (int idx = threadIdx.x; idx < MM * N; idx += 256) rts[idx] = Rg[idx];     // rows k < MM of the [RM][N] table are contiguous
    __syncthreads();
  }
  if (n >= N) return;
  const double* Rc = rt_in_smem ? rts : Rg;
  const double* zr = A + (long)c * strideA + (long)n * ld;
  const double* kr = Kx + ((long)c * N + n) * N;
  const double* cr = CK + ((long)c * N + n) * N;
  const double* al = alpha + (long)c * N;
  const double* yc = y + (long)c * N;
  double w[MM], v[MM], a = 0.0;
#pragma unroll
  for (int k = 0; k < MM; ++k) w[k] = v[k] = 0.0;
  // the row is streamed once from HBM by this warp alone: four 32-wide slices are loaded before any arithmetic so that
  // enough bytes are in flight (one slice per iteration ran at 1 TB/s)
  for (int j0 = 0; j0 < N; j0 += 128) {
    double z[4], zk[4], zc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u + lane;
      const bool ok = j < N;
      z[u] = ok ? (PASS == 0 ? zr[j] : al[j]) : 0.0;
      zk[u] = ok ? kr[j] : 0.0;
      zc[u] = ok ? cr[j] : 0.0;
      if (PASS == 0 && ok) a += z[u] * yc[j];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u + lane;
      if (j >= N) continue;
      const double wk = z[u] * zk[u], wc = z[u] * zc[u];
#pragma unroll
      for (int k = 0; k < MM; ++k) {
        const double r = Rc[(long)k * N + j];
        w[k] += wk * r;
        v[k] += wc * r;
      }
    }
  }
  if (PASS == 0) {
    a = warp_sum(a);
    if (lane == 0) alpha[(long)c * N + n] = a;
  }
#pragma unroll
  for (int k = 0; k < MM; ++k) {
    const double ws = warp_sum(w[k]), vs = warp_sum(v[k]);
    if (lane == 0) {
      Wo[((long)c * N + n) * RM + k] = ws;
      Vo[((long)c * N + n) * RM + k] = vs;
    }
  }
  if (lane == 0)
    for (int k = MM; k < RM; ++k) {      // the finish kernel reads all RM slots
      Wo[((long)c * N + n) * RM + k] = 0.0;
      Vo[((long)c * N + n) * RM + k] = 0.0;
    }
}

// values, priors and the gradient in the parameter layout of the variant (one CTA per subject)
__global__ void __launch_bounds__(256) had_finish_kernel(
    int variant, int N, int M, int P, const double* __restrict__ pars, const double* __restrict__ y,
    const int* __restrict__ indx, HyperConst h, const double* __restrict__ A, long strideA, int ld,
    const double* __restrict__ logdet, const int* __restrict__ info_in, const double* __restrict__ alpha,
    const double* __restrict__ s2v, const double* __restrict__ R, const double* __restrict__ Wo,
    const double* __restrict__ Vo, const double* __restrict__ Sa, const double* __restrict__ Ca,
    const double* __restrict__ Z0, const double* __restrict__ Z1, const double* __restrict__ G0,
    const double* __restrict__ G1, const double* __restrict__ hld0, const double* __restrict__ hld1,
    double* __restrict__ vals, double* __restrict__ grad, int* __restrict__ info) {
  __shared__ double scratch[40];
  __shared__ double dL[16 * 17 / 2 + 8];
  const int c = blockIdx.x;
  const int T = tril_size(M);
  const double* p = pars + (long)c * P;
  const double* al = alpha + (long)c * N;
  const double* yc = y + (long)c * N;
  const double* Ac = A + (long)c * strideA;
  const double* Rc = R + (long)c * N * RM;
  const int* ix = indx + (long)c * N;
  double q = 0.0, tr = 0.0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const double a = al[n];
    q += yc[n] * a;
    tr += -0.5 * Ac[(long)n * ld + n] + 0.5 * a * a;
  }
  const double quad = block_sum(q, scratch);
  const double trG = block_sum(tr, scratch);
  const double s2 = s2v[c], ts2 = p[P - 1];
  const double log2pi = 1.8378770664093453;
  const double loglik = -0.5 * logdet[c] - 0.5 * quad;                           // distributions.py:22
  // ---- priors
  double lp_l = 0.0, lp_2 = 0.0, lp_L = 0.0;
  const double* Lv = variant == 0 ? p + 2 * N : p + 2;                           // shared L_vec (variants 0, 2)
  if (variant == 2) {
    const double d = p[0] - h.s_loc;
    lp_l = -(d * d) / (2.0 * h.s_var) - h.s_logscale - h.half_log2pi;            // Normal(mu, sigma).log_prob, logpos.py:700
  } else {
    double z0 = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) { const double z = Z0[(long)c * N + n]; z0 += z * z; }
    lp_l = -0.5 * (N * log2pi + block_sum(z0, scratch)) - hld0[c];               // MVN.log_prob, logpos.py:539, 610
  }
  if (variant == 0) {
    double z1 = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) { const double z = Z1[(long)c * N + n]; z1 += z * z; }
    lp_2 = -0.5 * (N * log2pi + block_sum(z1, scratch)) - hld1[c];               // tilde_sigma, logpos.py:544-545
  }
  if (variant == 1) {
    double z1 = 0.0;
    for (int idx = threadIdx.x; idx < N * T; idx += blockDim.x) { const double z = Z1[(long)c * N * T + idx]; z1 += z * z; }
    lp_L = -0.5 * ((double)T * N * log2pi + block_sum(z1, scratch)) - (double)T * hld1[c];   // T columns, logpos.py:617
  } else {
    double lpu = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const double d = Lv[t] - h.n_loc;
      lpu += -(d * d) / (2.0 * h.n_var) - h.n_logscale - h.half_log2pi;          // Normal(0, c), logpos.py:549, 704
    }
    lp_L = block_sum(lpu, scratch);
  }
  const double lp_s = (-h.ig_a - 1.0) * log(s2) - h.ig_b / s2;   // inverse_gamma_logpdf_u: no normaliser (distributions.py:116-124)
  double res = loglik;
  if (h.prior) res += lp_l + lp_2 + lp_L + lp_s + ts2;
  const double pf = h.prior ? 1.0 : 0.0;
  if (threadIdx.x == 0) {
    double* v = vals + (long)c * 6;
    v[0] = -res; v[1] = loglik; v[2] = lp_l;
    if (variant == 0) { v[3] = lp_2; v[4] = lp_L; v[5] = lp_s; }
    else { v[3] = lp_L; v[4] = lp_s; v[5] = 0.0; }
    info[c] = info_in[c];
  }
  if (grad == nullptr) return;
  double* g = grad + (long)c * P;
  // ---- per-observation pieces: dR[n][k] = -W + alpha_n Sa;  dl[n] = 2 <r_n, -0.5 V + 0.5 alpha_n Ca>
  for (int t = threadIdx.x; t < T; t += blockDim.x) dL[t] = 0.0;
  __syncthreads();
  double sum_l = 0.0, sum_s = 0.0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const double a = al[n];
    const double Gnn = -0.5 * Ac[(long)n * ld + n] + 0.5 * a * a;
    const int m = ix[n];
    double dl = 0.0, rdr = 0.0, rr = 0.0;
    for (int k = 0; k < M; ++k) {
      const double r = Rc[(long)n * RM + k];
      const double dr = -Wo[((long)c * N + n) * RM + k] + a * Sa[((long)c * N + n) * RM + k];
      dl += r * (-0.5 * Vo[((long)c * N + n) * RM + k] + 0.5 * a * Ca[((long)c * N + n) * RM + k]);
      rdr += r * dr;
      rr += r * r;
      if (k <= m) {
        if (variant == 1) g[N + (long)n * T + m * (m + 1) / 2 + k] = -(dr - pf * G1[(long)c * N * T + (long)n * T + m * (m + 1) / 2 + k]);
        else atomicAdd(&dL[m * (m + 1) / 2 + k], dr);
      }
    }
    const double ds = rdr - 2.0 * kJitter * Gnn * rr;
    if (variant == 0) {
      g[n] = -(2.0 * dl - pf * G0[(long)c * N + n]);
      g[N + n] = -(ds - pf * G1[(long)c * N + n]);
    } else if (variant == 1) {
      g[n] = -(2.0 * dl - pf * G0[(long)c * N + n]);
      // entries of L_n outside row indx_n only see their prior
      for (int t = 0; t < T; ++t) {
        int mm, kk;
        tril_unrank(t, mm, kk);
        if (mm != m) g[N + (long)n * T + t] = pf * G1[(long)c * N * T + (long)n * T + t];
      }
    } else {
      sum_l += 2.0 * dl;
      sum_s += ds;
    }
  }
  if (variant == 2) {
    const double sl = block_sum(sum_l, scratch), ss = block_sum(sum_s, scratch);
    if (threadIdx.x == 0) {
      g[0] = -(sl + pf * (-(p[0] - h.s_loc) / h.s_var));
      g[1] = -ss;
    }
  }
  __syncthreads();
  if (variant != 1) {
    double* gL = variant == 0 ? g + 2 * N : g + 2;
    for (int t = threadIdx.x; t < T; t += blockDim.x) gL[t] = -(dL[t] + pf * (-(Lv[t] - h.n_loc) / h.n_var));
  }
  if (threadIdx.x == 0) g[P - 1] = -(s2 * trG + pf * ((-h.ig_a - 1.0) + h.ig_b / s2 + 1.0));
}

}  // namespace

#define NMGP_LAUNCH_CHECK()                 \
  do {                                      \
    NMGP_CUDA_TRY(cudaGetLastError());      \
    if (launches) ++*launches;              \
  } while (0)

int had_forward(int variant, int cs, int N, int M, const double* x, const int* indx, const double* pars, int P,
                const HyperConst& h, const Scratch& w, const BlockBatch& b, cudaStream_t st, long* launches) {
  if (cs <= 0) return 0;
  had_prep_kernel<<<cs, 256, 0, st>>>(variant, pars, P, N, M, indx, h.mu0, h.mu1, w.ell, w.sig, w.s2, w.Lst, w.Ua, w.R0, w.R1);
  NMGP_LAUNCH_CHECK();
  NMGP_TRY(launch_kx(x, w.ell, w.sig, cs, N, w.Kx, w.CK, st, launches));
  dim3 gb((b.nP + 127) / 128, (b.nP + HB_ROWS - 1) / HB_ROWS, cs);
#define NMGP_HB_CASE(MM) case MM: had_build_kernel<MM><<<gb, 128, 0, st>>>(w.Kx, w.Lst, w.Ua, w.s2, N, b.A, b.strideA(), b.nP); break;
  switch (M) { NMGP_FOR_EACH_M(NMGP_HB_CASE) default: set_last_error("M out of range"); return -1; }
#undef NMGP_HB_CASE
  NMGP_LAUNCH_CHECK();
  return 0;
}

int had_backward(int variant, int cs, int N, int M, const double* y, const int* indx, const double* pars, int P,
                 const HyperConst& h, const Scratch& w, const BlockBatch& b, const double* hld0, const double* hld1,
                 double* vals, double* grad, int* info, cudaStream_t st, long* launches) {
  if (cs <= 0) return 0;
  dim3 gc((N + 7) / 8, cs);
  // the factor table in shared memory when it fits beside a useful number of CTAs per SM (M N doubles <= 96 KB)
  const size_t rt_bytes = (size_t)M * N * sizeof(double);
  const int in_smem = rt_bytes <= 96u * 1024u ? 1 : 0;
  const size_t smem = in_smem ? rt_bytes : 0;
#define NMGP_HC_CASE(MM)                                                                                                      \
  case MM:                                                                                                                    \
    if (smem > 48u * 1024u) {                                                                                                 \
      NMGP_SMEM_ATTR_PER_DEVICE((had_contract_kernel<0, MM>), smem);                                                          \
      NMGP_SMEM_ATTR_PER_DEVICE((had_contract_kernel<1, MM>), smem);                                                          \
    }                                                                                                                         \
    had_contract_kernel<0, MM><<<gc, 256, smem, st>>>(b.A, b.strideA(), b.nP, N, y, w.Kx, w.CK, w.Ua, w.alpha, w.Wout, w.Vout, \
                                                      in_smem);                                                               \
    NMGP_LAUNCH_CHECK();                                                                                                      \
    if (grad != nullptr) {                                                                                                    \
      had_contract_kernel<1, MM><<<gc, 256, smem, st>>>(b.A, b.strideA(), b.nP, N, y, w.Kx, w.CK, w.Ua, w.alpha, w.Sa, w.Ca,   \
                                                        in_smem);                                                             \
      NMGP_LAUNCH_CHECK();                                                                                                    \
    }                                                                                                                         \
    break;
  switch (M) { NMGP_FOR_EACH_M(NMGP_HC_CASE) default: set_last_error("M out of range"); return -1; }
#undef NMGP_HC_CASE
  had_finish_kernel<<<cs, 256, 0, st>>>(variant, N, M, P, pars, y, indx, h, b.A, b.strideA(), b.nP, b.logdet, b.info, w.alpha,
                                        w.s2, w.Lst, w.Wout, w.Vout, w.Sa, w.Ca, w.Z0, w.Z1, w.G0, w.G1, hld0, hld1, vals, grad,
                                        info);
  NMGP_LAUNCH_CHECK();
  return 0;
}

}  // namespace nmgp
