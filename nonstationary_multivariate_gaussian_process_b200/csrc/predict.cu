// Posterior prediction for the nonseparable model (Utility/prediction.py:1038-1262).
//
// The reference rebuilds the n x n covariance, diagonalises it (`symeig`) and Cholesky-factors its inverse for EVERY grid
// point and EVERY sample (G * n_sample times, prediction.py:1130-1160) although none of that depends on them.  Here the
// hot-path engine factors and inverts Sigma once per subject, and the per-sample work is reduced to its sample-dependent
// core.  With L* the sampled M x M factor, l* the sampled length scale and kx[i] = Gibbs(x_i, x*; l_i, l*):
//
//   k_f[(i,m), m'] = kx[i] * sum_k L_i[m,k] L*[m',k]                                             (prediction.py:1148-1150)
//   mu_f[m']   = sum_k L*[m',k] q[k],            q[k]    = sum_i kx[i] (L_i^T alpha_i)[k],  alpha = Sigma^-1 y
//   B[m',m']   = sum_{k,k'} L*[m',k] L*[m',k'] Q[k,k'],  Q[k,k'] = sum_{i,j} kx[i] kx[j] H_ij[k,k'],  H_ij = L_i^T (Sigma^-1)_ij L_j
//   sigma2_y   = (1 + jitter) (L* L*^T)[m',m'] - B[m',m'] + sigma2_err                              (prediction.py:1153-1161)
//
// H (M x M per time pair) does not depend on the sample: it is formed once per subject, stored as T = M(M+1)/2 dense
// N x N matrices (Q is symmetric, pairs k <= k' suffice), and Q for all G * n_sample columns is a batched
// [N x N] x [N x C] product followed by a column-wise dot with kx -- T * 2 N^2 C flop instead of the 2 n^2 M C of the dense
// form (2M/(1+1/M) times fewer), on FP64 tensor cores (DMMA.8x8x4, cp.async double buffer).
#include "models.cuh"

namespace nmgp {

namespace {

constexpr int QT = 64;    // output tile (rows i, columns = samples)
constexpr int QK = 16;    // k-chunk
constexpr int QLD = 20;   // padded shared-memory row (doubles): conflict-free 8-byte fragment loads, 16-byte aligned rows

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------ GP-prior conditionals
// Conditional moments of one GP prior at G new inputs per subject (prediction.py:1060-1068 tilde_l, :1070-1081 uL):
//   k = alpha^2 exp(-0.5 |(x - x*)/beta|^2),  w = L^-1 k (forward substitution with the plan's cached factor),
//   mean_t = mu + w . z_t with z_t = L^-1 (v_t - mu) (already formed for the prior density),  sigma2 = (jitter + alpha^2) - w . w,
//   negative sigma2 -> 1e-6.  One WARP per (subject, new input): w lives in shared memory, row i of L is read coalesced and
//   the lanes split its dot product with w[0..i) (N dependent steps of a 5-shuffle reduction instead of N^2/2 serial FMAs).
__global__ void __launch_bounds__(128) pred_prior_kernel(const double* __restrict__ x, const double* __restrict__ Lp,
                                                         const double* __restrict__ Z, int N, int nv,
                                                         const double* __restrict__ xstar, int G, double alpha2,
                                                         double beta, double mu, double* __restrict__ mean_out,
                                                         double* __restrict__ s2_out) {
  extern __shared__ __align__(16) double psm[];
  const int wpb = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int c = blockIdx.y;
  if (g >= G) return;
  double* w = psm + (size_t)(threadIdx.x >> 5) * N;
  const double* xs = x + (long)c * N;
  const double* L = Lp + (long)c * N * N;
  const double b = xstar[(long)c * G + g] / beta;                       // kernels.py:38-39
  for (int i = lane; i < N; i += 32) {
    const double a = xs[i] / beta;
    w[i] = __dmul_rn(exp(__dmul_rn(-0.5, ref_sqdist(a, b))), alpha2);
  }
  __syncwarp();
  double ww = 0.0;
  for (int i = 0; i < N; ++i) {
    const double* Li = L + (long)i * N;
    double s0 = 0.0, s1 = 0.0;
    int j = lane;
    for (; j + 32 < i; j += 64) {
      s0 += Li[j] * w[j];
      s1 += Li[j + 32] * w[j + 32];
    }
    if (j < i) s0 += Li[j] * w[j];
    const double s = warp_sum(s0 + s1);
    const double wi = (w[i] - s) / Li[i];
    __syncwarp();
    if (lane == 0) w[i] = wi;
    __syncwarp();
    ww += wi * wi;
  }
  if (lane == 0) {
    double s2v = __dadd_rn(kJitter, alpha2) - ww;                       // RBF_cov(x*)[0,0] - proj . k
    if (s2v < 0.0) s2v = 1e-6;                                          // settings.precision, prediction.py:1066,1081
    s2_out[(long)c * G + g] = s2v;
  }
  const double* Zc = Z + (long)c * N * nv;
  for (int t = 0; t < nv; ++t) {
    double m = 0.0;
    for (int i = lane; i < N; i += 32) m += w[i] * Zc[(long)i * nv + t];
    m = warp_sum(m);
    if (lane == 0) mean_out[((long)c * G + g) * nv + t] = mu + m;
  }
}

// ------------------------------------------------------------------------------------------------ per-subject tables
// u_j = L_j^T alpha_j  -> U[c][j][k]   (alpha time-major [c][n])
__global__ void pred_u_kernel(const double* __restrict__ alpha, const double* __restrict__ Lst, int N, int M, int MT,
                              long total, double* __restrict__ U) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = (int)(idx % MT);
  const long cj = idx / MT;            // c * N + j
  double u = 0.0;
  if (k < M)
    for (int m = k; m < M; ++m) u += alpha[cj * M + m] * Lst[(cj * M + m) * MT + k];
  U[idx] = u;
}

// H_ij = L_i^T (Sigma^-1)_ij L_j for all time pairs, stored as T matrices Hm[pair(k<=k')][i][j] with zero padding to
// [Ni64][Npad].  One thread per (i, k', j): column k' of (Sigma^-1)_ij L_j, then the k <= k' entries of L_i^T (.).
__global__ void __launch_bounds__(128) pred_h_kernel(const double* __restrict__ A, long strideA, int ld,
                                                     const double* __restrict__ Lst, int N, int M, int MT, int Ni64,
                                                     int Npad, int c0, double* __restrict__ Hm, long hm_stride) {
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int i = blockIdx.y / M, kp = blockIdx.y % M;
  const int cc = blockIdx.z, c = c0 + cc;
  if (j >= Npad) return;
  double* H = Hm + (long)cc * hm_stride + (long)(kp * (kp + 1) / 2) * Ni64 * Npad + (long)i * Npad + j;
  const long pstride = (long)Ni64 * Npad;
  if (i >= N || j >= N) {
    for (int k = 0; k <= kp; ++k) H[k * pstride] = 0.0;
    return;
  }
  const double* Ac = A + (long)c * strideA + (long)(i * M) * ld + (long)j * M;
  const double* Lj = Lst + ((long)c * N + j) * M * MT;
  const double* Li = Lst + ((long)c * N + i) * M * MT;
  double tmp[16];
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    double t = 0.0;
    if (m < M)
      for (int m2 = kp; m2 < M; ++m2) t += Ac[(long)m * ld + m2] * Lj[m2 * MT + kp];
    tmp[m] = t;
  }
  for (int k = 0; k <= kp; ++k) {
    double h = 0.0;
#pragma unroll
    for (int m = 0; m < 16; ++m)
      if (m >= k && m < M) h += Li[m * MT + k] * tmp[m];
    H[k * pstride] = h;
  }
}

// KXT[cc][col][j] = Gibbs(x_j, x*_g; l_j, l*_col) (no jitter: cross-covariance, kernels.py:62-63), col = g * ns + s;
// zero for padded j / col.
// Separable / stationary models: sig != NULL, the pair carries sigma_j * exp(tilde_sigma*_col) (kernels.py:71).
__global__ void __launch_bounds__(128) pred_kx_kernel(const double* __restrict__ x, const double* __restrict__ ell,
                                                      const double* __restrict__ sig, const double* __restrict__ xstar,
                                                      const double* __restrict__ tl_star, const double* __restrict__ ts_star,
                                                      int N, int Npad, int G, int ns, int C, int c0,
                                                      double* __restrict__ KXT, long kx_stride) {
  const int col = blockIdx.x;
  const int j = blockIdx.y * 128 + threadIdx.x;
  const int cc = blockIdx.z, c = c0 + cc;
  if (j >= Npad) return;
  double v = 0.0;
  if (j < N && col < C) {
    const double ls = exp(tl_star[(long)c * C + col]);
    const double sij = sig ? __dmul_rn(sig[(long)c * N + j], exp(ts_star[(long)c * C + col])) : 1.0;
    double cf;
    gibbs_pair(x[(long)c * N + j], xstar[(long)c * G + col / ns], ell[(long)c * N + j], ls, sij, v, cf);
  }
  KXT[(long)cc * kx_stride + (long)col * Npad + j] = v;
}

// ------------------------------------------------------------------------------------------------ quadratic forms
// Q[cc][pair][col] = sum_i kx[i][col] * (sum_j Hm[pair][i][j] kx[j][col]).  CTA = (64 columns, one pair, one subject):
// 64 x 64 DMMA tiles of Hm * KX over the whole j range, one row tile i after the other; after each row tile the
// accumulators are multiplied by kx[i][col] and folded into per-column sums, so the N x C product is never stored.
// Both operands are K-major ([row][j], j contiguous) and move through a two-stage cp.async buffer.
__global__ void __launch_bounds__(128) pred_quad_kernel(const double* __restrict__ Hm, long hm_stride, int Ni64, int Npad,
                                                        const double* __restrict__ KXT, long kx_stride, int Cpad,
                                                        double* __restrict__ Q, int T, int N) {
  __shared__ __align__(16) double As[2][QT * QLD];
  __shared__ __align__(16) double Bs[2][QT * QLD];
  __shared__ double red[2][QT];
  const int cc = blockIdx.z, pair = blockIdx.y, col0 = blockIdx.x * QT;
  const double* H = Hm + (long)cc * hm_stride + (long)pair * Ni64 * Npad;
  const double* KX = KXT + (long)cc * kx_stride;
  // The quarter of the tile a warp owns is rotated per CTA: warps are pinned to SM sub-partitions (one DMMA pipe each), and
  // the quarters do unequal work when the last row tile is mostly padding.
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = ((tid >> 5) + blockIdx.x + blockIdx.y) & 3;
  const int m0 = (warp >> 1) * 32, n0 = (warp & 1) * 32;
  const int lr = lane >> 2, lk = lane & 3;
  const int nK = Npad / QK, total = (Ni64 / QT) * nK;

  auto load = [&](int q, int buf) {
    const int it = q / nK, ch = q - it * nK;
#pragma unroll
    for (int p = tid; p < QT * 8; p += 128) {
      const int row = p >> 3, seg = (p & 7) * 2;
      cp_async16(&As[buf][row * QLD + seg], H + (long)(it * QT + row) * Npad + ch * QK + seg);
      cp_async16(&Bs[buf][row * QLD + seg], KX + (long)(col0 + row) * Npad + ch * QK + seg);
    }
    cp_async_commit();
  };

  double acc[4][4][2], cs[4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
#pragma unroll
  for (int b = 0; b < 4; ++b) cs[b][0] = cs[b][1] = 0.0;

  load(0, 0);
  for (int q = 0; q < total; ++q) {
    const int buf = q & 1;
    if (q + 1 < total) {
      load(q + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* SA = &As[buf][(m0 + lr) * QLD + lk];
    const double* SB = &Bs[buf][(n0 + lr) * QLD + lk];
    // a warp whose 32 rows are all padding (N = 100: rows 96..127 of the second row tile) has nothing to add
    // 8-row groups of this warp that are all padding (N = 100: rows 104..127 of the second row tile) and k-steps of the last
    // chunk that only meet zero padding are skipped
    const int ksteps = min(QK / 4, (N - (q % nK) * QK + 3) / 4);
    const int na = min(4, max(0, (N - ((q / nK) * QT + m0) + 7) / 8));
    if (na > 0)
#pragma unroll
    for (int ks = 0; ks < QK / 4; ++ks) {
      if (ks >= ksteps) break;
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = SA[i * 8 * QLD + ks * 4];
        b[i] = SB[i * 8 * QLD + ks * 4];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < na)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();   // this buffer is refilled by the load issued at the top of the next iteration
    if ((q + 1) % nK == 0) {
      const int it = q / nK;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int row = it * QT + m0 + 8 * a + lr;
          const long col = col0 + n0 + 8 * b + 2 * lk;
          if (row < N) {   // padded rows of Hm are zero; their kx entries do not exist
            cs[b][0] += acc[a][b][0] * KX[col * Npad + row];
            cs[b][1] += acc[a][b][1] * KX[(col + 1) * Npad + row];
          }
          acc[a][b][0] = acc[a][b][1] = 0.0;
        }
    }
  }
  // fold the 8 row groups of the warp (lanes with equal lane & 3), then the two row halves of the tile
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = cs[b][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lr == 0) red[warp >> 1][n0 + 8 * b + 2 * lk + e] = v;
    }
  __syncthreads();
  if (tid < QT) Q[((long)cc * T + pair) * Cpad + col0 + tid] = red[0][tid] + red[1][tid];
}

// ------------------------------------------------------------------------------------------------ moments
// One warp per column (grid point, sample): q = U^T kx, then lanes m' < M form mu_f[m'] and sigma2_y[m'].
__global__ void __launch_bounds__(128) pred_finish_kernel(const double* __restrict__ KXT, long kx_stride, int Npad,
                                                          const double* __restrict__ U, const double* __restrict__ Q,
                                                          int Cpad, const double* __restrict__ uL_star,
                                                          const double* __restrict__ s2v, int N, int M, int MT, int T,
                                                          int C, int c0, int raw_factor, double* __restrict__ mu_f,
                                                          double* __restrict__ s2y) {
  const int col = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int cc = blockIdx.y, c = c0 + cc;
  if (col >= C) return;
  const double* kx = KXT + (long)cc * kx_stride + (long)col * Npad;
  const double* Uc = U + (long)c * N * MT;
  double q[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) q[k] = 0.0;
  for (int j = lane; j < N; j += 32) {
    const double kv = kx[j];
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < M) q[k] += kv * Uc[(long)j * MT + k];
  }
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if (k < M) q[k] = warp_sum(q[k]);
  if (lane >= M) return;
  const int mp = lane;
  const double* us = uL_star + ((long)c * C + col) * T + mp * (mp + 1) / 2;
  const double* Qc = Q + (long)cc * T * Cpad + col;
  double l[16];
#pragma unroll
  for (int k = 0; k < 16; ++k)   // utils.py:10-22; the history variant uses the sampled vector as the factor itself (:1311)
    l[k] = (k < mp) ? us[k] : (k == mp ? (raw_factor ? us[k] : exp(us[k])) : 0.0);
  double mu = 0.0, ll = 0.0, B = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (k > mp) continue;
    mu += l[k] * q[k];
    ll += l[k] * l[k];
    double r = 0.0;
    for (int k2 = 0; k2 <= mp; ++k2) {
      const int lo = k < k2 ? k : k2, hi = k < k2 ? k2 : k;
      r += l[k2] * Qc[(long)(hi * (hi + 1) / 2 + lo) * Cpad];
    }
    B += l[k] * r;
  }
  double v = (__dmul_rn(__dadd_rn(kJitter, 1.0), ll) - B) + s2v[c];       // prediction.py:1157-1161
  if (v <= 0.0) v = 1e-6;                                                  // :1163
  mu_f[((long)c * C + col) * M + mp] = mu;
  s2y[((long)c * C + col) * M + mp] = v;
}

// ------------------------------------------------------------------------------------------------ separable / stationary
// The M inverses S_m^-1 = (lam_m Kx + sigma2 I)^-1 of a subject out of the engine's padded layout into the table layout of
// pred_quad_kernel ([M][Ni64][Npad], zero padded).
__global__ void __launch_bounds__(128) pred_copy_inv_kernel(const double* __restrict__ A, long strideA, int ld, int N,
                                                            int M, int Ni64, int Npad, int c0,
                                                            double* __restrict__ Hm, long hm_stride) {
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int i = blockIdx.y / M, m = blockIdx.y % M;
  const int cc = blockIdx.z, c = c0 + cc;
  if (j >= Npad) return;
  const double v = (i < N && j < N) ? A[((long)c * M + m) * strideA + (long)i * ld + j] : 0.0;
  Hm[(long)cc * hm_stride + ((long)m * Ni64 + i) * Npad + j] = v;
}

// One warp per column: q1[m] = kx . alpha_m, q2[m] = kx^T S_m^-1 kx (from pred_quad_kernel); lanes m' < M then form
//   mu_f[m'] = sum_m lam_m V[m'][m] q1[m],   quad[m'] = sum_m (lam_m V[m'][m])^2 q2[m]
// (k_f = B (x) k_x, Sigma^-1 = sum_m v_m v_m^T (x) S_m^-1: prediction.py:95-110, 241-262, 1591-1596 in block form).
__global__ void __launch_bounds__(128) pred_finish_sep_kernel(const double* __restrict__ KXT, long kx_stride, int Npad,
                                                              const double* __restrict__ alpha,
                                                              const double* __restrict__ lam, const double* __restrict__ Vec,
                                                              const double* __restrict__ Q, int Cpad, int N, int M, int C,
                                                              int c0, double* __restrict__ mu_f, double* __restrict__ quad) {
  const int col = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int cc = blockIdx.y, c = c0 + cc;
  if (col >= C) return;
  const double* kx = KXT + (long)cc * kx_stride + (long)col * Npad;
  const double* al = alpha + (long)c * M * N;
  double q1[16];
#pragma unroll
  for (int m = 0; m < 16; ++m) q1[m] = 0.0;
  for (int j = lane; j < N; j += 32) {
    const double kv = kx[j];
#pragma unroll
    for (int m = 0; m < 16; ++m)
      if (m < M) q1[m] += kv * al[(long)m * N + j];
  }
#pragma unroll
  for (int m = 0; m < 16; ++m)
    if (m < M) q1[m] = warp_sum(q1[m]);
  if (lane >= M) return;
  const int mp = lane;
  double mu = 0.0, qd = 0.0;
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    if (m >= M) continue;
    const double f = lam[(long)c * M + m] * Vec[(long)c * M * M + mp * M + m];
    mu += f * q1[m];
    qd += f * f * Q[((long)cc * M + m) * Cpad + col];
  }
  mu_f[((long)c * C + col) * M + mp] = mu;
  quad[((long)c * C + col) * M + mp] = qd;
}

}  // namespace

// =================================================================================================== host side
#define NMGP_LAUNCH_CHECK()                 \
  do {                                      \
    NMGP_CUDA_TRY(cudaGetLastError());      \
    if (launches) ++*launches;              \
  } while (0)

size_t predict_prior_scratch_per_subject(int N, int G) { (void)N; (void)G; return 1; }   // w lives in shared memory

int launch_predict_prior(const double* x, const double* Lp, const double* Z, int cs, int N, int nv, const double* xstar,
                         int G, double alpha, double beta, double mu, double* scratch, double* mean_out, double* s2_out,
                         cudaStream_t st, long* launches) {
  (void)scratch;
  if (cs <= 0 || G <= 0) return 0;
  int wpb = 4;                                                 // warps per CTA, limited by their w vectors in shared memory
  while (wpb > 1 && (size_t)wpb * N * sizeof(double) > 200u * 1024u) wpb >>= 1;
  const size_t smem = (size_t)wpb * N * sizeof(double);
  if (smem > 227u * 1024u) { set_last_error("predict: N too large for the prior-conditional kernel"); return -1; }
  if (smem > 48u * 1024u) NMGP_SMEM_ATTR_PER_DEVICE(pred_prior_kernel, smem);
  dim3 grid((G + wpb - 1) / wpb, cs);
  pred_prior_kernel<<<grid, 32 * wpb, smem, st>>>(x, Lp, Z, N, nv, xstar, G, alpha * alpha, beta, mu, mean_out, s2_out);
  NMGP_LAUNCH_CHECK();
  return 0;
}

namespace {
struct PredLayout {
  int Ni64, Npad, Cpad, T;
  size_t hm, kx, q;   // doubles per subject
  size_t per_subject() const { return hm + kx + q; }
};
PredLayout pred_layout(int N, int M, long C) {
  PredLayout p;
  p.T = tril_size(M);
  p.Ni64 = (int)round_up(N, QT);
  p.Npad = (int)round_up(N, QK);
  p.Cpad = (int)round_up(C, QT);
  p.hm = (size_t)p.T * p.Ni64 * p.Npad;
  p.kx = (size_t)p.Cpad * p.Npad;
  p.q = (size_t)p.T * p.Cpad;
  return p;
}
}  // namespace

size_t predict_scratch_per_subject(int N, int M, long C) { return pred_layout(N, M, C).per_subject(); }

int predict_moments_chunk(int cs, int N, int M, const double* x, const double* Y, const Scratch& w, const BlockBatch& b,
                          const double* xstar, const double* tl_star, const double* uL_star, int G, int ns,
                          int raw_factor, double* scratch, size_t scratch_doubles, double* mu_f, double* s2y,
                          cudaStream_t st, long* launches) {
  if (cs <= 0 || G <= 0 || ns <= 0) return 0;
  const int n = N * M, MT = padded_M(M);
  const long C = (long)G * ns;
  if (C > 0x7fffffffL / 2) { set_last_error("predict: too many (grid point, sample) columns"); return -1; }
  const PredLayout pl = pred_layout(N, M, C);
  if ((long)pl.Ni64 * M > 65535) { set_last_error("predict: N * M too large for the table kernel's grid"); return -1; }
  long SB = (long)(scratch_doubles / pl.per_subject());
  if (SB < 1) { set_last_error("predict: scratch smaller than one subject"); return -3; }
  if (SB > cs) SB = cs;
  if (SB > 65535) SB = 65535;
  // alpha = Sigma^-1 y and u_j = L_j^T alpha_j for the whole chunk
  NMGP_TRY(launch_symv(b, n, Y, w.alpha, cs, st, launches));
  {
    const long total = (long)cs * N * MT;
    pred_u_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w.alpha, w.Lst, N, M, MT, total, w.Ua);
    NMGP_LAUNCH_CHECK();
  }
  for (int c0 = 0; c0 < cs; c0 += (int)SB) {
    const int sb = cs - c0 < SB ? cs - c0 : (int)SB;
    double* Hm = scratch;
    double* KXT = Hm + (size_t)sb * pl.hm;
    double* Qb = KXT + (size_t)sb * pl.kx;
    dim3 gh((pl.Npad + 127) / 128, pl.Ni64 * M, sb);
    pred_h_kernel<<<gh, 128, 0, st>>>(b.A, b.strideA(), b.nP, w.Lst, N, M, MT, pl.Ni64, pl.Npad, c0, Hm, (long)pl.hm);
    NMGP_LAUNCH_CHECK();
    dim3 gk(pl.Cpad, (pl.Npad + 127) / 128, sb);
    pred_kx_kernel<<<gk, 128, 0, st>>>(x, w.ell, nullptr, xstar, tl_star, nullptr, N, pl.Npad, G, ns, (int)C, c0, KXT,
                                       (long)pl.kx);
    NMGP_LAUNCH_CHECK();
    dim3 gq(pl.Cpad / QT, pl.T, sb);
    pred_quad_kernel<<<gq, 128, 0, st>>>(Hm, (long)pl.hm, pl.Ni64, pl.Npad, KXT, (long)pl.kx, pl.Cpad, Qb, pl.T, N);
    NMGP_LAUNCH_CHECK();
    dim3 gf((unsigned)((C + 3) / 4), sb);
    pred_finish_kernel<<<gf, 128, 0, st>>>(KXT, (long)pl.kx, pl.Npad, w.Ua, Qb, pl.Cpad, uL_star, w.s2, N, M, MT, pl.T,
                                           (int)C, c0, raw_factor, mu_f, s2y);
    NMGP_LAUNCH_CHECK();
  }
  return 0;
}

int predict_moments_sep_chunk(int cs, int N, int M, const double* x, const Scratch& w, const BlockBatch& b,
                              const double* xstar, const double* tl_star, const double* ts_star, int G, int ns,
                              double* scratch, size_t scratch_doubles, double* mu_f, double* quad, cudaStream_t st,
                              long* launches) {
  if (cs <= 0 || G <= 0 || ns <= 0) return 0;
  const long C = (long)G * ns;
  if (C > 0x7fffffffL / 2) { set_last_error("predict: too many (grid point, sample) columns"); return -1; }
  PredLayout pl = pred_layout(N, M, C);
  pl.T = M;                                  // M tables (one inverse per eigen-direction of B) instead of M(M+1)/2
  pl.hm = (size_t)M * pl.Ni64 * pl.Npad;
  pl.q = (size_t)M * pl.Cpad;
  if ((long)pl.Ni64 * M > 65535) { set_last_error("predict: N * M too large for the table kernel's grid"); return -1; }
  long SB = (long)(scratch_doubles / pl.per_subject());
  if (SB < 1) { set_last_error("predict: scratch smaller than one subject"); return -3; }
  if (SB > cs) SB = cs;
  if (SB > 65535) SB = 65535;
  // alpha_m = S_m^-1 (Y v_m) for the whole chunk (w.yv holds the rotated observations)
  NMGP_TRY(launch_symv(b, N, w.yv, w.alpha, cs * M, st, launches));
  for (int c0 = 0; c0 < cs; c0 += (int)SB) {
    const int sb = cs - c0 < SB ? cs - c0 : (int)SB;
    double* Hm = scratch;
    double* KXT = Hm + (size_t)sb * pl.hm;
    double* Qb = KXT + (size_t)sb * pl.kx;
    dim3 gh((pl.Npad + 127) / 128, pl.Ni64 * M, sb);
    pred_copy_inv_kernel<<<gh, 128, 0, st>>>(b.A, b.strideA(), b.nP, N, M, pl.Ni64, pl.Npad, c0, Hm, (long)pl.hm);
    NMGP_LAUNCH_CHECK();
    dim3 gk(pl.Cpad, (pl.Npad + 127) / 128, sb);
    pred_kx_kernel<<<gk, 128, 0, st>>>(x, w.ell, w.sig, xstar, tl_star, ts_star, N, pl.Npad, G, ns, (int)C, c0, KXT,
                                       (long)pl.kx);
    NMGP_LAUNCH_CHECK();
    dim3 gq(pl.Cpad / QT, M, sb);
    pred_quad_kernel<<<gq, 128, 0, st>>>(Hm, (long)pl.hm, pl.Ni64, pl.Npad, KXT, (long)pl.kx, pl.Cpad, Qb, M, N);
    NMGP_LAUNCH_CHECK();
    dim3 gf((unsigned)((C + 3) / 4), sb);
    pred_finish_sep_kernel<<<gf, 128, 0, st>>>(KXT, (long)pl.kx, pl.Npad, w.alpha, w.lam, w.Vec, Qb, pl.Cpad, N, M, (int)C,
                                               c0, mu_f, quad);
    NMGP_LAUNCH_CHECK();
  }
  return 0;
}

size_t predict_sep_scratch_per_subject(int N, int M, long C) {
  PredLayout pl = pred_layout(N, M, C);
  return (size_t)M * pl.Ni64 * pl.Npad + pl.kx + (size_t)M * pl.Cpad;
}

}  // namespace nmgp
