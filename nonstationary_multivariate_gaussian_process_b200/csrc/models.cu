// Model kernels of the NMGP hot path (FP64, sm_100a): everything between the flat parameter vector and the
// factorisation engine, and between the inverse and the gradient.  All kernels are batched over subjects.
//
// Internal ordering of the nonseparable covariance is TIME-major (row (i,m) -> i*M+m) -- a symmetric permutation
// of the reference's output-major ordering (Utility/logpos.py:347-348) that leaves log det, the quadratic form
// and (after un-permuting) the gradient unchanged, and makes the M x M cross-output blocks contiguous.
#include "models.cuh"
#include "jacobi.cuh"

#include <cmath>
#include <cstdlib>

namespace nmgp {

namespace {

constexpr int NB = kNB;

// ------------------------------------------------------------------------------------------ unit kernels
__global__ void rbf_cov_kernel(const double* __restrict__ x1, int N1, const double* __restrict__ x2, int N2,
                               double alpha2, double beta, int self, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= N2 || i >= N1) return;
  const double a = x1[i] / beta, b = x2[j] / beta;                 // kernels.py:38-39
  const double d = ref_sqdist(a, b);
  double v = __dmul_rn(exp(__dmul_rn(-0.5, d)), alpha2);            // exp(-0.5*dist)*alpha**2, kernels.py:42
  if (self && i == j) v = __dadd_rn(kJitter, v);                    // jitter identity is the accumulator, :35
  out[(long)i * N2 + j] = v;
}

__global__ void gibbs_cov_kernel(const double* __restrict__ x1, const double* __restrict__ s1,
                                 const double* __restrict__ l1, int N1, const double* __restrict__ x2,
                                 const double* __restrict__ s2, const double* __restrict__ l2, int N2, int self,
                                 double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= N2 || i >= N1) return;
  const double sij = (s1 ? s1[i] : 1.0) * (s2 ? s2[j] : 1.0);
  double k0, cf;
  gibbs_pair(x1[i], x2[j], l1[i], l2[j], sij, k0, cf);
  if (self && i == j) k0 = __dadd_rn(kJitter, k0);
  out[(long)i * N2 + j] = k0;
}

// prior covariance into the padded block layout (full symmetric, identity padding)
__global__ void prior_cov_blocks_kernel(const double* __restrict__ x, int N, double alpha2, double beta, double* A,
                                        long strideA, int ld) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  const int c = blockIdx.z;
  if (q >= ld) return;
  double v;
  if (p < N && q < N) {
    const double* xs = x + (long)c * N;
    const double a = xs[p] / beta, b = xs[q] / beta;
    v = __dmul_rn(exp(__dmul_rn(-0.5, ref_sqdist(a, b))), alpha2);
    if (p == q) v = __dadd_rn(kJitter, v);
  } else {
    v = (p == q) ? 1.0 : 0.0;
  }
  A[(long)c * strideA + (long)p * ld + q] = v;
}

// compact lower Cholesky factor [c][N][N] (zero strict upper) and half log-determinant out of the padded layout
__global__ void extract_factor_kernel(const double* __restrict__ A, long strideA, int ld, int N,
                                      const double* __restrict__ logdet, double* __restrict__ Lp,
                                      double* __restrict__ hld) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  const int c = blockIdx.z;
  if (p == 0 && q == 0) hld[c] = 0.5 * logdet[c];
  if (q >= N) return;
  Lp[((long)c * N + p) * N + q] = (q <= p) ? A[(long)c * strideA + (long)p * ld + q] : 0.0;
}

// Triangular solves with the cached prior factor, nv right-hand sides per subject, by blocked (64-row) right-looking
// substitution -- the arithmetic MultivariateNormal.log_prob and its autograd backward do (logpos.py:274,279,358,365):
//   TRANS=0:  L   X = rhs   (forward, blocks ascending)        TRANS=1:  L^T X = rhs   (backward, blocks descending)
// Two kernels per block step, both batched over subjects, so a single large subject still spreads over the GPU:
//   prior_diag_solve : X_k <- L_kk^-1 X_k  (or L_kk^-T)   one warp per right-hand side, L_kk in shared memory
//   prior_update     : X_i <- X_i - L(i,k) X_k  for all i > k   (or  X_i - L(k,i)^T X_k, i < k)   one CTA per row block
constexpr int PBS = 64;   // block rows
constexpr int PVC = 24;   // right-hand sides per CTA (static shared memory stays under 48 KB)

template <int TRANS>
__global__ void __launch_bounds__(128) prior_diag_solve_kernel(const double* __restrict__ Lp, double* X, int N, int nv,
                                                               int k0) {
  __shared__ double Ld[PBS][PBS + 1];
  __shared__ double Zb[PBS][PVC + 1];
  const int c = blockIdx.x;
  const int v0 = blockIdx.y * PVC;
  const int vc = min(PVC, nv - v0);
  const int rows = min(PBS, N - k0);
  const double* L = Lp + (long)c * N * N;
  double* Xc = X + (long)c * N * nv;
  for (int idx = threadIdx.x; idx < rows * rows; idx += 128) {
    const int r = idx / rows, cc = idx % rows;
    Ld[r][cc] = L[(long)(k0 + r) * N + k0 + cc];
  }
  for (int idx = threadIdx.x; idx < rows * vc; idx += 128) {
    const int r = idx / vc, v = idx % vc;
    Zb[r][v] = Xc[(long)(k0 + r) * nv + v0 + v];
  }
  __syncthreads();
  // one WARP per right-hand side (the four warps take them round-robin), lanes over the rows: 64 dependent steps of one
  // shared-memory round trip each instead of one thread's 2048 dependent FMAs (23 us -> ~2 us for a single right-hand side,
  // which is what the separable model's priors have)
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int v = warp; v < vc; v += 4) {
      if (TRANS) {     // L^T x = z, descending: x_r final, then z_i -= L[r][i] x_r for i < r   (row r of L_kk)
        for (int r = rows - 1; r >= 0; --r) {
          const double xr = Zb[r][v] / Ld[r][r];
          __syncwarp();
          for (int i = lane; i < r; i += 32) Zb[i][v] -= Ld[r][i] * xr;
          if (lane == 0) Zb[r][v] = xr;
          __syncwarp();
        }
      } else {         // L x = z, ascending: x_r final, then z_i -= L[i][r] x_r for i > r   (column r of L_kk)
        for (int r = 0; r < rows; ++r) {
          const double xr = Zb[r][v] / Ld[r][r];
          __syncwarp();
          for (int i = r + 1 + lane; i < rows; i += 32) Zb[i][v] -= Ld[i][r] * xr;
          if (lane == 0) Zb[r][v] = xr;
          __syncwarp();
        }
      }
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < rows * vc; idx += 128) {
    const int r = idx / vc, v = idx % vc;
    Xc[(long)(k0 + r) * nv + v0 + v] = Zb[r][v];
  }
}

template <int TRANS>
__global__ void __launch_bounds__(256) prior_update_kernel(const double* __restrict__ Lp, double* X, int N, int nv,
                                                           int k0) {
  __shared__ double Lt[PBS][PBS + 1];   // Lt[r][k]: coefficient of X_k[k] in row r of the target block
  __shared__ double Zk[PBS][PVC + 1];
  const int c = blockIdx.x;
  const int v0 = blockIdx.z * PVC;
  const int vc = min(PVC, nv - v0);
  const int krows = min(PBS, N - k0);
  // target block i: forward -> blocks after k; backward -> blocks before k
  const int i0 = TRANS ? blockIdx.y * PBS : k0 + PBS + blockIdx.y * PBS;
  const int irows = min(PBS, (TRANS ? k0 : N) - i0);
  const double* L = Lp + (long)c * N * N;
  double* Xc = X + (long)c * N * nv;
  for (int idx = threadIdx.x; idx < PBS * PBS; idx += 256) {
    double v = 0.0;
    if (TRANS) {   // L(k,i)^T : Lt[r][k] = L[k0+k][i0+r]; read with r fastest (coalesced)
      const int kk = idx / PBS, r = idx % PBS;
      if (kk < krows && r < irows) v = L[(long)(k0 + kk) * N + i0 + r];
      Lt[r][kk] = v;
    } else {       // L(i,k)   : Lt[r][k] = L[i0+r][k0+k]
      const int r = idx / PBS, kk = idx % PBS;
      if (kk < krows && r < irows) v = L[(long)(i0 + r) * N + k0 + kk];
      Lt[r][kk] = v;
    }
  }
  for (int idx = threadIdx.x; idx < krows * vc; idx += 256) {
    const int r = idx / vc, v = idx % vc;
    Zk[r][v] = Xc[(long)(k0 + r) * nv + v0 + v];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < irows * vc; idx += 256) {
    const int r = idx / vc, v = idx % vc;
    double acc = 0.0;
    for (int kk = 0; kk < krows; ++kk) acc += Lt[r][kk] * Zk[kk][v];
    Xc[(long)(i0 + r) * nv + v0 + v] -= acc;
  }
}

// Many small systems (the batched regime: N ~ 100, thousands of subjects): ONE WARP solves all right-hand sides of one
// subject, one right-hand side per lane, 32 rows of the solution in registers at a time.  The factor is staged through a
// 32 x 32 shared-memory block per warp and read back as uniform-address (broadcast) LDS.128 -- no bank conflicts, no
// shuffles, no CTA barriers; the already-solved part of the solution is re-read from the output array (written by the
// same lane, coalesced across lanes).
//   TRANS=0:  L x = b, block rows ascending:   acc[r] -= L[r][k] x[k]  (k < block),   staged TRANSPOSED ([k][r])
//   TRANS=1:  L^T x = b, block rows descending: acc[k] -= L[r][k] x[r]  (r > block),   staged as stored ([r][k])
constexpr int PWB = 4;   // warps (independent problems) per CTA

template <int TRANS>
__global__ void __launch_bounds__(32 * PWB) prior_solve_warp_kernel(const double* __restrict__ Lp,
                                                                    const double* __restrict__ rhs, double* X, int cs,
                                                                    int N, int nv, int nvc) {
  __shared__ __align__(16) double Ls_all[PWB][32 * 34];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int prob = blockIdx.x * PWB + warp;
  if (prob >= cs * nvc) return;          // whole warps leave together; only __syncwarp below
  const int c = prob / nvc, v = (prob % nvc) * 32 + lane;
  const bool vok = v < nv;
  double* Ls = Ls_all[warp];             // [32][34]
  const double* L = Lp + (long)c * N * N;
  const double* R = rhs + (long)c * N * nv;
  double* Xc = X + (long)c * N * nv;
  const int nblk = (N + 31) / 32;
  for (int bi = 0; bi < nblk; ++bi) {
    const int b = TRANS ? nblk - 1 - bi : bi;
    const int r0 = b * 32;
    double acc[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) acc[r] = (vok && r0 + r < N) ? R[(long)(r0 + r) * nv + v] : 0.0;
    // ---- off-diagonal blocks already solved
    for (int oi = 0; oi < bi; ++oi) {
      const int o = TRANS ? nblk - 1 - oi : oi;
      const int k0 = o * 32;
      __syncwarp();
      if (TRANS) {   // rows k0.. (the solved x[r]), columns r0..r0+31 of L:  Ls[rr][kk] = L[k0+rr][r0+kk]
#pragma unroll 16
        for (int rr = 0; rr < 32; ++rr)
          Ls[rr * 34 + lane] = (k0 + rr < N) ? L[(long)(k0 + rr) * N + r0 + lane] : 0.0;   // r0 + lane < k0 <= N
      } else {       // rows r0..r0+31, columns k0..k0+31 (solved x[k]) of L, transposed:  Ls[kk][rr] = L[r0+rr][k0+kk]
#pragma unroll 16
        for (int rr = 0; rr < 32; ++rr)
          Ls[lane * 34 + rr] = (r0 + rr < N) ? L[(long)(r0 + rr) * N + k0 + lane] : 0.0;
      }
      __syncwarp();
#pragma unroll 8
      for (int kk = 0; kk < 32; ++kk) {
        const double xk = (vok && k0 + kk < N) ? Xc[(long)(k0 + kk) * nv + v] : 0.0;
        const double* row = Ls + kk * 34;
#pragma unroll
        for (int r = 0; r < 32; r += 2) {
          const double2 l2 = *reinterpret_cast<const double2*>(row + r);
          acc[r] -= l2.x * xk;
          acc[r + 1] -= l2.y * xk;
        }
      }
    }
    // ---- diagonal block
    __syncwarp();
    if (TRANS) {
#pragma unroll 16
      for (int rr = 0; rr < 32; ++rr)
        Ls[rr * 34 + lane] = (r0 + rr < N && r0 + lane < N) ? L[(long)(r0 + rr) * N + r0 + lane] : (rr == lane ? 1.0 : 0.0);
    } else {
#pragma unroll 16
      for (int rr = 0; rr < 32; ++rr)
        Ls[lane * 34 + rr] = (r0 + rr < N && r0 + lane < N) ? L[(long)(r0 + rr) * N + r0 + lane] : (rr == lane ? 1.0 : 0.0);
    }
    __syncwarp();
    if (TRANS) {   // Ls[r][k] = L_bb[r][k]:  x[r] final (descending) -> acc[k] -= L[r][k] x[r], k < r
#pragma unroll
      for (int r = 31; r >= 0; --r) {
        acc[r] = acc[r] / Ls[r * 34 + r];
        const double xr = acc[r];
#pragma unroll
        for (int k = 0; k < r; ++k) acc[k] -= Ls[r * 34 + k] * xr;
      }
    } else {       // Ls[k][r] = L_bb[r][k]:  x[k] final (ascending) -> acc[r] -= L[r][k] x[k], r > k
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        acc[k] = acc[k] / Ls[k * 34 + k];
        const double xk = acc[k];
#pragma unroll
        for (int r = k + 1; r < 32; ++r) acc[r] -= Ls[k * 34 + r] * xk;
      }
    }
    if (vok) {
#pragma unroll
      for (int r = 0; r < 32; ++r)
        if (r0 + r < N) Xc[(long)(r0 + r) * nv + v] = acc[r];
    }
  }
}

// ------------------------------------------------------------------------------------------ shared by models
// Gibbs kernel matrix Kx (with jitter) and CK = c_ij * K0_ij for every subject of the chunk.
__global__ void kx_kernel(const double* __restrict__ x, const double* __restrict__ ell, const double* __restrict__ sig,
                          int N, double* __restrict__ Kx, double* __restrict__ CK) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  const int c = blockIdx.z;
  if (j >= N) return;
  const double* xs = x + (long)c * N;
  const double* ls = ell + (long)c * N;
  const double sij = sig ? sig[(long)c * N + i] * sig[(long)c * N + j] : 1.0;
  double k0, cf;
  gibbs_pair(xs[i], xs[j], ls[i], ls[j], sij, k0, cf);
  const long o = ((long)c * N + i) * N + j;
  CK[o] = cf * k0;
  Kx[o] = (i == j) ? __dadd_rn(kJitter, k0) : k0;
}

// alpha[b][p] = sum_q A[b][p][q] y[b][q]   (full symmetric A in the padded layout; one warp per row)
__global__ void __launch_bounds__(256) symv_kernel(const double* __restrict__ A, long strideA, int ld, int n,
                                                   const double* __restrict__ y, double* __restrict__ alpha) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= n) return;
  const double* row = A + (long)b * strideA + (long)p * ld;
  const double* yb = y + (long)b * n;
  double s = 0.0;
  for (int q = lane; q < n; q += 32) s += row[q] * yb[q];
  s = warp_sum(s);
  if (lane == 0) alpha[(long)b * n + p] = s;
}

// ------------------------------------------------------------------------------------------ nonseparable
// pars -> ell, Lst rows, sigma2, prior residuals    (Utility/utils.py:10-74, logpos.py:337-343)
__global__ void svc_prep_kernel(const double* __restrict__ pars, int P, int N, int M, int MT, double mu0, double mu1,
                                double* __restrict__ ell, double* __restrict__ Lst, double* __restrict__ s2,
                                double* __restrict__ R0, double* __restrict__ R1) {
  const int c = blockIdx.x;
  const int T = tril_size(M);
  const double* p = pars + (long)c * P;
  if (threadIdx.x == 0) s2[c] = exp(p[P - 1]);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double tl = p[i];
    ell[(long)c * N + i] = exp(tl);
    R0[(long)c * N + i] = tl - mu0;
  }
  for (int idx = threadIdx.x; idx < N * M * MT; idx += blockDim.x) {
    const int k = idx % MT, m = (idx / MT) % M, i = idx / (MT * M);
    double v = 0.0;
    if (k <= m) {
      const double u = p[N + (long)i * T + m * (m + 1) / 2 + k];
      v = (k == m) ? exp(u) : u;
    }
    Lst[((long)c * N * M + (long)i * M + m) * MT + k] = v;
  }
  for (int idx = threadIdx.x; idx < N * T; idx += blockDim.x) R1[(long)c * N * T + idx] = p[N + idx] - mu1;
}

// Sigma = Kx[i,j] * L_i L_j^T + sigma2 I, time-major, lower block triangle of the padded layout.
// One CTA per 64x64 tile; a warp writes one tile row as 32 double2 (512 B, coalesced); the factor rows of the tile's
// columns sit k-major in shared memory so the M-term dot products read it conflict-free.  Memory-bound on the stores.
template <int M>
__global__ void __launch_bounds__(256, 5) svc_build_kernel(const double* __restrict__ Kx, const double* __restrict__ Lst,
                                                           const double* __restrict__ s2v, int N, int MT, double* A,
                                                           long strideA, int ld) {
  constexpr int KS = (NB + M - 1) / M + 1;          // time points a 64-row / 64-column range can touch
  __shared__ double Lr[NB][M];
  __shared__ __align__(16) double LcT[M][NB];
  __shared__ double Ks[KS][KS + 1];                  // the Kx entries of this tile (no global load in the store loop)
  int ti = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((long)ti * (ti + 1) / 2 > (long)blockIdx.x) --ti;
  while ((long)(ti + 1) * (ti + 2) / 2 <= (long)blockIdx.x) ++ti;
  const int tj = blockIdx.x - ti * (ti + 1) / 2;
  const int c = blockIdx.y;
  const int n = N * M;
  const int p0 = ti * NB, q0 = tj * NB;
  const int i0 = p0 / M, j0 = q0 / M;
  const double* Ls = Lst + (long)c * n * MT;
  const double* Kc = Kx + (long)c * N * N;
  for (int idx = threadIdx.x; idx < NB * M; idx += 256) {
    const int r = idx / M, k = idx % M;
    Lr[r][k] = (p0 + r < n) ? Ls[(long)(p0 + r) * MT + k] : 0.0;
    LcT[k][r] = (q0 + r < n) ? Ls[(long)(q0 + r) * MT + k] : 0.0;
  }
  for (int idx = threadIdx.x; idx < KS * KS; idx += 256) {
    const int ii = idx / KS, jj = idx % KS;
    Ks[ii][jj] = (i0 + ii < N && j0 + jj < N) ? Kc[(long)(i0 + ii) * N + j0 + jj] : 0.0;
  }
  __syncthreads();
  const double s2 = s2v[c];
  double* Ac = A + (long)c * strideA;
#pragma unroll
  for (int it = 0; it < (NB * NB / 2) / 256; ++it) {
    const int idx = threadIdx.x + it * 256;
    const int r = idx >> 5, c2 = idx & 31;
    const int p = p0 + r, q = q0 + 2 * c2;
    double d0 = 0.0, d1 = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) {
      const double lr = Lr[r][k];
      const double2 lc = *reinterpret_cast<const double2*>(&LcT[k][2 * c2]);
      d0 += lr * lc.x;
      d1 += lr * lc.y;
    }
    double2 v;
    if (p < n) {
      const double* krow = Ks[p / M - i0];
      v.x = (q < n) ? krow[q / M - j0] * d0 + (p == q ? s2 : 0.0) : 0.0;
      v.y = (q + 1 < n) ? krow[(q + 1) / M - j0] * d1 + (p == q + 1 ? s2 : 0.0) : 0.0;
    } else {
      v.x = (p == q) ? 1.0 : 0.0;
      v.y = (p == q + 1) ? 1.0 : 0.0;
    }
    *reinterpret_cast<double2*>(Ac + (long)p * ld + q) = v;
  }
}

// Reference-ordered dense covariance for the parity entry point (output-major, full symmetric, no padding).
__global__ void svc_cov_reference_order_kernel(const double* __restrict__ x, const double* __restrict__ pars, int N,
                                               int M, double* __restrict__ out) {
  const int n = N * M, T = tril_size(M), P = N + N * T + 1;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  const int c = blockIdx.z;
  if (q >= n) return;
  const double* pr = pars + (long)c * P;
  const double* xs = x + (long)c * N;
  const int m = p / N, i = p % N, m2 = q / N, j = q % N;
  double k0, cf;
  gibbs_pair(xs[i], xs[j], exp(pr[i]), exp(pr[j]), 1.0, k0, cf);
  if (i == j) k0 = __dadd_rn(kJitter, k0);
  double dot = 0.0;
  const int kmax = min(m, m2);
  for (int k = 0; k <= kmax; ++k) {
    double a = pr[N + (long)i * T + m * (m + 1) / 2 + k];
    double b = pr[N + (long)j * T + m2 * (m2 + 1) / 2 + k];
    if (k == m) a = exp(a);
    if (k == m2) b = exp(b);
    dot += a * b;
  }
  double v = k0 * dot;
  if (p == q) v += exp(pr[P - 1]);
  out[((long)c * n + p) * n + q] = v;
}

// The gradient needs  W[p][k] = sum_q Kx[i,j] G[p,q] L_j[m'][k]  and  V[p][k] = sum_q CK[i,j] G[p,q] L_j[m'][k]
// (p=(i,m), q=(j,m')) with G = -0.5 Sigma^-1 + 0.5 alpha alpha^T.  G is never formed: the Sigma^-1 part is ONE pass over
// the inverse that also produces alpha = Sigma^-1 y (so the inverse is read exactly once); the rank-one part is an
// O(N^2 M) correction (svc_u_kernel + svc_alpha_terms_kernel).
// Mapping: a lane owns ROW p and keeps its 2M+1 partial sums in registers; the warp walks q = 0..n-1 and reads
// Z[q][p0..p0+31] -- the inverse is symmetric, so the column strip a lane needs is a contiguous 256-byte row segment:
// every load is perfectly coalesced and each byte of Sigma^-1 is fetched exactly once per subject.  The factors L_j and
// the observations y_j are uniform across the warp: staged in shared memory, read as broadcasts.  Kx[i_p][j] / CK[i_p][j]
// take ~32/M distinct rows per warp and walk along j, i.e. L1-resident lines.
template <int M>
__global__ void __launch_bounds__(128) svc_contract_kernel(const double* __restrict__ A, long strideA, int ld, int N,
                                                           int MT, int JC, const double* __restrict__ Y,
                                                           const double* __restrict__ Kx, const double* __restrict__ CK,
                                                           const double* __restrict__ Lst, double* __restrict__ alpha,
                                                           double* __restrict__ Wout, double* __restrict__ Vout) {
  extern __shared__ __align__(16) double csm[];   // [JC][LS] factors, then [JC][M] observations
  constexpr int LS = M * M;
  double* Lsm = csm;
  double* ysm = csm + (size_t)JC * LS;
  const int c = blockIdx.y;
  const int n = N * M;
  const int p = blockIdx.x * 128 + threadIdx.x;
  const bool valid = p < n;
  const int pc = valid ? p : n - 1;          // clamped: out-of-range lanes read a real column and discard the sums
  const int ip = pc / M;
  const double* Zc = A + (long)c * strideA + pc;
  const double* Ls = Lst + (long)c * n * MT;
  const double* y = Y + (long)c * n;
  const double* kr = Kx + ((long)c * N + ip) * N;
  const double* cr = CK + ((long)c * N + ip) * N;
  double aacc = 0.0, w[M], v[M];
#pragma unroll
  for (int k = 0; k < M; ++k) w[k] = v[k] = 0.0;
  for (int j0 = 0; j0 < N; j0 += JC) {
    const int jn = min(JC, N - j0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < jn * LS; idx += 128) {
      const int k = idx % M, m2 = (idx / M) % M, jj = idx / LS;
      Lsm[idx] = Ls[((long)(j0 + jj) * M + m2) * MT + k];
    }
    for (int idx = threadIdx.x; idx < jn * M; idx += 128) ysm[idx] = y[(long)j0 * M + idx];
    __syncthreads();
    // software pipeline: the M loads of time block j+1 are issued before the arithmetic of block j
    double zn[M];
#pragma unroll
    for (int m2 = 0; m2 < M; ++m2) zn[m2] = Zc[(long)(j0 * M + m2) * ld];
    double kxn = kr[j0], ckn = cr[j0];
#pragma unroll 2
    for (int jj = 0; jj < jn; ++jj) {
      double z[M];
#pragma unroll
      for (int m2 = 0; m2 < M; ++m2) z[m2] = zn[m2];
      const double kx = kxn, ck = ckn;
      const int jnx = min(j0 + jj + 1, N - 1);       // clamped prefetch (the last one is discarded)
#pragma unroll
      for (int m2 = 0; m2 < M; ++m2) zn[m2] = Zc[(long)(jnx * M + m2) * ld];
      kxn = kr[jnx];
      ckn = cr[jnx];
      const double* lj = Lsm + jj * LS;
      const double* yj = ysm + jj * M;
      double t[M];
#pragma unroll
      for (int k = 0; k < M; ++k) t[k] = 0.0;
#pragma unroll
      for (int m2 = 0; m2 < M; ++m2) {
        aacc += z[m2] * yj[m2];
#pragma unroll
        for (int k = 0; k <= m2; ++k) t[k] += z[m2] * lj[m2 * M + k];   // L_j is lower triangular
      }
#pragma unroll
      for (int k = 0; k < M; ++k) {
        w[k] += kx * t[k];
        v[k] += ck * t[k];
      }
    }
  }
  if (valid) {
    alpha[(long)c * n + p] = aacc;
#pragma unroll
    for (int k = 0; k < M; ++k) {
      Wout[((long)c * n + p) * MT + k] = w[k];
      Vout[((long)c * n + p) * MT + k] = v[k];
    }
  }
}

// rank-one part of the gradient sums:  u_j = L_j^T alpha_j  (one thread per time point), then
//   Sa[i][k] = sum_j Kx[i,j] u_j[k],  Ca[i][k] = sum_j CK[i,j] u_j[k]   (one warp per i, coalesced rows of Kx / CK / u)
template <int M>
__global__ void __launch_bounds__(128) svc_u_kernel(int N, int MT, const double* __restrict__ alpha,
                                                    const double* __restrict__ Lst, double* __restrict__ U) {
  const int c = blockIdx.y;
  const int j = blockIdx.x * 128 + threadIdx.x;
  if (j >= N) return;
  const double* al = alpha + ((long)c * N + j) * M;
  const double* Lj = Lst + ((long)c * N + j) * M * MT;
  double u[M];
#pragma unroll
  for (int k = 0; k < M; ++k) u[k] = 0.0;
#pragma unroll
  for (int m2 = 0; m2 < M; ++m2) {
    const double a = al[m2];
#pragma unroll
    for (int k = 0; k <= m2; ++k) u[k] += a * Lj[m2 * MT + k];
  }
#pragma unroll
  for (int k = 0; k < M; ++k) U[((long)c * N + j) * MT + k] = u[k];
}

template <int M>
__global__ void __launch_bounds__(256) svc_alpha_terms_kernel(int N, int MT, const double* __restrict__ U,
                                                              const double* __restrict__ Kx,
                                                              const double* __restrict__ CK, double* __restrict__ Sa,
                                                              double* __restrict__ Ca) {
  const int c = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= N) return;
  const double* Uc = U + (long)c * N * MT;
  const double* kr = Kx + ((long)c * N + i) * N;
  const double* cr = CK + ((long)c * N + i) * N;
  double sa[M], ca[M];
#pragma unroll
  for (int k = 0; k < M; ++k) sa[k] = ca[k] = 0.0;
  for (int j = lane; j < N; j += 32) {
    const double kx = kr[j], ck = cr[j];
#pragma unroll
    for (int k = 0; k < M; ++k) {
      const double u = Uc[(long)j * MT + k];
      sa[k] += kx * u;
      ca[k] += ck * u;
    }
  }
#pragma unroll
  for (int k = 0; k < M; ++k) {
    sa[k] = warp_sum(sa[k]);
    ca[k] = warp_sum(ca[k]);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < M; ++k) {
      Sa[((long)c * N + i) * MT + k] = sa[k];
      Ca[((long)c * N + i) * MT + k] = ca[k];
    }
  }
}

__global__ void __launch_bounds__(256) svc_finish_kernel(
    int N, int M, int MT, int P, const double* __restrict__ pars, const double* __restrict__ Y, HyperConst h,
    const double* __restrict__ A, long strideA, int ld, const double* __restrict__ logdet,
    const int* __restrict__ info_in, const double* __restrict__ alpha, const double* __restrict__ s2v,
    const double* __restrict__ Lst, const double* __restrict__ Wout, const double* __restrict__ Vout,
    const double* __restrict__ Sa, const double* __restrict__ Ca,
    const double* __restrict__ Z0, const double* __restrict__ Z1, const double* __restrict__ G0,
    const double* __restrict__ G1, const double* __restrict__ hld0, const double* __restrict__ hld1,
    double* __restrict__ vals, double* __restrict__ grad, int* __restrict__ info) {
  __shared__ double scratch[40];
  const int c = blockIdx.x;
  const int n = N * M, T = tril_size(M);
  const double* p = pars + (long)c * P;
  const double* y = Y + (long)c * n;  // time-major y == Y flattened row-major (logpos.py:336, permuted)
  const double* al = alpha + (long)c * n;
  const double* Ac = A + (long)c * strideA;
  double q = 0.0, tr = 0.0;
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const double a = al[r];
    q += y[r] * a;
    tr += -0.5 * Ac[(long)r * ld + r] + 0.5 * a * a;
  }
  const double quad = block_sum(q, scratch);
  const double trG = block_sum(tr, scratch);
  double z0 = 0.0, z1 = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) { const double z = Z0[(long)c * N + i]; z0 += z * z; }
  for (int idx = threadIdx.x; idx < N * T; idx += blockDim.x) { const double z = Z1[(long)c * N * T + idx]; z1 += z * z; }
  const double q0 = block_sum(z0, scratch);
  const double q1 = block_sum(z1, scratch);

  const double s2 = s2v[c];
  const double ts2 = p[P - 1];
  const double log2pi = 1.8378770664093453;
  const double loglik = -0.5 * logdet[c] - 0.5 * quad;                 // distributions.py:22
  const double lp_l = -0.5 * (N * log2pi + q0) - hld0[c];              // MVN.log_prob, logpos.py:358
  const double lp_L = -0.5 * ((double)T * N * log2pi + q1) - (double)T * hld1[c];   // logpos.py:365
  const double lp_s = (-h.ig_a - 1.0) * log(s2) - h.ig_b / s2 + h.ig_alogb - h.ig_lgamma;          // distributions.py:134
  double res = loglik;
  if (h.prior) res += lp_l + lp_L + lp_s + ts2;                        // logpos.py:359-376
  const double pf = h.prior ? 1.0 : 0.0;
  if (threadIdx.x == 0) {
    double* v = vals + (long)c * 6;
    v[0] = -res; v[1] = loglik; v[2] = lp_l; v[3] = lp_L; v[4] = lp_s; v[5] = 0.0;
    info[c] = info_in[c];
  }
  if (grad == nullptr) return;
  double* g = grad + (long)c * P;
  const double* Ls = Lst + (long)c * n * MT;
  const double* Wc = Wout + (long)c * n * MT;
  const double* Vc = Vout + (long)c * n * MT;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double s = 0.0;
    for (int m = 0; m < M; ++m) {
      const double ap = al[i * M + m];
      for (int k = 0; k <= m; ++k)
        s += Ls[((long)i * M + m) * MT + k] *
             (-0.5 * Vc[((long)i * M + m) * MT + k] + 0.5 * ap * Ca[((long)c * N + i) * MT + k]);
    }
    g[i] = -(2.0 * s - pf * G0[(long)c * N + i]);
  }
  for (int idx = threadIdx.x; idx < N * T; idx += blockDim.x) {
    const int i = idx / T, t = idx % T;
    int m = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while (m * (m + 1) / 2 > t) --m;
    while ((m + 1) * (m + 2) / 2 <= t) ++m;
    const int k = t - m * (m + 1) / 2;
    double d = -Wc[((long)i * M + m) * MT + k] + al[i * M + m] * Sa[((long)c * N + i) * MT + k];
    if (k == m) d *= Ls[((long)i * M + m) * MT + m];                   // d exp(u)/du on the diagonal slots
    g[N + idx] = -(d - pf * G1[(long)c * N * T + idx]);
  }
  if (threadIdx.x == 0) g[P - 1] = -(s2 * trG + pf * ((-h.ig_a - 1.0) + h.ig_b / s2 + 1.0));
}

// ------------------------------------------------------------------------------------------ separable / stationary
// pars -> ell, sigma, sigma2, L, eig(B), rotated observations, prior residuals   (logpos.py:249-257, 423-429)
__global__ void __launch_bounds__(256) sep_prep_kernel(int model, const double* __restrict__ pars, int P, int N, int M,
                                                       const double* __restrict__ Y, double mu0, double mu1,
                                                       double* __restrict__ ell, double* __restrict__ sig,
                                                       double* __restrict__ s2, double* __restrict__ Lmat,
                                                       double* __restrict__ lam, double* __restrict__ Vec,
                                                       double* __restrict__ yv, double* __restrict__ R0,
                                                       double* __restrict__ R1) {
  __shared__ double Ls[16 * 17], Bs[16 * 17], Vs[16 * 17];
  const int c = blockIdx.x;
  const int T = tril_size(M);
  const double* p = pars + (long)c * P;
  const bool stat = (model == 0);
  const double* uL = stat ? p + 2 : p + 2 * N;
  if (threadIdx.x == 0) s2[c] = exp(p[P - 1]);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double tl = stat ? p[0] : p[i];
    const double tsg = stat ? p[1] : p[N + i];
    ell[(long)c * N + i] = exp(tl);
    sig[(long)c * N + i] = exp(tsg);
    if (!stat) {
      R0[(long)c * N + i] = tl - mu0;
      R1[(long)c * N + i] = tsg - mu1;
    }
  }
  for (int idx = threadIdx.x; idx < M * M; idx += blockDim.x) {
    const int m = idx / M, k = idx % M;
    double v = 0.0;
    if (k <= m) {
      const double u = uL[m * (m + 1) / 2 + k];
      v = (k == m) ? exp(u) : u;
    }
    Ls[m * 17 + k] = v;
    Lmat[(long)c * M * M + idx] = v;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < M * M; idx += blockDim.x) {
    const int a = idx / M, b = idx % M;
    double s = 0.0;
    for (int k = 0; k < M; ++k) s += Ls[a * 17 + k] * Ls[b * 17 + k];
    Bs[a * 17 + b] = s;
  }
  __syncthreads();
  if (threadIdx.x < 32) jacobi_eig_warp(Bs, Vs, M, 17);
  __syncthreads();
  for (int idx = threadIdx.x; idx < M * M; idx += blockDim.x) Vec[(long)c * M * M + idx] = Vs[(idx / M) * 17 + idx % M];
  for (int m = threadIdx.x; m < M; m += blockDim.x) lam[(long)c * M + m] = Bs[m * 17 + m];
  // rotated observations  yv[m][i] = sum_m' Y[i][m'] V[m'][m]
  for (int idx = threadIdx.x; idx < M * N; idx += blockDim.x) {
    const int m = idx / N, i = idx % N;
    double s = 0.0;
    for (int k = 0; k < M; ++k) s += Y[((long)c * N + i) * M + k] * Vs[k * 17 + m];
    yv[((long)c * M + m) * N + i] = s;
  }
}

// S_m = lam_m Kx + sigma2 I in the padded layout (full symmetric lower+upper of the lower block triangle)
__global__ void sep_build_kernel(const double* __restrict__ Kx, const double* __restrict__ lam,
                                 const double* __restrict__ s2v, int N, int M, double* A, long strideA, int ld) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  const int b = blockIdx.z;  // c*M + m
  if (q >= ld) return;
  const int c = b / M;
  double v;
  if (p < N && q < N) {
    v = lam[b] * Kx[((long)c * N + p) * N + q];
    if (p == q) v += s2v[c];
  } else {
    v = (p == q) ? 1.0 : 0.0;
  }
  A[(long)b * strideA + (long)p * ld + q] = v;
}

// One pass over the M inverses S_m^-1 of a subject, one warp per row i:
//   gl[i] = sum_j GK_ij CK_ij,  gs[i] = sum_j GK_ij K0_ij,  rowSK[i][m] = sum_j S_m^-1[i,j] Kx[i,j],
//   KA[i][m] = sum_j Kx[i,j] alpha_m[j],     GK = sum_m lam_m (-0.5 S_m^-1 + 0.5 alpha_m alpha_m^T)
__global__ void __launch_bounds__(256) sep_contract_kernel(const double* __restrict__ A, long strideA, int ld, int N,
                                                           int M, const double* __restrict__ alpha,
                                                           const double* __restrict__ lam,
                                                           const double* __restrict__ Kx, const double* __restrict__ CK,
                                                           const double* __restrict__ sig,
                                                           double* __restrict__ gl, double* __restrict__ gs,
                                                           double* __restrict__ rowSK, double* __restrict__ KA) {
  const int c = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= N) return;
  const double* kr = Kx + ((long)c * N + i) * N;
  const double* cr = CK + ((long)c * N + i) * N;
  const double sii = sig[(long)c * N + i] * sig[(long)c * N + i];
  double agl = 0.0, ags = 0.0;
  for (int m = 0; m < M; ++m) {
    const double* row = A + ((long)c * M + m) * strideA + (long)i * ld;
    const double* al = alpha + ((long)c * M + m) * N;
    const double lm = lam[(long)c * M + m];
    const double ai = al[i];
    double sk = 0.0, ka = 0.0;
    for (int j = lane; j < N; j += 32) {
      const double sinv = row[j];
      const double kx = kr[j];
      const double k0 = (j == i) ? sii : kx;            // K0_ii = sigma_i^2 exactly (2B/A == 1, exp(0) == 1)
      const double gk = lm * (-0.5 * sinv + 0.5 * ai * al[j]);
      agl += gk * cr[j];
      ags += gk * k0;
      sk += sinv * kx;
      ka += kx * al[j];
    }
    sk = warp_sum(sk);
    ka = warp_sum(ka);
    if (lane == 0) {
      rowSK[((long)c * N + i) * M + m] = sk;
      KA[((long)c * N + i) * M + m] = ka;
    }
  }
  agl = warp_sum(agl);
  ags = warp_sum(ags);
  if (lane == 0) {
    gl[(long)c * N + i] = agl;
    gs[(long)c * N + i] = ags;
  }
}

__global__ void __launch_bounds__(256) sep_finish_kernel(
    int model, int N, int M, int P, const double* __restrict__ pars, HyperConst h, const double* __restrict__ A,
    long strideA, int ld, const double* __restrict__ logdet, const int* __restrict__ info_in,
    const double* __restrict__ alpha, const double* __restrict__ yv, const double* __restrict__ s2v,
    const double* __restrict__ sig, const double* __restrict__ Lmat, const double* __restrict__ lam,
    const double* __restrict__ Vec, const double* __restrict__ gl, const double* __restrict__ gs,
    const double* __restrict__ rowSK, const double* __restrict__ KA, const double* __restrict__ Z0,
    const double* __restrict__ Z1, const double* __restrict__ G0, const double* __restrict__ G1,
    const double* __restrict__ hld0, const double* __restrict__ hld1, double* __restrict__ vals,
    double* __restrict__ grad, int* __restrict__ info) {
  __shared__ double scratch[40];
  __shared__ double D[16 * 16], dB[16 * 16], tmp[16 * 16];
  const int c = blockIdx.x;
  const int T = tril_size(M);
  const bool stat = (model == 0);
  const double* p = pars + (long)c * P;
  const double* uL = stat ? p + 2 : p + 2 * N;
  const double* Vc = Vec + (long)c * M * M;
  const double* Lc = Lmat + (long)c * M * M;

  // per-output reductions
  double quad = 0.0, trG = 0.0, ldsum = 0.0;
  int bad = 0;
  for (int m = 0; m < M; ++m) {
    const double* al = alpha + ((long)c * M + m) * N;
    const double* ym = yv + ((long)c * M + m) * N;
    const double* Am = A + ((long)c * M + m) * strideA;
    double q = 0.0, t = 0.0, sk = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const double a = al[i];
      q += ym[i] * a;
      t += -0.5 * Am[(long)i * ld + i] + 0.5 * a * a;
      sk += rowSK[((long)c * N + i) * M + m];
    }
    quad += block_sum(q, scratch);
    trG += block_sum(t, scratch);
    const double trSK = block_sum(sk, scratch);
    ldsum += logdet[(long)c * M + m];
    if (info_in[(long)c * M + m] != 0 && bad == 0) bad = info_in[(long)c * M + m];
    // D = 0.5 A^T Kx A - 0.5 diag(tr(S_m^-1 Kx))   (eigenbasis of B)
    for (int m2 = 0; m2 < M; ++m2) {
      double s = 0.0;
      for (int i = threadIdx.x; i < N; i += blockDim.x) s += al[i] * KA[((long)c * N + i) * M + m2];
      const double aka = block_sum(s, scratch);
      if (threadIdx.x == 0) D[m * 16 + m2] = 0.5 * aka - (m == m2 ? 0.5 * trSK : 0.0);
    }
  }
  __syncthreads();
  // dB = V D V^T ; dL = 2 tril(dB L)
  for (int idx = threadIdx.x; idx < M * M; idx += blockDim.x) {
    const int a = idx / M, b = idx % M;
    double s = 0.0;
    for (int k = 0; k < M; ++k) s += Vc[a * M + k] * D[k * 16 + b];
    tmp[a * 16 + b] = s;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < M * M; idx += blockDim.x) {
    const int a = idx / M, b = idx % M;
    double s = 0.0;
    for (int k = 0; k < M; ++k) s += tmp[a * 16 + k] * Vc[b * M + k];
    dB[a * 16 + b] = s;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < M * M; idx += blockDim.x) {
    const int a = idx / M, b = idx % M;
    double s = 0.0;
    for (int k = 0; k < M; ++k) s += 0.5 * (dB[a * 16 + k] + dB[k * 16 + a]) * Lc[k * M + b];
    tmp[a * 16 + b] = 2.0 * s;
  }
  __syncthreads();

  // priors
  const double s2 = s2v[c];
  const double ts2 = p[P - 1];
  const double log2pi = 1.8378770664093453;
  const double half_log2pi = h.half_log2pi;
  double lp_l, lp_sig = 0.0;
  if (stat) {
    const double dl = p[0] - h.s_loc;
    lp_l = -(dl * dl) / (2.0 * h.s_var) - h.s_logscale - half_log2pi;          // Normal.log_prob, logpos.py:446
  } else {
    double z0 = 0.0, z1 = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const double a = Z0[(long)c * N + i], b = Z1[(long)c * N + i];
      z0 += a * a;
      z1 += b * b;
    }
    const double q0 = block_sum(z0, scratch), q1 = block_sum(z1, scratch);
    lp_l = -0.5 * (N * log2pi + q0) - hld0[c];                                 // logpos.py:274
    lp_sig = -0.5 * (N * log2pi + q1) - hld1[c];                               // logpos.py:279
  }
  double lpu = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const double d = uL[t] - h.n_loc;
    lpu += -(d * d) / (2.0 * h.n_var) - h.n_logscale - half_log2pi;            // logpos.py:283,450
  }
  const double lp_uL = block_sum(lpu, scratch);
  const double lp_s = (-h.ig_a - 1.0) * log(s2) - h.ig_b / s2 + h.ig_alogb - h.ig_lgamma;
  const double loglik = -0.5 * ldsum - 0.5 * quad;
  double res = loglik;
  if (h.prior) res += lp_l + lp_sig + lp_uL + lp_s + ts2;
  const double pf = h.prior ? 1.0 : 0.0;
  if (threadIdx.x == 0) {
    double* v = vals + (long)c * 6;
    v[0] = -res; v[1] = loglik; v[2] = lp_l;
    if (stat) { v[3] = lp_uL; v[4] = lp_s; v[5] = 0.0; }
    else { v[3] = lp_sig; v[4] = lp_uL; v[5] = lp_s; }
    info[c] = bad;
  }
  if (grad == nullptr) return;
  double* g = grad + (long)c * P;
  double* guL = stat ? g + 2 : g + 2 * N;
  if (stat) {
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) { a += gl[(long)c * N + i]; b += gs[(long)c * N + i]; }
    const double sl = block_sum(a, scratch), ss = block_sum(b, scratch);
    if (threadIdx.x == 0) {
      g[0] = -(2.0 * sl + pf * (-(p[0] - h.s_loc) / h.s_var));
      g[1] = -(2.0 * ss);
    }
  } else {
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      g[i] = -(2.0 * gl[(long)c * N + i] - pf * G0[(long)c * N + i]);
      g[N + i] = -(2.0 * gs[(long)c * N + i] - pf * G1[(long)c * N + i]);
    }
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    int m = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while (m * (m + 1) / 2 > t) --m;
    while ((m + 1) * (m + 2) / 2 <= t) ++m;
    const int k = t - m * (m + 1) / 2;
    double d = tmp[m * 16 + k];
    if (k == m) d *= Lc[m * M + m];
    guL[t] = -(d + pf * (-(uL[t] - h.n_loc) / h.n_var));
  }
  if (threadIdx.x == 0) g[P - 1] = -(s2 * trG + pf * ((-h.ig_a - 1.0) + h.ig_b / s2 + 1.0));
}

}  // namespace

// =================================================================================================== host side
int padded_M(int M) {
  const int buckets[] = {2, 3, 4, 5, 6, 8, 10, 12, 16};
  for (int b : buckets)
    if (M <= b) return b;
  return -1;
}


#define NMGP_LAUNCH_CHECK()                 \
  do {                                      \
    NMGP_CUDA_TRY(cudaGetLastError());      \
    if (launches) ++*launches;              \
  } while (0)

int launch_rbf_cov(const double* x1, int N1, const double* x2, int N2, double alpha, double beta, double* out,
                   cudaStream_t st) {
  const int self = (x2 == nullptr);
  if (self) { x2 = x1; N2 = N1; }
  if (N1 <= 0 || N2 <= 0) return 0;
  dim3 grid((N2 + 127) / 128, N1);
  rbf_cov_kernel<<<grid, 128, 0, st>>>(x1, N1, x2, N2, alpha * alpha, beta, self, out);
  NMGP_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_gibbs_cov(const double* x1, const double* s1, const double* l1, int N1, const double* x2, const double* s2,
                     const double* l2, int N2, double* out, cudaStream_t st) {
  const int self = (x2 == nullptr);
  if (self) { x2 = x1; s2 = s1; l2 = l1; N2 = N1; }
  if (N1 <= 0 || N2 <= 0) return 0;
  dim3 grid((N2 + 127) / 128, N1);
  gibbs_cov_kernel<<<grid, 128, 0, st>>>(x1, s1, l1, N1, x2, s2, l2, N2, self, out);
  NMGP_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_prior_cov_blocks(const double* x, int cs, int N, double alpha, double beta, const BlockBatch& b,
                            cudaStream_t st, long* launches) {
  if (cs <= 0) return 0;
  dim3 grid((b.nP + 127) / 128, b.nP, cs);
  prior_cov_blocks_kernel<<<grid, 128, 0, st>>>(x, N, alpha * alpha, beta, b.A, b.strideA(), b.nP);
  NMGP_LAUNCH_CHECK();
  return 0;
}

int launch_extract_factor(const BlockBatch& b, int N, double* Lp, double* hld, cudaStream_t st, long* launches) {
  if (b.batch <= 0) return 0;
  dim3 grid((N + 127) / 128, N, b.batch);
  extract_factor_kernel<<<grid, 128, 0, st>>>(b.A, b.strideA(), b.nP, N, b.logdet, Lp, hld);
  NMGP_LAUNCH_CHECK();
  return 0;
}

int launch_prior_solve(const double* Lp, const double* rhs, double* out, int cs, int N, int nv, int trans,
                       cudaStream_t st, long* launches) {
  if (cs <= 0 || nv <= 0) return 0;
  const int nvc32 = (nv + 31) / 32;
  // warp-per-problem kernel only when there are enough problems to fill the GPU (>= one warp per SM sub-partition); for
  // a handful of subjects one warp's serial N^2/2 FMAs lose to the blocked path below (measured: N=500 4.9 vs 1.4 ms)
  // (also for long series: at N = 600, 3000 subjects, 21 right-hand sides the blocked path below is slower still -- 31.5 vs
  // 24.1 ms for the four solves, profiles/r02_hadamard.txt; both are far from the 17 GB they read)
  if ((long)cs * nvc32 >= 592) {
    const int grid = (int)(((long)cs * nvc32 + PWB - 1) / PWB);
    if (trans) prior_solve_warp_kernel<1><<<grid, 32 * PWB, 0, st>>>(Lp, rhs, out, cs, N, nv, nvc32);
    else prior_solve_warp_kernel<0><<<grid, 32 * PWB, 0, st>>>(Lp, rhs, out, cs, N, nv, nvc32);
    NMGP_LAUNCH_CHECK();
    return 0;
  }
  NMGP_CUDA_TRY(cudaMemcpyAsync(out, rhs, (size_t)cs * N * nv * sizeof(double), cudaMemcpyDeviceToDevice, st));
  const int nblk = (N + PBS - 1) / PBS;
  const int nvc = (nv + PVC - 1) / PVC;
  for (int bk = 0; bk < nblk; ++bk) {
    const int kb = trans ? nblk - 1 - bk : bk;
    const int k0 = kb * PBS;
    dim3 gd(cs, nvc);
    if (trans) prior_diag_solve_kernel<1><<<gd, 128, 0, st>>>(Lp, out, N, nv, k0);
    else prior_diag_solve_kernel<0><<<gd, 128, 0, st>>>(Lp, out, N, nv, k0);
    NMGP_LAUNCH_CHECK();
    const int others = trans ? kb : nblk - 1 - kb;
    if (others > 0) {
      dim3 gu(cs, others, nvc);
      if (trans) prior_update_kernel<1><<<gu, 256, 0, st>>>(Lp, out, N, nv, k0);
      else prior_update_kernel<0><<<gu, 256, 0, st>>>(Lp, out, N, nv, k0);
      NMGP_LAUNCH_CHECK();
    }
  }
  return 0;
}

int launch_nonseparable_cov_reference_order(const double* x, const double* pars, int batch, int N, int M, double* out,
                                            cudaStream_t st) {
  if (batch <= 0) return 0;
  const int n = N * M;
  dim3 grid((n + 127) / 128, n, batch);
  svc_cov_reference_order_kernel<<<grid, 128, 0, st>>>(x, pars, N, M, out);
  NMGP_CUDA_TRY(cudaGetLastError());
  return 0;
}

namespace {
__global__ void reduce_info_kernel(const int* __restrict__ in, int cs, int nmat, int* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cs) return;
  int f = 0;
  for (int m = 0; m < nmat && f == 0; ++m) f = in[c * nmat + m];
  out[c] = f;
}
}  // namespace

int launch_reduce_info(const int* info_mat, int cs, int nmat, int* info_out, cudaStream_t st, long* launches) {
  if (cs <= 0) return 0;
  reduce_info_kernel<<<(cs + 127) / 128, 128, 0, st>>>(info_mat, cs, nmat, info_out);
  NMGP_LAUNCH_CHECK();
  return 0;
}

int launch_kx(const double* x, const double* ell, const double* sig, int cs, int N, double* Kx, double* CK, cudaStream_t st,
              long* launches) {
  if (cs <= 0) return 0;
  dim3 gk((N + 127) / 128, N, cs);
  kx_kernel<<<gk, 128, 0, st>>>(x, ell, sig, N, Kx, CK);
  NMGP_LAUNCH_CHECK();
  return 0;
}

int launch_svc_prep(int cs, int N, int M, const double* pars, int P, const HyperConst& h, const Scratch& w, cudaStream_t st,
                    long* launches) {
  if (cs <= 0) return 0;
  svc_prep_kernel<<<cs, 256, 0, st>>>(pars, P, N, M, padded_M(M), h.mu0, h.mu1, w.ell, w.Lst, w.s2, w.R0, w.R1);
  NMGP_LAUNCH_CHECK();
  return 0;
}

int launch_symv(const BlockBatch& b, int n, const double* y, double* alpha, int batch, cudaStream_t st, long* launches) {
  if (batch <= 0) return 0;
  dim3 gs((n + 7) / 8, batch);
  symv_kernel<<<gs, 256, 0, st>>>(b.A, b.strideA(), b.nP, n, y, alpha);
  NMGP_LAUNCH_CHECK();
  return 0;
}

int svc_forward(int cs, int N, int M, const double* x, const double* pars, int P, const HyperConst& h, const Scratch& w,
                const BlockBatch& b, cudaStream_t st, long* launches) {
  if (cs <= 0) return 0;
  const int MT = padded_M(M);
  NMGP_TRY(launch_svc_prep(cs, N, M, pars, P, h, w, st, launches));
  dim3 gk((N + 127) / 128, N, cs);
  kx_kernel<<<gk, 128, 0, st>>>(x, w.ell, nullptr, N, w.Kx, w.CK);
  NMGP_LAUNCH_CHECK();
  dim3 gb(b.Kt * (b.Kt + 1) / 2, cs);
#define NMGP_BUILD_CASE(MM) case MM: svc_build_kernel<MM><<<gb, 256, 0, st>>>(w.Kx, w.Lst, w.s2, N, MT, b.A, b.strideA(), b.nP); break;
  switch (M) { NMGP_FOR_EACH_M(NMGP_BUILD_CASE) default: set_last_error("M out of range"); return -1; }
#undef NMGP_BUILD_CASE
  NMGP_LAUNCH_CHECK();
  return 0;
}

int svc_backward(int cs, int N, int M, const double* Y, const double* pars, int P, const HyperConst& h,
                 const Scratch& w, const BlockBatch& b, const double* hld0, const double* hld1, double* vals,
                 double* grad, int* info, cudaStream_t st, long* launches) {
  if (cs <= 0) return 0;
  const int n = N * M, MT = padded_M(M);
  // one pass over Sigma^-1: alpha and the Sigma^-1 part of the gradient sums
  const int LS = M * M;
  int JC = 32;    // time blocks staged per pass: small, so that registers (not shared memory) bound the occupancy
  if (JC > N) JC = N;
  dim3 gc((n + 127) / 128, cs);
  const size_t smem = (size_t)JC * (LS + M) * sizeof(double);
#define NMGP_CONTRACT_CASE(MM) case MM: svc_contract_kernel<MM><<<gc, 128, smem, st>>>(b.A, b.strideA(), b.nP, N, MT, JC, Y, w.Kx, w.CK, w.Lst, w.alpha, w.Wout, w.Vout); break;
  switch (M) { NMGP_FOR_EACH_M(NMGP_CONTRACT_CASE) default: set_last_error("M out of range"); return -1; }
#undef NMGP_CONTRACT_CASE
  NMGP_LAUNCH_CHECK();
  if (grad != nullptr) {
    dim3 gu((N + 127) / 128, cs);
#define NMGP_U_CASE(MM) case MM: svc_u_kernel<MM><<<gu, 128, 0, st>>>(N, MT, w.alpha, w.Lst, w.Ua); break;
    switch (M) { NMGP_FOR_EACH_M(NMGP_U_CASE) default: break; }
#undef NMGP_U_CASE
    NMGP_LAUNCH_CHECK();
    dim3 ga((N + 7) / 8, cs);
#define NMGP_ALPHA_CASE(MM) case MM: svc_alpha_terms_kernel<MM><<<ga, 256, 0, st>>>(N, MT, w.Ua, w.Kx, w.CK, w.Sa, w.Ca); break;
    switch (M) { NMGP_FOR_EACH_M(NMGP_ALPHA_CASE) default: break; }
#undef NMGP_ALPHA_CASE
    NMGP_LAUNCH_CHECK();
  }
  svc_finish_kernel<<<cs, 256, 0, st>>>(N, M, MT, P, pars, Y, h, b.A, b.strideA(), b.nP, b.logdet, b.info, w.alpha, w.s2,
                                        w.Lst, w.Wout, w.Vout, w.Sa, w.Ca, w.Z0, w.Z1, w.G0, w.G1, hld0, hld1, vals, grad, info);
  NMGP_LAUNCH_CHECK();
  return 0;
}

int launch_sep_prep(int model, int cs, int N, int M, const double* Y, const double* pars, int P, const HyperConst& h,
                    const Scratch& w, cudaStream_t st, long* launches) {
  if (cs <= 0) return 0;
  sep_prep_kernel<<<cs, 256, 0, st>>>(model, pars, P, N, M, Y, h.mu0, h.mu1, w.ell, w.sig, w.s2, w.Lst, w.lam, w.Vec, w.yv,
                                      w.R0, w.R1);
  NMGP_LAUNCH_CHECK();
  return 0;
}

int sep_forward(int model, int cs, int N, int M, const double* x, const double* Y, const double* pars, int P,
                const HyperConst& h, const Scratch& w, const BlockBatch& b, cudaStream_t st, long* launches) {
  if (cs <= 0) return 0;
  NMGP_TRY(launch_sep_prep(model, cs, N, M, Y, pars, P, h, w, st, launches));
  dim3 gk((N + 127) / 128, N, cs);
  kx_kernel<<<gk, 128, 0, st>>>(x, w.ell, w.sig, N, w.Kx, w.CK);
  NMGP_LAUNCH_CHECK();
  dim3 gb((b.nP + 127) / 128, b.nP, cs * M);
  sep_build_kernel<<<gb, 128, 0, st>>>(w.Kx, w.lam, w.s2, N, M, b.A, b.strideA(), b.nP);
  NMGP_LAUNCH_CHECK();
  return 0;
}

int sep_backward(int model, int cs, int N, int M, const double* pars, int P, const HyperConst& h, const Scratch& w,
                 const BlockBatch& b, const double* hld0, const double* hld1, double* vals, double* grad, int* info,
                 cudaStream_t st, long* launches) {
  if (cs <= 0) return 0;
  dim3 gs((N + 7) / 8, cs * M);
  symv_kernel<<<gs, 256, 0, st>>>(b.A, b.strideA(), b.nP, N, w.yv, w.alpha);
  NMGP_LAUNCH_CHECK();
  dim3 gc((N + 7) / 8, cs);
  sep_contract_kernel<<<gc, 256, 0, st>>>(b.A, b.strideA(), b.nP, N, M, w.alpha, w.lam, w.Kx, w.CK, w.sig, w.gl, w.gs,
                                          w.Wout, w.Vout);
  NMGP_LAUNCH_CHECK();
  sep_finish_kernel<<<cs, 256, 0, st>>>(model, N, M, P, pars, h, b.A, b.strideA(), b.nP, b.logdet, b.info, w.alpha, w.yv,
                                        w.s2, w.sig, w.Lst, w.lam, w.Vec, w.gl, w.gs, w.Wout, w.Vout, w.Z0, w.Z1, w.G0,
                                        w.G1, hld0, hld1, vals, grad, info);
  NMGP_LAUNCH_CHECK();
  return 0;
}

}  // namespace nmgp
