// Stand-alone Kronecker helpers of the reference (Utility/kronecker_operation.py:5-85, Utility/kernels.py:5-21,
// Utility/distributions.py:26-113) on the GPU with this library's own kernels.  They are COLD (SURVEY.md 8a rows a5, a8-a10:
// prediction.py / sim.py / deviance call them a handful of times), so they are written for clarity; the hot path never
// forms a Kronecker product.
//
// kron_eig_solve is the block formulation of the separable likelihood for ARBITRARY symmetric B (M x M) and
// K (N x N):  with B = V diag(lam) V^T (one-warp cyclic Jacobi),
//     sigma2 I + B (x) K = (V (x) I) blockdiag(S_m) (V (x) I)^T,   S_m = lam_m K + sigma2 I,
// so  log det = sum_m log det S_m,   inverse = sum_m (v_m v_m^T) (x) S_m^-1,   r^T inverse r = sum_m rt_m^T S_m^-1 rt_m with
// rt = (V^T (x) I) r.  The M matrices S_m go through the batched Cholesky + inverse engine (engine.cu) -- no N x N
// eigensolver, unlike the reference's two symeig calls (kronecker_operation.py:45-47, 67-68; distributions.py:37,40).
#include "engine.cuh"
#include "jacobi.cuh"
#include "models.cuh"

namespace nmgp {

namespace {

__global__ void pairwise_sqdist_kernel(const double* __restrict__ x1, int N1, const double* __restrict__ x2, int N2,
                                       double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j < N2) out[(long)i * N2 + j] = ref_sqdist(x1[i], x2[j]);   // (x_i^2 + y_j^2) - 2 x_i y_j, kernels.py:13-20
}

// out[(a h2 + i)][(b w2 + j)] = t1[a][b] * t2[i][j]   (kronecker_operation.py:5-22)
__global__ void kron_kernel(const double* __restrict__ t1, int h1, int w1, const double* __restrict__ t2, int h2, int w2,
                            double* __restrict__ out) {
  const long W = (long)w1 * w2;
  const long col = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long row = blockIdx.y;
  if (col >= W) return;
  const int a = (int)(row / h2), i = (int)(row % h2), b = (int)(col / w2), j = (int)(col % w2);
  out[row * W + col] = t1[(long)a * w1 + b] * t2[(long)i * w2 + j];
}

// out[i][j] = sum_k L[i][k] L[j][k]   (`generate_K_index_SVC`: stacked factors times their transpose, logpos.py:111-118)
__global__ void gram_kernel(const double* __restrict__ L, int R, int M, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= R) return;
  double s = 0.0;
  for (int k = 0; k < M; ++k) s += L[(long)i * M + k] * L[(long)j * M + k];
  out[(long)i * R + j] = s;
}

// stage 1 of kron_mv: T[m2][n1] = sum_n2 K[n1][n2] y[m2 N2 + n2]      (K Y, kronecker_operation.py:83)
__global__ void kron_mv_ky_kernel(const double* __restrict__ K, int N1, int N2, const double* __restrict__ y, int M2,
                                  double* __restrict__ Tm) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N1 * M2) return;
  const int n1 = warp % N1, m2 = warp / N1;
  double s = 0.0;
  for (int n2 = lane; n2 < N2; n2 += 32) s += K[(long)n1 * N2 + n2] * y[(long)m2 * N2 + n2];
  s = warp_sum(s);
  if (lane == 0) Tm[(long)m2 * N1 + n1] = s;
}
// stage 2: out[m1 N1 + n1] = sum_m2 T[m2][n1] B[m1][m2]               ((K Y) B^T, transposed back: :83-84)
__global__ void kron_mv_b_kernel(const double* __restrict__ Tm, const double* __restrict__ B, int M1, int M2, int N1,
                                 double* __restrict__ out) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M1 * N1) return;
  const int m1 = (int)(idx / N1), n1 = (int)(idx % N1);
  double s = 0.0;
  for (int m2 = 0; m2 < M2; ++m2) s += Tm[(long)m2 * N1 + n1] * B[(long)m1 * M2 + m2];
  out[idx] = s;
}

// eigen-decomposition of one symmetric M x M matrix (M <= 16): eigenvalues ascending (as torch.symeig returns them),
// eigenvectors in the columns of V [M][M]
__global__ void __launch_bounds__(32) sym_eig_kernel(const double* __restrict__ B, int M, double* __restrict__ lam,
                                                     double* __restrict__ V) {
  __shared__ double Bs[16 * 17], Vs[16 * 17];
  __shared__ int order[16];
  const int lane = threadIdx.x;
  for (int idx = lane; idx < M * M; idx += 32) {
    const int a = idx / M, b = idx % M;
    Bs[a * 17 + b] = 0.5 * (B[a * M + b] + B[b * M + a]);
  }
  __syncwarp();
  jacobi_eig_warp(Bs, Vs, M, 17);
  if (lane == 0) {
    for (int m = 0; m < M; ++m) order[m] = m;
    for (int a = 0; a < M; ++a)
      for (int b = a + 1; b < M; ++b)
        if (Bs[order[b] * 17 + order[b]] < Bs[order[a] * 17 + order[a]]) { const int t = order[a]; order[a] = order[b]; order[b] = t; }
  }
  __syncwarp();
  for (int m = lane; m < M; m += 32) lam[m] = Bs[order[m] * 17 + order[m]];
  for (int idx = lane; idx < M * M; idx += 32) V[idx] = Vs[(idx / M) * 17 + order[idx % M]];
}

// S_m = lam_m K + sigma2 I in the engine's padded layout (identity on the padding)
__global__ void kron_build_kernel(const double* __restrict__ K, const double* __restrict__ lam, double sigma2, int N,
                                  double* __restrict__ A, long strideA, int ld) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y, m = blockIdx.z;
  if (q >= ld) return;
  double v;
  if (p < N && q < N) {
    v = lam[m] * (0.5 * (K[(long)p * N + q] + K[(long)q * N + p]));
    if (p == q) v += sigma2;
  } else {
    v = (p == q) ? 1.0 : 0.0;
  }
  A[(long)m * strideA + (long)p * ld + q] = v;
}

// rt[m][i] = sum_m' V[m'][m] r[m' N + i]        ((V^T (x) I) r, r output-major)
__global__ void kron_rotate_kernel(const double* __restrict__ r, const double* __restrict__ V, int M, int N,
                                   double* __restrict__ rt) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * N) return;
  const int m = (int)(idx / N), i = (int)(idx % N);
  double s = 0.0;
  for (int k = 0; k < M; ++k) s += V[k * M + m] * r[(long)k * N + i];
  rt[idx] = s;
}

// one warp per (m, i): partial[m N + i] = rt[m][i] * sum_j Sinv_m[i][j] rt[m][j]
__global__ void kron_quad_kernel(const double* __restrict__ A, long strideA, int ld, const double* __restrict__ rt, int M,
                                 int N, double* __restrict__ partial) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M * N) return;
  const int m = warp / N, i = warp % N;
  const double* row = A + (long)m * strideA + (long)i * ld;
  double s = 0.0;
  for (int j = lane; j < N; j += 32) s += row[j] * rt[(long)m * N + j];
  s = warp_sum(s);
  if (lane == 0) partial[warp] = s * rt[(long)m * N + i];
}

// out[0] = sum of logdet[0..M), out[1] = sum of partial[0..n)   (fixed order: one CTA)
__global__ void __launch_bounds__(256) kron_finish_kernel(const double* __restrict__ logdet, int M,
                                                          const double* __restrict__ partial, long n,
                                                          double* __restrict__ out) {
  __shared__ double scratch[40];
  double s = 0.0;
  for (long i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  const double q = block_sum(s, scratch);
  if (threadIdx.x == 0) {
    double ld = 0.0;
    for (int m = 0; m < M; ++m) ld += logdet[m];
    out[0] = ld;
    out[1] = q;
  }
}

// inv[(a N + i)][(b N + j)] = sum_m V[a][m] V[b][m] Sinv_m[i][j]
__global__ void kron_assemble_kernel(const double* __restrict__ A, long strideA, int ld, const double* __restrict__ V,
                                     int M, int N, double* __restrict__ out) {
  const long n = (long)M * N;
  const long col = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long row = blockIdx.y;
  if (col >= n) return;
  const int a = (int)(row / N), i = (int)(row % N), b = (int)(col / N), j = (int)(col % N);
  double s = 0.0;
  for (int m = 0; m < M; ++m) s += V[a * M + m] * V[b * M + m] * A[(long)m * strideA + (long)i * ld + j];
  out[row * n + col] = s;
}

}  // namespace

int launch_pairwise_sqdist(const double* x1, int N1, const double* x2, int N2, double* out, cudaStream_t st) {
  if (!x2) { x2 = x1; N2 = N1; }
  if (N1 <= 0 || N2 <= 0) return 0;
  dim3 grid((N2 + 127) / 128, N1);
  pairwise_sqdist_kernel<<<grid, 128, 0, st>>>(x1, N1, x2, N2, out);
  NMGP_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_kron(const double* t1, int h1, int w1, const double* t2, int h2, int w2, double* out, cudaStream_t st) {
  if (h1 <= 0 || w1 <= 0 || h2 <= 0 || w2 <= 0) return 0;
  const long W = (long)w1 * w2, H = (long)h1 * h2;
  if (H > 65535) { set_last_error("nmgp_kron: more than 65535 output rows"); return -1; }
  dim3 grid((unsigned)((W + 127) / 128), (unsigned)H);
  kron_kernel<<<grid, 128, 0, st>>>(t1, h1, w1, t2, h2, w2, out);
  NMGP_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_gram(const double* L, int R, int M, double* out, cudaStream_t st) {
  if (R <= 0 || M <= 0) return 0;
  if (R > 65535) { set_last_error("nmgp_gram: more than 65535 rows"); return -1; }
  dim3 grid((R + 127) / 128, R);
  gram_kernel<<<grid, 128, 0, st>>>(L, R, M, out);
  NMGP_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_kron_mv(const double* B, int M1, int M2, const double* K, int N1, int N2, const double* y, double* out,
                   double* scratch, cudaStream_t st) {
  if (M1 <= 0 || M2 <= 0 || N1 <= 0 || N2 <= 0) return 0;
  const long warps = (long)N1 * M2;
  kron_mv_ky_kernel<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, st>>>(K, N1, N2, y, M2, scratch);
  NMGP_CUDA_TRY(cudaGetLastError());
  kron_mv_b_kernel<<<(unsigned)(((long)M1 * N1 + 127) / 128), 128, 0, st>>>(scratch, B, M1, M2, N1, out);
  NMGP_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_sym_eig(const double* B, int M, double* lam, double* V, cudaStream_t st) {
  if (M <= 0 || M > 16) { set_last_error("nmgp_sym_eig: 1 <= M <= 16"); return -1; }
  sym_eig_kernel<<<1, 32, 0, st>>>(B, M, lam, V);
  NMGP_CUDA_TRY(cudaGetLastError());
  return 0;
}

int kron_eig_solve(const double* B, int M, const double* K, int N, double sigma2, double* inv_out, const double* r,
                   double* out2, int* info_out, cudaStream_t st) {
  if (M <= 0 || M > 16 || N <= 0) { set_last_error("nmgp_kron_eig_solve: need 1 <= M <= 16, N >= 1"); return -1; }
  BlockBatch b;
  b.n = N; b.nP = padded_dim(N); b.Kt = b.nP / kNB; b.NB = kNB; b.batch = M;
  const bool need_inverse = inv_out != nullptr || r != nullptr;
  const size_t nA = (size_t)M * b.strideA(), nD = (size_t)M * b.strideD();
  const size_t small = (size_t)M + (size_t)M * M + (size_t)M + 2 * (size_t)M * N;   // lam, V, logdet, rt, partial
  double* ws = nullptr;
  int* inf = nullptr;
  if (cudaMalloc(&ws, (nA + nD + small) * sizeof(double)) != cudaSuccess || cudaMalloc(&inf, M * sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    if (ws) cudaFree(ws);
    set_last_error("nmgp_kron_eig_solve: out of device memory");
    return -3;
  }
  b.A = ws; b.Dinv = ws + nA; b.info = inf;
  double* lam = ws + nA + nD;
  double* V = lam + M;
  b.logdet = V + (size_t)M * M;
  double* rt = b.logdet + M;
  double* partial = rt + (size_t)M * N;
  int rc = launch_sym_eig(B, M, lam, V, st);
  if (rc == 0) {
    dim3 gb((b.nP + 127) / 128, b.nP, M);
    kron_build_kernel<<<gb, 128, 0, st>>>(K, lam, sigma2, N, b.A, b.strideA(), b.nP);
    if (cudaGetLastError() != cudaSuccess) { set_last_error("nmgp_kron_eig_solve: launch failed"); rc = -2; }
  }
  if (rc == 0) rc = engine_potrf(b, st, nullptr);
  if (rc == 0 && need_inverse) rc = engine_potri(b, st, nullptr);
  long npart = 0;
  if (rc == 0 && r) {
    const long tot = (long)M * N;
    kron_rotate_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(r, V, M, N, rt);
    kron_quad_kernel<<<(unsigned)((tot * 32 + 127) / 128), 128, 0, st>>>(b.A, b.strideA(), b.nP, rt, M, N, partial);
    npart = tot;
  }
  if (rc == 0 && out2) kron_finish_kernel<<<1, 256, 0, st>>>(b.logdet, M, partial, npart, out2);
  if (rc == 0 && inv_out) {
    const long n = (long)M * N;
    if (n > 65535) { set_last_error("nmgp_kron_eig_solve: inverse with more than 65535 rows"); rc = -1; }
    else {
      dim3 ga((unsigned)((n + 127) / 128), (unsigned)n);
      kron_assemble_kernel<<<ga, 128, 0, st>>>(b.A, b.strideA(), b.nP, V, M, N, inv_out);
    }
  }
  if (rc == 0 && info_out) {
    // first failing pivot over the M factorisations (0 = all positive definite)
    rc = launch_reduce_info(b.info, 1, M, info_out, st, nullptr);
  }
  if (rc == 0 && cudaGetLastError() != cudaSuccess) { set_last_error("nmgp_kron_eig_solve: launch failed"); rc = -2; }
  cudaStreamSynchronize(st);
  cudaFree(ws);
  cudaFree(inf);
  return rc;
}

}  // namespace nmgp
