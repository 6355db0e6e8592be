// Left-looking batched Cholesky + "Takahashi" inverse for many mid-size matrices (the 10 000 x (n=600) regime).
//
// The right-looking tile tasks of engine.cu read AND write every trailing tile at every block step: at n = 600 that is
// ~80 MB of traffic per matrix and the kernels sit on the HBM roof at ~45 % of the FP64 tensor peak (profiles/r01).
// Here every output tile is produced ONCE: a CTA accumulates its 64 x 64 tile over the whole K range in registers
// (DMMA.8x8x4), streaming the two operand row-panels through a 3-stage cp.async pipeline in shared memory.
//
//   potrf, block column k:   A(i,k) <- A(i,k) - sum_{j<k} A(i,j) A(k,j)^T        (LL_UPDATE,  i >= k)
//                            diag_kernel: A(k,k) = L L^T, W_kk = L^-1, log det
//                            A(i,k) <- A(i,k) W_kk^T                              (LL_SOLVE,   i > k)
//   inverse, block column j descending (Z = Sigma^-1, stored full symmetric, in place over L):
//                            P(c)   = -L(c,j) W_jj                                (TK_PANEL,   c > j, side buffer)
//                            Z(i,j) = sum_{c>j} Z(i,c) P(c),  Z(j,i) = Z(i,j)^T   (TK_COL,     i > j)
//                            Z(j,j) = W_jj^T W_jj + sum_{c>j} P(c)^T Z(c,j)       (TK_DIAG)
// Flops: n^3/3 + 2n^3/3, the same as potrf + potri; every A/B operand tile is read once per use, C never re-read.
#include "engine.cuh"

namespace nmgp {

namespace {

constexpr int NB = kNB;      // 64
constexpr int KC = 16;       // k-chunk per pipeline stage
constexpr int STAGES = 3;
constexpr int LDK = KC + 4;  // k-major chunk [64][20]   (== 4 mod 16: conflict-free fragment loads)
constexpr int LDM = NB + 4;  // m-major chunk [16][68]
constexpr int OPSZ = NB * LDK;  // 1280 doubles >= KC*LDM = 1088
constexpr int THREADS = 128;
constexpr size_t LL_SMEM = (size_t)STAGES * 2 * OPSZ * sizeof(double);  // 61 440 B -> 3 CTAs / SM

enum LLMode : int { LL_UPDATE = 0, LL_SOLVE = 1, TK_PANEL = 2, TK_COL = 3, TK_DIAG = 4 };

struct LLArgs {
  double* A;
  double* Dinv;
  double* Pbuf;
  long strideA, strideD, strideP;
  int ld, Kt, batch, step, n8;
};

__device__ __forceinline__ void cp_async16(double* smem, const double* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// issue the copies of one k-chunk of one operand tile
template <bool KM>
__device__ __forceinline__ void issue_chunk(double* S, const double* tile, int ld, int kc) {
  if (KM) {  // rows 0..63, columns kc*16 .. +16  ->  S[row][LDK]
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = threadIdx.x + it * THREADS;  // 0..511
      const int r = idx >> 3, c2 = idx & 7;
      cp_async16(S + r * LDK + 2 * c2, tile + (long)r * ld + kc * KC + 2 * c2);
    }
  } else {   // rows kc*16 .. +16, columns 0..63  ->  S[k][LDM]
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = threadIdx.x + it * THREADS;
      const int r = idx >> 5, c2 = idx & 31;
      cp_async16(S + r * LDM + 2 * c2, tile + (long)(kc * KC + r) * ld + 2 * c2);
    }
  }
}

template <bool A_KM, bool B_KM>
__device__ __forceinline__ void mma_chunk(const double* __restrict__ SA, const double* __restrict__ SB, int m0, int n0,
                                          double (&acc)[4][4][2]) {
  const int lane = threadIdx.x & 31;
  const int lr = lane >> 2, lk = lane & 3;
  const double* pa = A_KM ? SA + (m0 + lr) * LDK + lk : SA + lk * LDM + m0 + lr;
  const double* pb = B_KM ? SB + (n0 + lr) * LDK + lk : SB + lk * LDM + n0 + lr;
  constexpr int a_sub = A_KM ? 8 * LDK : 8;
  constexpr int b_sub = B_KM ? 8 * LDK : 8;
  constexpr int a_k = A_KM ? 4 : 4 * LDM;
  constexpr int b_k = B_KM ? 4 : 4 * LDM;
#pragma unroll
  for (int k = 0; k < KC; k += 4) {
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = pa[i * a_sub];
      b[i] = pb[i * b_sub];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    pa += a_k;
    pb += b_k;
  }
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 3) panel_gemm_kernel(LLArgs g) {
  extern __shared__ __align__(16) double smem[];
  constexpr bool A_KM = (MODE == LL_UPDATE || MODE == LL_SOLVE || MODE == TK_PANEL || MODE == TK_COL);
  constexpr bool B_KM = (MODE == LL_UPDATE || MODE == LL_SOLVE);
  const int s = g.step;
  const int t = blockIdx.x;
  const int last = g.Kt - 1;
  const int vlast = g.n8 - last * NB;  // valid rows/cols of the last block (multiple of 8)

  // output tile (i,j) and the k-block sequence
  int i, j, nkb;
  if (MODE == LL_UPDATE) { i = s + t; j = s; nkb = s; }
  else if (MODE == LL_SOLVE) { i = s + 1 + t; j = s; nkb = 1; }
  else if (MODE == TK_PANEL) { i = s + 1 + t; j = s; nkb = 1; }
  else if (MODE == TK_COL) { i = s + 1 + t; j = s; nkb = last - s; }
  else { i = s; j = s; nkb = last - s + 1; }
  // is the last k-block of the sequence the ragged block Kt-1 ?
  const bool ragged_k = (MODE == TK_COL) || (MODE == TK_DIAG && nkb > 1);
  const int nchunks = (nkb - 1) * (NB / KC) + ((ragged_k ? vlast : NB) + KC - 1) / KC;
  const int rows_valid = (i == last) ? vlast : NB;
  const int cols_valid = (j == last) ? vlast : NB;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (warp >> 1) * 32, n0 = (warp & 1) * 32;
  const bool active = (m0 < rows_valid) && (n0 < cols_valid);

  for (int mat = blockIdx.y; mat < g.batch; mat += gridDim.y) {
    double* Am = g.A + (long)mat * g.strideA;
    const double* Dm = g.Dinv + (long)mat * g.strideD + (long)s * NB * NB;
    double* Pm = g.Pbuf + (long)mat * g.strideP;

    auto operands = [&](int kb, const double*& ta, int& lda, const double*& tb, int& ldb) {
      if (MODE == LL_UPDATE) {
        ta = Am + ((long)i * NB) * g.ld + (long)kb * NB; lda = g.ld;
        tb = Am + ((long)s * NB) * g.ld + (long)kb * NB; ldb = g.ld;
      } else if (MODE == LL_SOLVE) {
        ta = Am + ((long)i * NB) * g.ld + (long)s * NB; lda = g.ld;
        tb = Dm; ldb = NB;
      } else if (MODE == TK_PANEL) {
        ta = Am + ((long)i * NB) * g.ld + (long)s * NB; lda = g.ld;
        tb = Dm; ldb = NB;
      } else if (MODE == TK_COL) {
        const int c = s + 1 + kb;
        ta = Am + ((long)i * NB) * g.ld + (long)c * NB; lda = g.ld;
        tb = Pm + (long)c * NB * NB; ldb = NB;
      } else {
        if (kb == 0) { ta = Dm; lda = NB; tb = Dm; ldb = NB; }
        else {
          const int c = s + kb;
          ta = Pm + (long)c * NB * NB; lda = NB;
          tb = Am + ((long)c * NB) * g.ld + (long)s * NB; ldb = g.ld;
        }
      }
    };
    auto issue = [&](int q) {
      const int kb = q / (NB / KC), kc = q % (NB / KC);
      const double *ta, *tb;
      int lda, ldb;
      operands(kb, ta, lda, tb, ldb);
      double* S = smem + (q % STAGES) * 2 * OPSZ;
      issue_chunk<A_KM>(S, ta, lda, kc);
      issue_chunk<B_KM>(S + OPSZ, tb, ldb, kc);
    };

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

#pragma unroll
    for (int q = 0; q < STAGES - 1; ++q) {
      if (q < nchunks) issue(q);
      cp_async_commit();
    }
    for (int q = 0; q < nchunks; ++q) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      if (q + STAGES - 1 < nchunks) issue(q + STAGES - 1);
      cp_async_commit();
      if (active) {
        const double* S = smem + (q % STAGES) * 2 * OPSZ;
        mma_chunk<A_KM, B_KM>(S, S + OPSZ, m0, n0, acc);
      }
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- epilogue
    const int r = lane >> 2, c = 2 * (lane & 3);
    if (MODE == TK_PANEL) {
      double* C = Pm + (long)i * NB * NB;
      if (active) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            double2 v; v.x = -acc[a][b][0]; v.y = -acc[a][b][1];
            *reinterpret_cast<double2*>(C + (long)(m0 + 8 * a + r) * NB + n0 + 8 * b + c) = v;
          }
      }
    } else {
      double* C = Am + ((long)i * NB) * g.ld + (long)j * NB;
      if (active) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            double2* p = reinterpret_cast<double2*>(C + (long)(m0 + 8 * a + r) * g.ld + n0 + 8 * b + c);
            double2 v;
            if (MODE == LL_UPDATE) {
              v = *p;
              v.x -= acc[a][b][0];
              v.y -= acc[a][b][1];
            } else {
              v.x = acc[a][b][0];
              v.y = acc[a][b][1];
            }
            *p = v;
            acc[a][b][0] = v.x;
            acc[a][b][1] = v.y;
          }
      }
      if (MODE == TK_COL) {  // mirror: Z(j,i) = Z(i,j)^T through shared memory (pipeline buffers are free now)
        double* T = smem;    // [64][65]
        if (active) {
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              T[(m0 + 8 * a + r) * (NB + 1) + n0 + 8 * b + c] = acc[a][b][0];
              T[(m0 + 8 * a + r) * (NB + 1) + n0 + 8 * b + c + 1] = acc[a][b][1];
            }
        }
        __syncthreads();
        double* U = Am + ((long)j * NB) * g.ld + (long)i * NB;  // rows = cols of the tile, cols = its rows
        const int rv = (rows_valid + 31) & ~31, cv = (cols_valid + 31) & ~31;  // what the active warps produced
        for (int idx = threadIdx.x; idx < cv * rv; idx += THREADS) {
          const int ur = idx / rv, uc = idx % rv;
          U[(long)ur * g.ld + uc] = T[uc * (NB + 1) + ur];
        }
      }
    }
    __syncthreads();
  }
}

template <int MODE>
int launch_ll(const LLArgs& g, int ntiles, cudaStream_t st, long* launches) {
  if (ntiles <= 0 || g.batch <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    NMGP_CUDA_TRY(cudaFuncSetAttribute(panel_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LL_SMEM));
    configured = true;
  }
  dim3 grid(ntiles, g.batch < 65535 ? g.batch : 65535);
  panel_gemm_kernel<MODE><<<grid, THREADS, LL_SMEM, st>>>(g);
  NMGP_CUDA_TRY(cudaGetLastError());
  if (launches) ++*launches;
  return 0;
}

LLArgs make_ll(const BlockBatch& b) {
  LLArgs g;
  g.A = b.A; g.Dinv = b.Dinv; g.Pbuf = b.Pbuf;
  g.strideA = b.strideA(); g.strideD = b.strideD(); g.strideP = b.strideD();
  g.ld = b.nP; g.Kt = b.Kt; g.batch = b.batch; g.step = 0;
  g.n8 = (int)round_up(b.n, 8);
  return g;
}

}  // namespace

int engine_potrf_ll(const BlockBatch& b, cudaStream_t st, long* launches) {
  if (b.batch <= 0) return 0;
  LLArgs g = make_ll(b);
  for (int k = 0; k < b.Kt; ++k) {
    g.step = k;
    if (k > 0) NMGP_TRY(launch_ll<LL_UPDATE>(g, b.Kt - k, st, launches));
    NMGP_TRY(engine_diag_step(b, k, st, launches));
    NMGP_TRY(launch_ll<LL_SOLVE>(g, b.Kt - k - 1, st, launches));
  }
  return 0;
}

int engine_potri_ll(const BlockBatch& b, cudaStream_t st, long* launches) {
  if (b.batch <= 0) return 0;
  if (!b.Pbuf) { set_last_error("engine_potri_ll: no panel buffer"); return -1; }
  LLArgs g = make_ll(b);
  for (int j = b.Kt - 1; j >= 0; --j) {
    g.step = j;
    NMGP_TRY(launch_ll<TK_PANEL>(g, b.Kt - 1 - j, st, launches));
    NMGP_TRY(launch_ll<TK_COL>(g, b.Kt - 1 - j, st, launches));
    NMGP_TRY(launch_ll<TK_DIAG>(g, 1, st, launches));
  }
  return 0;
}

}  // namespace nmgp
