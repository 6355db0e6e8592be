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
constexpr int KC = 16;       // k-chunk per pipeline stage (tools/dmma_pipe.cu: 16 x 3 stages beats 32 x 2 and 8 x 4)
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
  int ld, Kt, batch, step, n8, ntiles;
};

// cp.async with an L2 eviction-priority hint.  The streamed operand (tiles of the matrix that this launch reads once)
// is marked evict_first, the operand every row tile of a block column re-reads (the P / row-s panel, the W_kk tile)
// evict_last: without the hints the streamed tiles flush the shared panel out of the 126 MB L2 before its next use
// (ncu: 157 GB of DRAM traffic per sweep in the Takahashi column kernel against 106 GB algorithmic).
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cp_async16(double* smem, const double* gmem, unsigned long long policy) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "l"(policy));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// issue the copies of one k-chunk of one operand tile
template <bool KM>
__device__ __forceinline__ void issue_chunk(double* S, const double* tile, int ld, int kc, unsigned long long policy) {
  constexpr int V = KC / 2;               // 16-byte vectors per row of a k-major chunk
  constexpr int ITERS = NB * V / THREADS; // copies per thread (both layouts move NB*KC doubles)
  if (KM) {  // rows 0..63, columns kc*KC .. +KC  ->  S[row][LDK]
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int idx = threadIdx.x + it * THREADS;
      const int r = idx / V, c2 = idx % V;
      cp_async16(S + r * LDK + 2 * c2, tile + (long)r * ld + kc * KC + 2 * c2, policy);
    }
  } else {   // rows kc*KC .. +KC, columns 0..63  ->  S[k][LDM]
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int idx = threadIdx.x + it * THREADS;
      const int r = idx >> 5, c2 = idx & 31;
      cp_async16(S + r * LDM + 2 * c2, tile + (long)(kc * KC + r) * ld + 2 * c2, policy);
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

// Measured dead ends (profiles/README.md): prefetch.global.L2 of the operand tiles 3 k-blocks ahead (-2 % : slower),
// KC = 32 x 2 stages (-1 %), generating the covariance tile inside the LL_UPDATE epilogue instead of reading it
// (+8 ms in potrf for 6 ms saved in the build kernel: the epilogue's loads sit on the critical path of the tensor pipe).
// Also measured: fusing the diagonal-block factorisation and the panel solve of a block column into one kernel
// (one CTA per matrix: pivots, then the W_kk^T multiplies) is 7 ms SLOWER per sweep -- a CTA slot (1/3 of an SM at this
// register / shared-memory footprint) sits in the latency-bound pivot chain for ~45 us; as its own kernel the chain runs
// at 4-5 CTAs per SM.
// One CTA walks a LIST of output tiles of one matrix (tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of this launch) and
// streams all their k-chunks through ONE cp.async ring: the loads of the next tile are already in flight while the
// current tile's epilogue runs, so the pipeline never drains between tiles (with one tile per CTA the fill/drain cost
// 20-30 % of the short panel modes).  The epilogue works from registers only (the mirror of the symmetric inverse is a
// direct transposed store: 8 lanes write 64 contiguous bytes), so shared memory belongs to the ring at all times.
template <int MODE>
__global__ void __launch_bounds__(THREADS, 3) panel_gemm_kernel(LLArgs g) {
  extern __shared__ __align__(16) double smem[];
  constexpr bool A_KM = (MODE == LL_UPDATE || MODE == LL_SOLVE || MODE == TK_PANEL || MODE == TK_COL);
  constexpr bool B_KM = (MODE == LL_UPDATE || MODE == LL_SOLVE);
  constexpr int CPB = NB / KC;   // chunks per k-block
  const int s = g.step;
  const int last = g.Kt - 1;
  const int vlast = g.n8 - last * NB;  // valid rows/cols of the last block (multiple of 8)
  const int j = s;                     // every mode writes block column s

  // k-block sequence (the same for every tile of a launch)
  int nkb;
  if (MODE == LL_UPDATE) nkb = s;
  else if (MODE == LL_SOLVE || MODE == TK_PANEL) nkb = 1;
  else if (MODE == TK_COL) nkb = last - s;
  else nkb = last - s + 1;
  // is the last k-block of the sequence the ragged block Kt-1 ?
  const bool ragged_k = (MODE == TK_COL) || (MODE == TK_DIAG && nkb > 1);
  const int nchunks = (nkb - 1) * CPB + ((ragged_k ? vlast : NB) + KC - 1) / KC;
  const int cols_valid = (j == last) ? vlast : NB;

  // Quarter of the tile owned by this warp.  The quarters do unequal work in the triangular modes below and every
  // warp is pinned to one SM sub-partition (one DMMA pipe each), so the assignment is rotated per CTA: over the CTAs
  // resident on an SM the light and heavy quarters then spread over all four pipes.
  const int warp = ((threadIdx.x >> 5) + blockIdx.x + blockIdx.y) & 3, lane = threadIdx.x & 31;
  const int m0 = (warp >> 1) * 32, n0 = (warp & 1) * 32;
  const int r = lane >> 2, c = 2 * (lane & 3);
  const bool upper_quarter = (m0 == 0 && n0 == NB / 2);
  // k-chunks that only multiply structural zeros of the triangular W = L_kk^-1 (W[a][b] = 0 for b > a):
  //   LL_SOLVE  C[m][n] = sum_k A[m][k] W[n][k]   -> columns n < 32 need k < 32 only
  //   TK_PANEL  C[m][n] = sum_k L[m][k] W[k][n]   -> columns n >= 32 need k >= 32 only
  //   TK_DIAG   (first k-block) sum_k W[k][m] W[k][n] -> any quarter touching rows/cols >= 32 needs k >= 32 only
  auto zero_chunk = [&](int qq) -> bool {
    constexpr int HALF = (NB / 2) / KC;   // chunks per half block
    if (MODE == LL_SOLVE) return n0 == 0 && qq >= HALF;
    if (MODE == TK_PANEL) return n0 == NB / 2 && qq < HALF;
    if (MODE == TK_DIAG) return (m0 == NB / 2 || n0 == NB / 2) && qq < HALF;
    return false;
  };
  // block row of the tt-th tile of this CTA
  auto tile_row = [&](int tt) -> int {
    const int t = blockIdx.x + tt * gridDim.x;
    if (MODE == LL_UPDATE) return s + t;
    if (MODE == TK_DIAG) return s;
    return s + 1 + t;
  };
  const int nmy = ((int)blockIdx.x < g.ntiles) ? (g.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int total = nmy * nchunks;

  const unsigned long long pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  for (int mat = blockIdx.y; mat < g.batch; mat += gridDim.y) {
    double* Am = g.A + (long)mat * g.strideA;
    const double* Dm = g.Dinv + (long)mat * g.strideD + (long)s * NB * NB;
    double* Pm = g.Pbuf + (long)mat * g.strideP;

    int itt = 0, iqq = 0;   // tile / chunk-in-tile of the next chunk to be issued
    auto issue = [&](int q) {
      const int tt = itt, qq = iqq;
      if (++iqq == nchunks) { iqq = 0; ++itt; }
      const int i = tile_row(tt);
      const int kb = qq / CPB, kc = qq % CPB;
      const double *ta, *tb;
      int lda, ldb;
      if (MODE == LL_UPDATE) {
        ta = Am + ((long)i * NB) * g.ld + (long)kb * NB; lda = g.ld;
        tb = Am + ((long)s * NB) * g.ld + (long)kb * NB; ldb = g.ld;
      } else if (MODE == LL_SOLVE || MODE == TK_PANEL) {
        ta = Am + ((long)i * NB) * g.ld + (long)s * NB; lda = g.ld;
        tb = Dm; ldb = NB;
      } else if (MODE == TK_COL) {
        const int cb = s + 1 + kb;
        ta = Am + ((long)i * NB) * g.ld + (long)cb * NB; lda = g.ld;
        tb = Pm + (long)cb * NB * NB; ldb = NB;
      } else {
        if (kb == 0) { ta = Dm; lda = NB; tb = Dm; ldb = NB; }
        else {
          const int cb = s + kb;
          ta = Pm + (long)cb * NB * NB; lda = NB;
          tb = Am + ((long)cb * NB) * g.ld + (long)s * NB; ldb = g.ld;
        }
      }
      double* S = smem + (q % STAGES) * 2 * OPSZ;
      // which operand is the streamed one: A everywhere except TK_DIAG, where A is the shared P panel / W tile
      issue_chunk<A_KM>(S, ta, lda, kc, MODE == TK_DIAG ? pol_keep : pol_stream);
      issue_chunk<B_KM>(S + OPSZ, tb, ldb, kc, (MODE == TK_DIAG && kb > 0) ? pol_stream : pol_keep);
    };

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

#pragma unroll
    for (int q = 0; q < STAGES - 1; ++q) {
      if (q < total) issue(q);
      cp_async_commit();
    }
    int tt = 0, qq = 0;
    for (int q = 0; q < total; ++q) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      if (q + STAGES - 1 < total) issue(q + STAGES - 1);
      cp_async_commit();
      const int i = tile_row(tt);
      const int rows_valid = (i == last) ? vlast : NB;
      // structurally unnecessary quarter: strictly-upper 32 x 32 of a diagonal tile (potrf reads only the lower triangle
      // of A(k,k); Z(j,j) is symmetric and mirrored below)
      const bool tri_skip = upper_quarter && i == j && (MODE == LL_UPDATE || MODE == TK_DIAG);
      const bool active = (m0 < rows_valid) && (n0 < cols_valid) && !tri_skip;
      if (active && !zero_chunk(qq)) {
        const double* S = smem + (q % STAGES) * 2 * OPSZ;
        mma_chunk<A_KM, B_KM>(S, S + OPSZ, m0, n0, acc);
      }
      if (++qq < nchunks) continue;
      // ---- epilogue of tile (i, j)
      if (active) {
        if (MODE == TK_PANEL) {
          double* C = Pm + (long)i * NB * NB;
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              double2 v; v.x = -acc[a][b][0]; v.y = -acc[a][b][1];
              *reinterpret_cast<double2*>(C + (long)(m0 + 8 * a + r) * NB + n0 + 8 * b + c) = v;
            }
        } else {
          double* C = Am + ((long)i * NB) * g.ld + (long)j * NB;
          // the mirror tile: Z(j,i) = Z(i,j)^T, or the upper-right quarter of the symmetric diagonal tile
          const bool mirror = (MODE == TK_COL) || (MODE == TK_DIAG && m0 == NB / 2 && n0 == 0);
          double* U = Am + ((long)j * NB) * g.ld + (long)i * NB;
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const int row = m0 + 8 * a + r, col = n0 + 8 * b + c;
              double2* p = reinterpret_cast<double2*>(C + (long)row * g.ld + col);
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
              if (mirror) {
                U[(long)col * g.ld + row] = v.x;
                U[(long)(col + 1) * g.ld + row] = v.y;
              }
            }
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
      qq = 0;
      ++tt;
    }
    cp_async_wait<0>();
    __syncthreads();   // the ring is reused by the next matrix
  }
}

// tiles of one matrix are split over gx CTAs; enough CTAs for ~8 waves of the 148 x 3 resident slots
template <int MODE>
int launch_ll(const LLArgs& g0, int ntiles, cudaStream_t st, long* launches) {
  if (ntiles <= 0 || g0.batch <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    NMGP_CUDA_TRY(cudaFuncSetAttribute(panel_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LL_SMEM));
    configured = true;
  }
  LLArgs g = g0;
  g.ntiles = ntiles;
  const int gy = g.batch < 65535 ? g.batch : 65535;
  int gx = (148 * 3 * 8 + gy - 1) / gy;
  if (gx > ntiles) gx = ntiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, gy);
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
  g.ntiles = 0;
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
