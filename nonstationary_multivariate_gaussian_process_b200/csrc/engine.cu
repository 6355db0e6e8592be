// Batched blocked FP64 Cholesky + inverse on B200, written as "tile tasks".
//
// Every O(n^3) step of potrf / trtri / lauum is a rank-NB update  C(i,j) = beta*C(i,j) + alpha * opA * opB
// on NB x NB tiles of the padded matrix, one CTA per (tile, matrix), computed with DMMA.8x8x4
// (mma.sync m8n8k4 f64 -- the FP64 tensor-core instruction of sm_100a; tcgen05 has no f64 kind).
// A launch covers the tiles of ONE block step for ALL matrices of the batch, so a 10 000-subject
// sweep costs the same number of launches as one subject.  The only non-GEMM work is the NB x NB
// diagonal block (factor + inverse in shared memory).
//
// Replaces, for the whole batch at once: torch.inverse + torch.logdet (LU) at Utility/logpos.py:352-353,
// torch.symeig at Utility/distributions.py:37,40 (block formulation) and the Cholesky
// inside MultivariateNormal at logpos.py:274,279,358,365.
#include "engine.cuh"
#include "tile_mma.cuh"

#include <cstdlib>

namespace nmgp {

namespace {

constexpr int NB = kNB;          // 64
using tile::LDS;
using tile::TILE_THREADS;
using tile::load_tile;
using tile::load_tiles2;
using tile::warp_mma;
constexpr size_t TILE_SMEM = 2ull * NB * LDS * sizeof(double);

enum Mode : int {
  POTRF_PANEL = 0,  // A(i,k)  <- A(i,k) * W_kk^T                      i > k
  POTRF_SYRK = 1,   // A(i,j) -= A(i,k) * A(j,k)^T                     k < j <= i
  TRTRI_ROW = 2,    // A(c,j)  <- -W_cc * A(c,j)                       j < c
  TRTRI_UPD = 3,    // A(i,j) += A(i,c) * A(c,j)                       i > c, j < c
  TRTRI_PANEL = 4,  // A(i,c)  <- A(i,c) * W_cc                        i > c
  LAUUM_UPD = 5,    // A(i,j) += A(r,i)^T * A(r,j)                     j <= i < r
  LAUUM_ROW = 6,    // A(r,j)  <- W_rr^T * A(r,j) (j<r);  A(r,r) <- W_rr^T W_rr
  // POTRF_SYRK split for the look-ahead: the next block column alone, and everything to the right of it
  POTRF_SYRK_COL = 7,   // A(i,k+1) -= A(i,k) * A(k+1,k)^T               i >= k+1
  POTRF_SYRK_REST = 8,  // A(i,j)   -= A(i,k) * A(j,k)^T                 k+2 <= j <= i
};

struct EngineArgs {
  double* A;
  double* Dinv;
  long strideA, strideD;
  int ld, Kt, batch, step;
};

__device__ __forceinline__ void tri_decode(int t, int& a, int& b) {
  // t = a(a+1)/2 + b, 0 <= b <= a
  a = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((long)a * (a + 1) / 2 > t) --a;
  while ((long)(a + 1) * (a + 2) / 2 <= t) ++a;
  b = t - a * (a + 1) / 2;
}

// CHAIN: the launch sits on the critical chain of a large factorisation with a handful of CTAs: round trips, not bytes or
// occupancy, are its time (both operands in one global round trip, 64 more registers)
template <int MODE, bool CHAIN = false>
__global__ void __launch_bounds__(TILE_THREADS) tile_kernel(EngineArgs g) {
  extern __shared__ __align__(16) double smem[];
  double* SA = smem;
  double* SB = smem + NB * LDS;

  const int s = g.step;
  const int t = blockIdx.x;
  int i, j;          // output block coordinates
  int ai, aj;        // block coordinates of operand A inside the matrix (or -1: Dinv[step])
  int bi, bj;        // same for operand B
  if (MODE == POTRF_PANEL) {
    i = s + 1 + t; j = s; ai = i; aj = s; bi = -1; bj = -1;
  } else if (MODE == POTRF_SYRK) {
    int a, b; tri_decode(t, a, b);
    i = s + 1 + a; j = s + 1 + b; ai = i; aj = s; bi = j; bj = s;
  } else if (MODE == POTRF_SYRK_COL) {
    i = s + 1 + t; j = s + 1; ai = i; aj = s; bi = j; bj = s;
  } else if (MODE == POTRF_SYRK_REST) {
    int a, b; tri_decode(t, a, b);
    i = s + 2 + a; j = s + 2 + b; ai = i; aj = s; bi = j; bj = s;
  } else if (MODE == TRTRI_ROW) {
    i = s; j = t; ai = -1; aj = -1; bi = s; bj = j;
  } else if (MODE == TRTRI_UPD) {
    const int a = t / s, b = t % s;
    i = s + 1 + a; j = b; ai = i; aj = s; bi = s; bj = j;
  } else if (MODE == TRTRI_PANEL) {
    i = s + 1 + t; j = s; ai = i; aj = s; bi = -1; bj = -1;
  } else if (MODE == LAUUM_UPD) {
    int a, b; tri_decode(t, a, b);
    i = a; j = b; ai = s; aj = i; bi = s; bj = j;
  } else {  // LAUUM_ROW
    i = s; j = t; ai = -1; aj = -1;
    if (j < s) { bi = s; bj = j; } else { bi = -1; bj = -1; }
  }
  constexpr bool SYRK = (MODE == POTRF_SYRK || MODE == POTRF_SYRK_COL || MODE == POTRF_SYRK_REST);
  constexpr bool A_KM = (MODE == POTRF_PANEL || SYRK || MODE == TRTRI_ROW || MODE == TRTRI_UPD || MODE == TRTRI_PANEL);
  constexpr bool B_KM = (MODE == POTRF_PANEL || SYRK);
  constexpr double alpha = (SYRK || MODE == TRTRI_ROW) ? -1.0 : 1.0;
  constexpr bool accumulate = (SYRK || MODE == TRTRI_UPD || MODE == LAUUM_UPD);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (warp >> 1) * 32, n0 = (warp & 1) * 32;

  if (CHAIN) {   // may be a programmatic dependent of the diagonal-block kernel (common.cuh)
    pdl_launch_dependents();
    pdl_wait();
  }
  for (int mat = blockIdx.y; mat < g.batch; mat += gridDim.y) {
    double* Am = g.A + (long)mat * g.strideA;
    const double* Dm = g.Dinv + (long)mat * g.strideD + (long)s * 2 * NB * NB;   // slot 0: W_kk
    const double* srcA = (ai < 0) ? Dm : Am + ((long)ai * NB) * g.ld + (long)aj * NB;
    const double* srcB = (bi < 0) ? Dm : Am + ((long)bi * NB) * g.ld + (long)bj * NB;
    if (CHAIN) {
      load_tiles2(SA, srcA, ai < 0 ? NB : g.ld, SB, srcB, bi < 0 ? NB : g.ld);
    } else {
      load_tile(SA, srcA, ai < 0 ? NB : g.ld);
      load_tile(SB, srcB, bi < 0 ? NB : g.ld);
    }
    __syncthreads();

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    warp_mma<A_KM, B_KM>(SA, SB, m0, n0, acc);

    double* C = Am + ((long)i * NB) * g.ld + (long)j * NB;
    const int r = lane >> 2, c = 2 * (lane & 3);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        double2* p = reinterpret_cast<double2*>(C + (long)(m0 + 8 * a + r) * g.ld + n0 + 8 * b + c);
        double2 v;
        if (accumulate) {
          v = *p;
          v.x += alpha * acc[a][b][0];
          v.y += alpha * acc[a][b][1];
        } else {
          v.x = alpha * acc[a][b][0];
          v.y = alpha * acc[a][b][1];
        }
        *p = v;
      }
    __syncthreads();  // shared tiles are reused by the next matrix
  }
}

// Mirror the lower block triangle into the upper one: A(j,i) = A(i,j)^T for i > j.
__global__ void __launch_bounds__(256) symmetrize_kernel(EngineArgs g) {
  __shared__ double T[NB][NB + 1];
  int a, b;
  tri_decode(blockIdx.x, a, b);
  const int i = a + 1, j = b;  // strictly lower blocks
  for (int mat = blockIdx.y; mat < g.batch; mat += gridDim.y) {
    double* Am = g.A + (long)mat * g.strideA;
    const double* src = Am + ((long)i * NB) * g.ld + (long)j * NB;
    double* dst = Am + ((long)j * NB) * g.ld + (long)i * NB;
    for (int idx = threadIdx.x; idx < NB * NB; idx += 256) {
      const int r = idx / NB, c = idx % NB;
      T[r][c] = src[(long)r * g.ld + c];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < NB * NB; idx += 256) {
      const int r = idx / NB, c = idx % NB;
      dst[(long)r * g.ld + c] = T[c][r];
    }
    __syncthreads();
  }
}

// Trailing update by a PANEL of nkb block columns at once:  A(i,j) -= sum_{kb} A(i,kb0+kb) A(j,kb0+kb)^T, accumulated in
// registers over the panel, so the C tile is read and written once per panel instead of once per block column (the rank-64
// update is memory-bound: 4 flop/B).  TRI: all tiles i >= j >= ja (tri_decode);  !TRI: the columns ja <= j < jb only.
template <bool TRI>
__global__ void __launch_bounds__(TILE_THREADS) syrk_wide_kernel(EngineArgs g, int ja, int jb, int kb0, int nkb) {
  extern __shared__ __align__(16) double smem[];
  double* SA = smem;
  double* SB = smem + NB * LDS;
  int i, j;
  if (TRI) {
    int a, b;
    tri_decode(blockIdx.x, a, b);
    i = ja + a; j = ja + b;
  } else {
    int t = blockIdx.x;
    j = ja;
    while (j < jb - 1 && t >= g.Kt - j) { t -= g.Kt - j; ++j; }
    i = j + t;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (warp >> 1) * 32, n0 = (warp & 1) * 32;
  for (int mat = blockIdx.y; mat < g.batch; mat += gridDim.y) {
    double* Am = g.A + (long)mat * g.strideA;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    double* C = Am + ((long)i * NB) * g.ld + (long)j * NB;
    const int r = lane >> 2, c = 2 * (lane & 3);
    // !TRI = the few-tile launches on the critical chain (in-panel and NEXT updates): the C tile is fetched under the
    // products and both operands of a k-block come in one global round trip; TRI (the bulk update) keeps its registers for
    // occupancy
    double2 cpre[TRI ? 1 : 4][TRI ? 1 : 4];
    if (!TRI) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          cpre[TRI ? 0 : a][TRI ? 0 : b] = *reinterpret_cast<const double2*>(C + (long)(m0 + 8 * a + r) * g.ld + n0 + 8 * b + c);
    }
    for (int kb = 0; kb < nkb; ++kb) {
      const double* pa = Am + ((long)i * NB) * g.ld + (long)(kb0 + kb) * NB;
      const double* pb = Am + ((long)j * NB) * g.ld + (long)(kb0 + kb) * NB;
      if (!TRI) {
        load_tiles2(SA, pa, g.ld, SB, pb, g.ld);
      } else {
        load_tile(SA, pa, g.ld);
        load_tile(SB, pb, g.ld);
      }
      __syncthreads();
      warp_mma<true, true>(SA, SB, m0, n0, acc);
      __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        double2* p = reinterpret_cast<double2*>(C + (long)(m0 + 8 * a + r) * g.ld + n0 + 8 * b + c);
        double2 v = TRI ? *p : cpre[TRI ? 0 : a][TRI ? 0 : b];
        v.x -= acc[a][b][0];
        v.y -= acc[a][b][1];
        *p = v;
      }
  }
}

// Backward-stable panel solve  X L_kk^T = A(i,k)  by substitution (one thread per row), used instead of the
// multiply-by-inverse POTRF_PANEL when the matrix is ill-conditioned (GP-prior covariances, cond ~ 1e10):
// multiplying by an explicit inverse of the diagonal block inflates the backward error by cond(L_kk).
__global__ void __launch_bounds__(NB) trsm_panel_kernel(EngineArgs g) {
  extern __shared__ __align__(16) double smem[];
  double (*Ls)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smem);
  double (*Xs)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(smem + NB * (NB + 1));
  const int k = g.step;
  const int i = k + 1 + blockIdx.x;
  const int r = threadIdx.x;
  for (int mat = blockIdx.y; mat < g.batch; mat += gridDim.y) {
    double* Am = g.A + (long)mat * g.strideA;
    const double* Lkk = Am + ((long)k * NB) * g.ld + (long)k * NB;
    double* X = Am + ((long)i * NB) * g.ld + (long)k * NB;
    for (int idx = threadIdx.x; idx < NB * NB; idx += NB) {
      const int rr = idx / NB, cc = idx % NB;
      Ls[rr][cc] = Lkk[(long)rr * g.ld + cc];
      Xs[rr][cc] = X[(long)rr * g.ld + cc];
    }
    __syncthreads();
    for (int c = 0; c < NB; ++c) {
      double sacc = Xs[r][c];
      for (int c2 = 0; c2 < c; ++c2) sacc -= Xs[r][c2] * Ls[c][c2];
      Xs[r][c] = sacc / Ls[c][c];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < NB * NB; idx += NB) {
      const int rr = idx / NB, cc = idx % NB;
      X[(long)rr * g.ld + cc] = Xs[rr][cc];
    }
    __syncthreads();
  }
}

template <int MODE, bool CHAIN = false>
int launch_tiles(const EngineArgs& g, int ntiles, cudaStream_t st, long* launches, bool pdl = false) {
  if (ntiles <= 0 || g.batch <= 0) return 0;
  NMGP_SMEM_ATTR_PER_DEVICE((tile_kernel<MODE, CHAIN>), TILE_SMEM);
  dim3 grid(ntiles, g.batch < 65535 ? g.batch : 65535);
  NMGP_CUDA_TRY(launch_kernel_pdl(tile_kernel<MODE, CHAIN>, grid, dim3(TILE_THREADS), TILE_SMEM, st, pdl && CHAIN, g));
  NMGP_CUDA_TRY(cudaGetLastError());
  if (launches) ++*launches;
  return 0;
}

EngineArgs make_args(const BlockBatch& b) {
  EngineArgs g;
  g.A = b.A; g.Dinv = b.Dinv; g.strideA = b.strideA(); g.strideD = b.strideD();
  g.ld = b.nP; g.Kt = b.Kt; g.batch = b.batch; g.step = 0;
  return g;
}

}  // namespace

namespace {
int launch_panel(const BlockBatch& b, const EngineArgs& g, int r, bool stable_panel, cudaStream_t st, long* launches,
                 bool chain = false, bool pdl = false) {
  if (!stable_panel) return chain ? launch_tiles<POTRF_PANEL, true>(g, r, st, launches, pdl) : launch_tiles<POTRF_PANEL>(g, r, st, launches);
  if (r <= 0) return 0;
  dim3 pg(r, b.batch < 65535 ? b.batch : 65535);
  constexpr size_t kTrsmSmem = 2ull * NB * (NB + 1) * sizeof(double);
  NMGP_SMEM_ATTR_PER_DEVICE(trsm_panel_kernel, kTrsmSmem);
  trsm_panel_kernel<<<pg, NB, kTrsmSmem, st>>>(g);
  NMGP_CUDA_TRY(cudaGetLastError());
  if (launches) ++*launches;
  return 0;
}

template <bool TRI>
int launch_syrk_wide(const EngineArgs& g, int ja, int jb, int kb0, int nkb, cudaStream_t st, long* launches) {
  int ntiles = 0;
  if (TRI) { const int r = g.Kt - ja; ntiles = r * (r + 1) / 2; }
  else for (int j = ja; j < jb; ++j) ntiles += g.Kt - j;
  if (ntiles <= 0 || nkb <= 0 || g.batch <= 0) return 0;
  NMGP_SMEM_ATTR_PER_DEVICE(syrk_wide_kernel<TRI>, TILE_SMEM);
  dim3 grid(ntiles, g.batch < 65535 ? g.batch : 65535);
  syrk_wide_kernel<TRI><<<grid, TILE_THREADS, TILE_SMEM, st>>>(g, ja, jb, kb0, nkb);
  NMGP_CUDA_TRY(cudaGetLastError());
  if (launches) ++*launches;
  return 0;
}

// Right-looking potrf for a handful of large matrices: panels of W block columns with a look-ahead.
//
// Every block column is a chain of dependent kernels -- diagonal block (19 us for one matrix: 64 dependent pivots), panel,
// update of what the next diagonal block needs -- and the chain, not the arithmetic, sets the time (n = 5000 started at
// 79 steps x 74 us, profiles/r01_c3_chain.txt).  Two measures:
//  * the trailing matrix is updated once per PANEL (rank 64 W, accumulated in registers) instead of once per block column:
//    the rank-64 update is memory-bound (4 flop/B), the C tiles dominate its traffic;
//  * look-ahead: the chain (diag -> panel -> in-panel update, then the update of the NEXT panel's columns) runs on a
//    HIGH-priority helper stream, the rest of the trailing update (REST) on the caller's stream beside the next panel's
//    chain.  (The caller's stream usually has the lowest priority already, which is why the chain, not REST, moves.)
// Orderings: REST(P) after panel P is complete [evP]; NEXT(P+1) after REST(P) [evR], because both update the columns of
// panel P+2; the caller's stream finally waits for the helper.
int potrf_lookahead(const BlockBatch& b, cudaStream_t st, long* launches, bool stable_panel) {
  EngineArgs g = make_args(b);
  const int Kt = b.Kt;
  int W = Kt >= 128 ? 8 : 2;   // measured (profiles/r01_potrf_panel_width.txt): n = 5000 is chain-bound for any W, n = 16 384 gains up to W = 8
  if (const char* ev = getenv("NMGP_POTRF_W")) { const int w = atoi(ev); if (w >= 1 && w <= 16) W = w; }   // A/B timing
  // 128-wide diagonal steps (diag.cu:diag128_kernel; needs the side buffer for W21 and the multiply-by-inverse panel): half
  // the chain links and a third fewer launches, but NOT faster -- the two dependent 64 x 64 Choleskys (17.5 us each) dominate the
  // link either way, and the four small products between them run on one SM (measured, profiles/r02_diag128.txt: n = 5000
  // 4.11 vs 4.05 ms, n = 16 384 53.4 vs 53.8 ms).  Kept for A/B timing (NMGP_DIAG128=1), off by default.
  static const bool want_wide = getenv("NMGP_DIAG128") != nullptr && atoi(getenv("NMGP_DIAG128")) != 0;
  const bool wide_diag = !stable_panel && b.Pbuf != nullptr && want_wide && W >= 2;
  // helper stream and events: created once per (host thread, device) and reused by every call
  struct Helper {
    int device = -1;
    cudaStream_t crit = nullptr;
    cudaEvent_t evP[2] = {nullptr, nullptr}, evR[2] = {nullptr, nullptr}, ev0 = nullptr;
  };
  static thread_local Helper hp[16];
  int dev = 0;
  NMGP_CUDA_TRY(cudaGetDevice(&dev));
  Helper& H = hp[dev & 15];
  if (H.device != dev) {
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    bool made = cudaStreamCreateWithPriority(&H.crit, cudaStreamNonBlocking, greatest) == cudaSuccess &&
                cudaEventCreateWithFlags(&H.ev0, cudaEventDisableTiming) == cudaSuccess;
    for (int e = 0; e < 2 && made; ++e)
      made = cudaEventCreateWithFlags(&H.evP[e], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&H.evR[e], cudaEventDisableTiming) == cudaSuccess;
    if (!made) { set_last_error("engine_potrf: creating the look-ahead stream / events failed"); return -2; }
    H.device = dev;
  }
  cudaStream_t crit = H.crit;
  cudaEvent_t *evP = H.evP, *evR = H.evR, ev0 = H.ev0;
  int rc = 0;
  auto cu = [&](cudaError_t e) {
    if (e != cudaSuccess && rc == 0) { rc = -2; set_last_error(std::string("engine_potrf look-ahead: ") + cudaGetErrorString(e)); }
    return rc == 0;
  };
  if (rc == 0) { cu(cudaEventRecord(ev0, st)); cu(cudaStreamWaitEvent(crit, ev0, 0)); }   // after whatever built the matrices
  // the updates A(i,j) -= sum_c A(i,c) A(j,c)^T, ja <= j < jb, i >= j: on the TMA-ring kernel of engine_ll.cu where the batch has
  // (or can build) tensor maps -- n = 16 384: potrf 53.0 -> 48.8 ms with the bulk update there (27 -> 30 TFLOP/s) -- else the plain
  // tile kernels.  NMGP_REST_LL: 0 = plain kernels, 1 = bulk update only, 2 (default) = chain updates too.
  static const int upd_ll = getenv("NMGP_REST_LL") ? atoi(getenv("NMGP_REST_LL")) : 2;
  const bool can_ll = !stable_panel && (b.maps || (b.Dinv && b.Pbuf));
  // The kernels of the chain as programmatic dependents of one another (scheduled while their predecessor still runs,
  // griddepcontrol.wait before they read: common.cuh): hides most of the launch gap between the three dependent launches of a
  // block column.  Measured (profiles/r02_potrf_crossover.txt): it pays where the chain IS the time -- one matrix of n = 3000:
  // 1.63 -> 1.37 ms, n = 5000: 3.07 -> 2.87 ms -- and costs 2-6 % where a bulk update runs beside the chain (n = 8000: 7.48 ->
  // 7.62 ms, n = 16 384: 47.0 -> 48.4 ms, 4 x n = 4096: 4.03 -> 4.28 ms: the early-scheduled CTAs wait in slots the update
  // could use).  Hence: batch * Kt <= 96.  NMGP_PDL = 0 / 1 forces it off / on.
  static const int want_pdl = getenv("NMGP_PDL") ? atoi(getenv("NMGP_PDL")) : -1;
  const bool pdl = can_ll && upd_ll >= 2 && !stable_panel && (want_pdl >= 0 ? want_pdl != 0 : (long)b.batch * Kt <= 96);
  auto update = [&](int ja, int jb, int kb0, int nkb, cudaStream_t s, bool bulk, bool dep = false) -> int {
    if (can_ll && upd_ll >= (bulk ? 1 : 2)) return engine_syrk_update_ll(b, ja, jb, kb0, nkb, s, launches, dep && pdl);
    return bulk ? launch_syrk_wide<true>(g, ja, jb, kb0, nkb, s, launches) : launch_syrk_wide<false>(g, ja, jb, kb0, nkb, s, launches);
  };
  bool rest_pending = false;
  int last_e = -1, panel = 0;
  for (int p0 = 0; rc == 0 && p0 < Kt; p0 += W, ++panel) {
    const int p1 = p0 + W < Kt ? p0 + W : Kt, e = panel & 1;
    for (int k = p0; rc == 0 && k < p1;) {
      g.step = k;
      if (wide_diag && k + 1 < p1) {
        // two block columns per chain link: 128 x 128 diagonal block + the panel below it, then the rest of this panel (rank 128)
        if ((rc = engine_diag128_step(b, k, crit, launches))) break;
        if (k + 2 < p1) rc = launch_syrk_wide<false>(g, k + 2, p1, k, 2, crit, launches);
        k += 2;
        continue;
      }
      // (the first diagonal block follows the caller's work, the one after a panel boundary follows NEXT: a kernel either way,
      //  but the very first launch of the chain has an event wait in front of it)
      if ((rc = engine_diag_step(b, k, crit, launches, stable_panel, pdl && k > 0))) break;
      if ((rc = launch_panel(b, g, Kt - k - 1, stable_panel, crit, launches, true, pdl))) break;
      if (k + 1 < p1) rc = update(k + 1, p1, k, 1, crit, false, true);                    // rest of this panel, rank 64
      ++k;
    }
    if (rc != 0 || !cu(cudaEventRecord(evP[e], crit))) break;
    last_e = e;
    if (p1 >= Kt) break;
    if (rest_pending && !cu(cudaStreamWaitEvent(crit, evR[e ^ 1], 0))) break;              // REST(P-1) done
    rest_pending = false;
    const int n1 = p1 + W < Kt ? p1 + W : Kt;
    if ((rc = update(p1, n1, p0, p1 - p0, crit, false))) break;                           // NEXT: the next panel's columns
    if (n1 < Kt) {
      if (!cu(cudaStreamWaitEvent(st, evP[e], 0))) break;
      if ((rc = update(n1, Kt, p0, p1 - p0, st, true))) break;                              // REST: the bulk of the update
      if (!cu(cudaEventRecord(evR[e], st))) break;
      rest_pending = true;
    }
  }
  // the caller's stream continues only after the helper's chain (also on errors, so that nothing is left racing)
  if (last_e >= 0) {
    cudaEventRecord(evP[last_e], crit);
    cudaStreamWaitEvent(st, evP[last_e], 0);
  }
  return rc;
}
}  // namespace

int engine_potrf(const BlockBatch& b, cudaStream_t st, long* launches, bool stable_panel) {
  if (b.batch <= 0) return 0;
  if (b.NB != NB || b.nP != b.Kt * NB) { set_last_error("engine_potrf: bad block layout"); return -1; }
  // look-ahead pays when a block step does not fill the GPU for long: few matrices, several block columns
  if (b.Kt >= 8 && ((long)b.batch * b.Kt < 2048 || (b.batch < 96 && b.Kt >= 32) || getenv("NMGP_FORCE_LOOKAHEAD")))
    return potrf_lookahead(b, st, launches, stable_panel);
  EngineArgs g = make_args(b);
  for (int k = 0; k < b.Kt; ++k) {
    g.step = k;
    NMGP_TRY(engine_diag_step(b, k, st, launches, stable_panel));
    const int r = b.Kt - k - 1;
    NMGP_TRY(launch_panel(b, g, r, stable_panel, st, launches));
    NMGP_TRY(launch_tiles<POTRF_SYRK>(g, r * (r + 1) / 2, st, launches));
  }
  return 0;
}

int engine_trtri(const BlockBatch& b, cudaStream_t st, long* launches) {
  if (b.batch <= 0) return 0;
  EngineArgs g = make_args(b);
  // W = L^-1 (strictly-lower blocks in A, diagonal blocks stay in Dinv)
  for (int c = 0; c < b.Kt; ++c) {
    g.step = c;
    const int below = b.Kt - c - 1;
    NMGP_TRY(launch_tiles<TRTRI_ROW>(g, c, st, launches));
    NMGP_TRY(launch_tiles<TRTRI_UPD>(g, below * c, st, launches));
    NMGP_TRY(launch_tiles<TRTRI_PANEL>(g, below, st, launches));
  }
  return 0;
}

int engine_potri(const BlockBatch& b, cudaStream_t st, long* launches) {
  if (b.batch <= 0) return 0;
  NMGP_TRY(engine_trtri(b, st, launches));
  EngineArgs g = make_args(b);
  // A = W^T W (lower block triangle, diagonal blocks complete)
  for (int r = 0; r < b.Kt; ++r) {
    g.step = r;
    NMGP_TRY(launch_tiles<LAUUM_UPD>(g, r * (r + 1) / 2, st, launches));
    NMGP_TRY(launch_tiles<LAUUM_ROW>(g, r + 1, st, launches));
  }
  if (b.Kt > 1) {
    dim3 grid(b.Kt * (b.Kt - 1) / 2, b.batch < 65535 ? b.batch : 65535);
    symmetrize_kernel<<<grid, 256, 0, st>>>(g);
    NMGP_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
  }
  return 0;
}

}  // namespace nmgp
