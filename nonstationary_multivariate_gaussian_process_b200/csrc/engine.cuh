// Batched blocked FP64 Cholesky / inverse engine (tile-task kernels on DMMA).  See engine.cu.
#pragma once
#include "common.cuh"

namespace nmgp {

// A batch of symmetric matrices in the engine's padded block layout:
//   A      [batch][nP][nP] row-major, nP = Kt*NB >= n; rows/cols >= n hold the identity, so the padded
//          matrix is diag(Sigma, I): its factor is diag(L, I), its inverse diag(Sigma^-1, I), and no kernel
//          needs edge predication.
//   Dinv   [batch][Kt][2][NB*NB]  W_kk = inverse of the diagonal block of L (full tile, zero upper) and its transpose
//          W_kk^T, written by the diagonal-block kernel.  Both exist so that every operand of the left-looking
//          engine is K-MAJOR (row = output index, column = summation index): one TMA box shape, one swizzle.
struct BlockBatch {
  double* A = nullptr;
  double* A2 = nullptr;      // [batch][nP][nP]  second matrix buffer of the stable left-looking inverse (W^T tiles)
  double* Dinv = nullptr;
  double* Pbuf = nullptr;    // [batch][Kt][NB*NB]  panel side buffer of the left-looking inverse: P(c)^T tiles (zero-initialised once)
  long strideP() const { return (long)Kt * NB * NB; }
  const void* maps = nullptr;  // TMA descriptors of (A, A2, Dinv, Pbuf) for the FULL workspace (engine_maps_create); owned by the plan
  double* logdet = nullptr;  // [batch]
  int* info = nullptr;       // [batch]
  // [batch] smallest / largest pivot (diagonal entry of L) seen by the diagonal-block steps; optional.  Their squared ratio
  // is the conditioning indicator the guarded inverse uses to route a matrix (engine_potri_ll_guarded).
  double* pivmin = nullptr;
  double* pivmax = nullptr;
  int* sched = nullptr;      // [2] zero-initialised scheduler word of the left-looking engine's dynamic tile queue; optional
  int n = 0;                 // logical dimension
  int nP = 0;                // padded dimension (multiple of NB)
  int Kt = 0;                // blocks per side
  int NB = 64;
  int batch = 0;
  long strideA() const { return (long)nP * nP; }
  long strideD() const { return (long)Kt * 2 * NB * NB; }
};

constexpr int kNB = 64;

inline int padded_dim(int n, int NB = kNB) { return (int)round_up(n, NB); }

// In place: lower triangle of every A <- L;  logdet <- log det;  info <- 0 or failing pivot (1-based).
// stable_panel: solve the panel by substitution instead of multiplying by the inverse of the diagonal block
// (slower; for matrices with cond >> 1e6 such as the GP-prior covariances).
int engine_potrf(const BlockBatch& b, cudaStream_t st, long* launches, bool stable_panel = false);
// After potrf: strictly-lower blocks of A <- blocks of W = L^-1 (its diagonal blocks are already in Dinv).
int engine_trtri(const BlockBatch& b, cudaStream_t st, long* launches);
// After potrf: A <- inverse (both triangles filled).  trtri + lauum + symmetrize.
int engine_potri(const BlockBatch& b, cudaStream_t st, long* launches);

// One diagonal-block step (factor A(k,k), W_kk = L_kk^-1 into Dinv, log det, info) -- shared by both engines.
// pdl: programmatic dependent launch behind the previous kernel of `st` (common.cuh); only the tensor-pipe kernel honours it
int engine_diag_step(const BlockBatch& b, int k, cudaStream_t st, long* launches, bool accurate = false, bool pdl = false);

// 128-wide diagonal step for a handful of large matrices: blocks k and k+1 in ONE launch (both diagonal factorisations, the
// tile between them, W21 of the 128 x 128 inverse into Pbuf slot k) followed by the panel below both block columns.
int engine_diag128_step(const BlockBatch& b, int k, cudaStream_t st, long* launches);

// Left-looking potrf / Takahashi inverse for large batches of mid-size matrices (engine_ll.cu).  Same results layout as
// engine_potrf / engine_potri.  Requires b.Pbuf, and A's padding and upper block triangle to hold finite values
// (the plan zero-fills the workspace once).
int engine_potrf_ll(const BlockBatch& b, cudaStream_t st, long* launches);
int engine_potri_ll(const BlockBatch& b, cudaStream_t st, long* launches);
// Same result as engine_potri_ll by W = L^-1 (row-wise, all rows in parallel) and Z = W^T W (all tiles in parallel): three
// launches, no recursion over block columns, stable for any size.  Requires b.A2 (finite-initialised).
int engine_potri_ll_stable(const BlockBatch& b, cudaStream_t st, long* launches);
// Same result with W = L^-1 formed level by level (two launches of independent long-K tiles per level): for a few large
// matrices.  Accepts the factor of either potrf engine.  Requires b.A2 and b.Pbuf (tensor maps).
int engine_potri_ll_recursive(const BlockBatch& b, cudaStream_t st, long* launches);
// Trailing update A(i,j) -= sum_{c in [kb0, kb0+nkb)} A(i,c) A(j,c)^T over the lower tiles i >= j, ja <= j < jb (jb = 0: all
// remaining block columns) on the TMA-ring kernel: the updates of the right-looking potrf.  Needs b.maps or the buffers to
// build them (A, Dinv, Pbuf).
int engine_syrk_update_ll(const BlockBatch& b, int ja, int jb, int kb0, int nkb, cudaStream_t st, long* launches,
                          bool pdl = false);
// TMA descriptors of a workspace, built once (the maps cover the whole workspace, so they serve every sub-batch of it).
// `out` receives nullptr when the batch has no panel buffer.  engine_*_ll build a temporary set when b.maps is null.
int engine_maps_create(const BlockBatch& b, void** out);
void engine_maps_destroy(void* maps);
// The production inverse for thousands of mid-size matrices: every matrix takes the Takahashi sweep (fewest bytes) unless its
// factor says it is ill-conditioned -- (min pivot / max pivot)^2 < threshold(Kt) -- in which case the backward-stable W^T W
// path handles it; both launch sequences are queued, each kernel skips the matrices of the other (decided on the device: no
// host synchronisation, graph-capturable).  Measured (profiles/r02_takahashi_stress.txt): the sweep's gradient error grows
// from 4e-13 at noise variance e^-4 to 4e-9 at e^-8 and 5e-3 at e^-10 for n = 600 (10 block columns), and ~500x per two more
// block columns; the W^T W path stays below 2e-7 everywhere.  Requires b.pivmin / b.pivmax (written by potrf), b.A2, b.Pbuf.
int engine_potri_ll_guarded(const BlockBatch& b, cudaStream_t st, long* launches);
// largest number of 64-blocks per side for which the Takahashi recursion of engine_potri_ll is used (see api.cu)
constexpr int kTakahashiMaxBlocks = 16;

// Heuristic: the left-looking path launches (Kt - k) * batch CTAs per block column, each a long K loop; it needs a few
// waves of them (148 SMs x 4 resident CTAs) to keep the tensor pipes busy: thousands of mid-size matrices.  For a few dozen
// LARGE matrices the right-looking path with panels and look-ahead (engine.cu), whose updates run on the same TMA-ring kernel
// since the end of round 2, is ahead.  Measured on B200, potrf only, right- vs left-looking (profiles/r02_potrf_crossover.txt):
// n = 2048 x 64: 7.2 vs 8.0 ms; n = 4096 x 32: 25.9 vs 29.5, x 64: 50.9 vs 51.4; n = 8192 x 16: 91 vs 111; n = 16 384 x 8:
// 355 vs 431, x 23: 1022 vs 1067 ms.  (Round 1, before that kernel: n = 16 384 batch 8 equal, n = 2048 batch 64 left-looking
// 7.9 vs 9.0 ms -- profiles/r01_potrf_panel_width.txt.)
inline bool prefer_left_looking(const BlockBatch& b) {
  return b.Pbuf != nullptr && (b.batch >= 96 || (b.Kt < 32 && (long)b.batch * b.Kt >= 2048));
}

}  // namespace nmgp
