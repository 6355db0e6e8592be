// Left-looking batched Cholesky + "Takahashi" inverse for many mid-size matrices (the 10 000 x (n=600) regime) and for
// batches of large ones.  FP64 tensor cores (DMMA.8x8x4) fed by TMA.
//
// The right-looking tile tasks of engine.cu read AND write every trailing tile at every block step.  Here every output
// tile is produced ONCE: a CTA accumulates its 64 x 64 tile over the whole K range in registers, streaming the two
// operand row-panels through a shared-memory ring.
//
//   potrf, block column k:   A(i,k) <- A(i,k) - sum_{j<k} A(i,j) A(k,j)^T        (LL_UPDATE,  i >= k)
//                            diag64_kernel: A(k,k) = L L^T, W_kk = L^-1, W_kk^T, log det
//                            A(i,k) <- A(i,k) W_kk^T                              (LL_SOLVE,   i > k)
//   inverse, block column j descending (Z = Sigma^-1, stored full symmetric, in place over L):
//                            P(c)   = -L(c,j) W_jj   (stored transposed)          (TK_PANEL,   c > j, side buffer)
//                            Z(i,j) = sum_{c>j} Z(i,c) P(c),  Z(j,i) = Z(i,j)^T   (TK_COL,     i > j)
//                            Z(j,j) = W_jj^T W_jj + sum_{c>j} P(c)^T Z(c,j)       (TK_DIAG)
// Flops: n^3/3 + 2n^3/3, the same as potrf + potri; every operand tile is read once per use, C never re-read.
//
// Operand staging.  Every product is arranged so that BOTH operands are K-MAJOR tiles (row = output index, column =
// summation index): W_kk^T and P(c)^T are stored beside / instead of W_kk and P(c), and Z(c,j) is read through its
// mirror Z(j,c).  One TMA box shape (16 k x 64 rows = 64 rows of 128 bytes) then serves everything:
// cp.async.bulk.tensor.3d {k, row, matrix} with the hardware 128-byte swizzle, issued by ONE elected thread per chunk,
// completion signalled on an mbarrier -- no per-thread address arithmetic, no padding in shared memory (48 KB ring,
// 3 stages).  The DMMA fragments are read bank-conflict free from the swizzled rows by choosing WHICH four k's a k-step
// multiplies: {2s, 2s+1, 2s+8, 2s+9} for s = 0..3 (the order of the summation index is free as long as A and B agree).
// tools/dmma_pipe.cu measures this ring at 35.7 TFLOP/s (96 % of the 37.15 DMMA peak) against 34.6 for the cp.async
// ring with padded rows it replaces, bit-identical results.
//
// One CTA walks a LIST of output tiles of one matrix (tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of this launch) and
// streams all their k-chunks through ONE ring: the loads of the next tile are already in flight while the current
// tile's epilogue runs.  The epilogue works from registers only (the mirror of the symmetric inverse and P^T are
// direct transposed stores: 8 lanes write 64 contiguous bytes), so shared memory belongs to the ring at all times.
//
// Measured dead ends (profiles/README.md): prefetch.global.L2 of the operand tiles 3 k-blocks ahead (-2 %), KC = 32 x 2
// stages (-1 %), generating the covariance tile inside the LL_UPDATE epilogue instead of reading it (+8 ms in potrf for
// 6 ms saved in the build kernel), fusing the diagonal-block factorisation into the panel-solve kernel (+7 ms: a CTA
// slot idles in the pivot chain).
#include "engine.cuh"

#include <cuda.h>

#include <cstdint>
#include <atomic>
#include <cstdlib>
#include <new>

namespace nmgp {

namespace {

constexpr int NB = kNB;      // 64
constexpr int KC = 16;       // k-chunk per pipeline stage = one 128-byte swizzle row
constexpr int CPB = NB / KC; // chunks per k-block
#ifndef NMGP_LL_STAGES
#define NMGP_LL_STAGES 3
#endif
constexpr int STAGES = NMGP_LL_STAGES;
constexpr int OPB = NB * KC * 8;   // bytes of one operand chunk in shared memory (dense, swizzled): 8192
constexpr int THREADS = 128;
constexpr size_t LL_SMEM = (size_t)STAGES * 2 * OPB + 128 + 1024;   // ring + mbarriers (full, empty) + slack for 1024-byte alignment

enum LLMode : int { LL_UPDATE = 0, LL_SOLVE = 1, TK_PANEL = 2, TK_COL = 3, TK_DIAG = 4 };

struct LLArgs {
  double* A;
  double* A2;
  double* Pbuf;
  long strideA, strideP;
  int ld, Kt, batch, step, n8, ntiles;
  // routing of the guarded inverse: 0 = every matrix; 1 = only well-conditioned matrices (Takahashi kernels);
  // 2 = only ill-conditioned ones (W^T W kernels).  ill-conditioned: pivmin^2 < guard_thr * pivmax^2.
  int guard;
  double guard_thr;
  const double* pivmin;
  const double* pivmax;
  int* sched;   // [2] {next item, CTAs out of work}, zero between launches (the kernel resets them); nullptr = static grid
  // SYRK_UPD: updated tiles (i, j), i >= j, ua <= j < ujb (ujb = 0: up to Kt, the whole trailing triangle), k-block range
  // [ukb0, ukb0 + unkb)
  int ua = 0, ujb = 0, ukb0 = 0, unkb = 0;
};

__device__ __forceinline__ bool guard_skips(const LLArgs& g, int mat) {
  if (g.guard == 0) return false;
  const double a = g.pivmin[mat], b = g.pivmax[mat];
  const bool ill = a * a < g.guard_thr * b * b;
  return g.guard == 1 ? ill : !ill;
}

// ------------------------------------------------------------------------------------------------ TMA / mbarrier PTX
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, int parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one 16 (k) x 64 (rows) box of matrix / tile `c2`, landing as 64 swizzled 128-byte rows
__device__ __forceinline__ void tma_load_chunk(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                               unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

// acc(32 x 32 of this warp) += A(32 x 16) * B(32 x 16)^T for one chunk.  SA / SB point at row (m0 + lr) / (n0 + lr) of the
// operand chunks; off[s] is the lane's swizzled byte offset inside a 128-byte row for k-step s.
__device__ __forceinline__ void mma_chunk(const unsigned char* __restrict__ SA, const unsigned char* __restrict__ SB,
                                          const int (&off)[4], double (&acc)[4][4][2]) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = *reinterpret_cast<const double*>(SA + i * 1024 + off[s]);   // next 8-row group: 8 rows x 128 B
      b[i] = *reinterpret_cast<const double*>(SB + i * 1024 + off[s]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

// ---- panel_gemm_kernel
// Work items of one launch: (matrix, tile), w = matrix * ntiles + tile, the same k-chunk sequence for every item.
//  * Scheduling.  With a scheduler word (LLArgs::sched, plans) the grid is the set of co-resident CTAs and every CTA FETCHES
//    its next item with an atomic counter: dynamic balance like the hardware's CTA scheduler, but the TMA ring of a CTA never
//    drains between tiles and the CTA start-up (guard loads, barrier init, first TMA round trip: ~10 % of a one-tile CTA in the
//    source-level profile) is paid once per launch instead of once per tile.  Items of one matrix are fetched by neighbouring
//    CTAs at the same time, so the operand they share comes from HBM once.  The last CTA to run dry resets the counters.
//    Without a scheduler word (unit entry points): grid (gx, batch), CTA (bx, by) takes tiles bx, bx + gx, ... of matrix by.
//  * Inside a CTA the four warps only meet through the ring: `full[slot]` is completed by the TMA transaction, `empty[slot]`
//    by one arrival per warp, so a warp in its epilogue does not hold up the other three (no CTA-wide barrier in the loop).
//    The producer is thread 0: it issues chunk q + STAGES - 1 once `empty` says everyone is done with chunk q - 1, and
//    publishes the item of every tile it starts in a 4-entry shared ring that the consumers read after the tile's first
//    `full` wait; a negative entry ends the CTA.
// Measured alternatives (profiles/r02_engine_ab.txt; the code is in commit 92ea92f): STATIC round-robin deal over a persistent
// grid +3 % time, dedicated producer warp +1.5 % (3 CTAs/SM) / +20 % (4 CTAs/SM at 96 registers), 4 ring stages equal.
constexpr int PG_THREADS = THREADS;
constexpr int PG_CTAS = 4;

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(PG_THREADS, PG_CTAS) panel_gemm_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                         const __grid_constant__ CUtensorMap mapD,
                                                                         const __grid_constant__ CUtensorMap mapP,
                                                                         LLArgs g) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ int item_ring[4];
  // 1024-byte alignment by OFFSET arithmetic on the shared pointer (an integer round trip would make every operand fetch a
  // generic LD instead of LDS: the compiler loses the address space)
  unsigned char* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + STAGES * 2 * OPB);
  unsigned long long* empty = full + STAGES;

  const bool dynamic = g.sched != nullptr;
  if (!dynamic && g.guard != 0 && guard_skips(g, blockIdx.y)) return;   // static grid: gridDim.y == batch, one matrix per CTA
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
  const long nitems = (long)g.batch * g.ntiles;
  // (matrix, block row) of item w
  auto decode = [&](long w, int& mat, int& i) {
    mat = (int)(w / g.ntiles);
    const int t = (int)(w - (long)mat * g.ntiles);
    if (MODE == LL_UPDATE) i = s + t;
    else if (MODE == TK_DIAG) i = s;
    else i = s + 1 + t;
  };

  if (threadIdx.x == 0) {
    for (int st = 0; st < STAGES; ++st) { mbar_init(&full[st], 1); mbar_init(&empty[st], THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- producer (thread 0): chunk p of this CTA's sequence -> ring slot p % STAGES
  int p = 0, p_qq = 0, p_mat = 0, p_i = 0, p_ts = 0, p_static = 0;
  bool p_done = false;
  // static grid: tiles bx, bx + gx, ... of matrices by, by + gy, ...
  const int per = (!dynamic && (int)blockIdx.x < g.ntiles) ? (g.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int nmat_my = (!dynamic && (int)blockIdx.y < g.batch) ? (g.batch - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y : 0;
  auto next_item = [&]() -> long {       // next item of this CTA, or -1
    for (;;) {
      long w;
      if (dynamic) {
        w = (long)atomicAdd(g.sched, 1);
        if (w >= nitems) return -1;
      } else {
        if (p_static >= per * nmat_my) return -1;
        const int mi = p_static / per, ti = p_static - mi * per;
        ++p_static;
        w = (long)((int)blockIdx.y + mi * (int)gridDim.y) * g.ntiles + (int)blockIdx.x + ti * (int)gridDim.x;
      }
      decode(w, p_mat, p_i);
      if (!dynamic || g.guard == 0 || !guard_skips(g, p_mat)) return w;   // the guarded inverse: items of the other path are skipped
    }
  };
  auto issue_next = [&]() {
    if (p_done) return;
    const int slot = p % STAGES;
    if (p >= STAGES) mbar_wait(&empty[slot], ((p / STAGES) - 1) & 1);   // all four warps are done with its previous contents
    if (p_qq == 0) {
      const long w = next_item();
      item_ring[p_ts & 3] = (int)w;
      ++p_ts;
      if (w < 0) {                       // out of work: wake the consumers with an empty phase, retire the CTA
        p_done = true;
        mbar_arrive(&full[slot]);
        if (dynamic) {
          __threadfence();
          const int ncta = (int)(gridDim.x * gridDim.y);
          if (atomicAdd(g.sched + 1, 1) == ncta - 1) {   // every CTA has fetched its last item: reset for the next launch
            g.sched[0] = 0;
            g.sched[1] = 0;
            __threadfence();
          }
        }
        return;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int qq = p_qq, i = p_i, mat = p_mat;
    if (++p_qq == nchunks) p_qq = 0;
    ++p;
    const int kb = qq / CPB, kcol = (qq % CPB) * KC;
    const int dslot = (mat * g.Kt + s) * 2;    // W_ss (slot + 0) and W_ss^T (slot + 1) in the Dinv tensor
    const int pslot = mat * g.Kt;              // P(c)^T tiles of this matrix in the Pbuf tensor
    unsigned char* SA = ring + slot * 2 * OPB;
    unsigned char* SB = SA + OPB;
    unsigned long long* bar = &full[slot];
    mbar_expect_tx(bar, 2 * OPB);
    if (MODE == LL_UPDATE) {          // A(i,kb) . A(s,kb)^T
      tma_load_chunk(SA, &mapA, kb * NB + kcol, i * NB, mat, bar);
      tma_load_chunk(SB, &mapA, kb * NB + kcol, s * NB, mat, bar);
    } else if (MODE == LL_SOLVE) {    // A(i,s) . W^T          (B[n][k] = W[n][k])
      tma_load_chunk(SA, &mapA, s * NB + kcol, i * NB, mat, bar);
      tma_load_chunk(SB, &mapD, kcol, 0, dslot, bar);
    } else if (MODE == TK_PANEL) {    // L(i,s) . W            (B[n][k] = W^T[n][k])
      tma_load_chunk(SA, &mapA, s * NB + kcol, i * NB, mat, bar);
      tma_load_chunk(SB, &mapD, kcol, 0, dslot + 1, bar);
    } else if (MODE == TK_COL) {      // Z(i,cb) . P(cb)       (B[n][k] = P^T(cb)[n][k])
      const int cb = s + 1 + kb;
      tma_load_chunk(SA, &mapA, cb * NB + kcol, i * NB, mat, bar);
      tma_load_chunk(SB, &mapP, kcol, 0, pslot + cb, bar);
    } else {
      if (kb == 0) {                  // W^T W                 (A[m][k] = W^T[m][k], B[n][k] = W^T[n][k])
        tma_load_chunk(SA, &mapD, kcol, 0, dslot + 1, bar);
        tma_load_chunk(SB, &mapD, kcol, 0, dslot + 1, bar);
      } else {                        // P(cb)^T Z(cb,s)       (A[m][k] = P^T(cb)[m][k], B[n][k] = Z(s,cb)[n][k])
        const int cb = s + kb;
        tma_load_chunk(SA, &mapP, kcol, 0, pslot + cb, bar);
        tma_load_chunk(SB, &mapA, cb * NB + kcol, s * NB, mat, bar);
      }
    }
  };
  if (threadIdx.x == 0) {
#pragma unroll 1
    for (int k = 0; k < STAGES - 1; ++k) issue_next();
  }

  // ---- consumers.  The quarter of the tile owned by a warp is rotated per ITEM: the quarters do unequal work in the
  // triangular modes below and every warp is pinned to one SM sub-partition (one DMMA pipe each).
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = lane >> 2, c = 2 * (lane & 3);
  const int lr = lane >> 2, lk = lane & 3;
  // per-lane swizzled byte offsets of the four k-steps: 16-byte slot (s + 4*(lk>>1)) ^ (row & 7), half lk & 1
  int off[4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) off[ks] = (((ks + 4 * (lk >> 1)) ^ lr) << 4) + (lk & 1) * 8;

  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  int q = 0;                      // chunks consumed by this CTA so far
  for (int ts = 0;; ++ts) {       // tiles
    int mat = 0, i = 0, m0 = 0, n0 = 0;
    bool active = false;
    for (int qq = 0; qq < nchunks; ++qq, ++q) {
      if (threadIdx.x == 0) issue_next();          // chunk q + STAGES - 1
      const int slot = q % STAGES;
      mbar_wait(&full[slot], (q / STAGES) & 1);
      if (qq == 0) {                               // first chunk of a tile: which item is it?
        const int w = item_ring[ts & 3];
        if (w < 0) return;                         // out of work (nothing is in flight: the producer stopped here too)
        decode(w, mat, i);
        const int warp = (wid + w) & 3;
        m0 = (warp >> 1) * 32;
        n0 = (warp & 1) * 32;
        const int rows_valid = (i == last) ? vlast : NB;
        // structurally unnecessary quarter: strictly-upper 32 x 32 of a diagonal tile (potrf reads only the lower triangle
        // of A(k,k); Z(j,j) is symmetric and mirrored in the epilogue)
        const bool tri_skip = (m0 == 0 && n0 == NB / 2) && i == j && (MODE == LL_UPDATE || MODE == TK_DIAG);
        active = (m0 < rows_valid) && (n0 < cols_valid) && !tri_skip;
      }
      // k-chunks that only multiply structural zeros of the triangular W = L_kk^-1 (W[a][b] = 0 for b > a):
      //   LL_SOLVE  C[m][n] = sum_k A[m][k] W[n][k]   -> columns n < 32 need k < 32 only
      //   TK_PANEL  C[m][n] = sum_k L[m][k] W[k][n]   -> columns n >= 32 need k >= 32 only
      //   TK_DIAG   (first k-block) sum_k W[k][m] W[k][n] -> any quarter touching rows/cols >= 32 needs k >= 32 only
      constexpr int HALF = (NB / 2) / KC;   // chunks per half block
      bool zero = false;
      if (MODE == LL_SOLVE) zero = n0 == 0 && qq >= HALF;
      else if (MODE == TK_PANEL) zero = n0 == NB / 2 && qq < HALF;
      else if (MODE == TK_DIAG) zero = (m0 == NB / 2 || n0 == NB / 2) && qq < HALF;
      if (active && !zero) {
        const unsigned char* SA = ring + slot * 2 * OPB;
        mma_chunk(SA + (m0 + lr) * 128, SA + OPB + (n0 + lr) * 128, off, acc);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);   // this warp's reads of the slot are complete
    }
    // ---- epilogue of tile (i, j) of matrix `mat`: registers only
    if (active) {
      double* Am = g.A + (long)mat * g.strideA;
      if (MODE == TK_PANEL) {   // P(i)^T = -(L(i,s) W)^T
        double* PT = g.Pbuf + (long)mat * g.strideP + (long)i * NB * NB;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int row = m0 + 8 * a + r, col = n0 + 8 * b + c;
            PT[(long)col * NB + row] = -acc[a][b][0];
            PT[(long)(col + 1) * NB + row] = -acc[a][b][1];
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
            double2* pp = reinterpret_cast<double2*>(C + (long)row * g.ld + col);
            double2 v;
            if (MODE == LL_UPDATE) {
              v = *pp;
              v.x -= acc[a][b][0];
              v.y -= acc[a][b][1];
            } else {
              v.x = acc[a][b][0];
              v.y = acc[a][b][1];
            }
            *pp = v;
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
  }
}

// ------------------------------------------------------------------------------------------------ stable inverse
// Z = W^T W with W = L^-1, in three launches over ALL block columns at once (no sweep, no error recursion: the Takahashi
// sweep above amplifies rounding errors geometrically in the number of block columns and is only used for Kt <= 16):
//   PANEL_ALL  Lh(c,j) = L(c,j) W_jj for every c > j            -> transposed into the upper tile (j,c) of A
//   TRTRI_ROW  row i of W:  W(i,j) = -sum_{c=j+1..i} W(i,c) Lh(c,j),  j = i-1 .. 0  (W(i,i) = W_ii from the diagonal step)
//              one CTA per row, its tiles in sequence; W(i,j) -> lower tile (i,j) of A (operand of the row's later
//              tiles) and transposed -> upper tile (j,i) of A2 (operand of LAUUM)
//   LAUUM      Z(i,j) = sum_{c>=i} W(c,i)^T W(c,j), i >= j      -> A, both triangles
// Same ring, same K-major operands, same fragment addressing as panel_gemm_kernel; the tiles of a CTA now have different
// K lengths, so producer and consumer each walk the (tile, chunk) sequence.
//
// Recursive (level-synchronous) triangular inverse, for a handful of LARGE matrices where one CTA per block row is far too
// little parallelism:  with L = [[L11, 0], [L21, L22]],  W = [[W11, 0], [-W22 (L21 W11), W22]].  Level s = 1, 2, 4, ...
// (half-size in blocks) handles ALL pairs of adjacent s-blocks at once in two launches whose tiles are all independent:
//   REC_T   T(i,j)  = sum_{c=j..e} L(i,c) W(c,j)      i in the pair's second half, j in its first half [.., e]
//                     -> transposed into the upper tile (j,i) of A            (W(c,j)^T: upper tiles of A2, W_jj^T from Dinv)
//   REC_W   W(i,j)  = -sum_{c=b..i} W(i,c) T(c,j)     b = first block of the second half
//                     -> lower tile (i,j) of A (over L21) and transposed -> upper tile (j,i) of A2 (operand of LAUUM)
// 2 log2(Kt) launches instead of a 3 Kt sweep, every tile a long-K accumulation in registers.
enum InvMode : int { PANEL_ALL = 0, TRTRI_ROW = 1, LAUUM = 2, REC_T = 3, REC_W = 4, SYRK_UPD = 5 };

struct Tile {
  int i, j, nch;
  int c0;   // REC_T: first k-block (ascending);  REC_W: first k-block (descending from i)
};

__device__ __forceinline__ void tri_decode(int t, int& a, int& b) {
  // t = a(a+1)/2 + b, 0 <= b <= a
  a = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((long)a * (a + 1) / 2 > t) --a;
  while ((long)(a + 1) * (a + 2) / 2 <= t) ++a;
  b = t - a * (a + 1) / 2;
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 4) inverse_kernel(const __grid_constant__ CUtensorMap mapA,
                                                             const __grid_constant__ CUtensorMap mapA2,
                                                             const __grid_constant__ CUtensorMap mapD, LLArgs g) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment by OFFSET arithmetic on the shared pointer (an integer round trip would make every operand fetch a
  // generic LD instead of LDS: the compiler loses the address space)
  unsigned char* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + STAGES * 2 * OPB);

  const int Kt = g.Kt, last = g.Kt - 1;
  const int vlast = g.n8 - last * NB;
  const int ragged_chunks = (vlast + KC - 1) / KC;
  const int warp = ((threadIdx.x >> 5) + blockIdx.x + blockIdx.y) & 3, lane = threadIdx.x & 31;
  const int m0 = (warp >> 1) * 32, n0 = (warp & 1) * 32;
  const int r = lane >> 2, c = 2 * (lane & 3);
  const int lr = lane >> 2, lk = lane & 3;
  int off[4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) off[ks] = (((ks + 4 * (lk >> 1)) ^ lr) << 4) + (lk & 1) * 8;
  constexpr int HALF = (NB / 2) / KC;

  // tiles of this CTA
  int nmy;
  const int hs = g.step;                                          // REC_*: half-size of the level, in blocks
  if (MODE == TRTRI_ROW) nmy = Kt - 1 - (int)blockIdx.x;        // CTA bx owns row i = Kt-1-bx (longest rows first)
  else nmy = ((int)blockIdx.x < g.ntiles) ? (g.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto tile_at = [&](int tt) -> Tile {
    Tile t;
    if (MODE == PANEL_ALL) {          // strictly lower pairs (c, j)
      int a, b;
      tri_decode(blockIdx.x + tt * gridDim.x, a, b);
      t.i = a + 1; t.j = b; t.nch = CPB;
    } else if (MODE == TRTRI_ROW) {   // row i, columns descending
      t.i = Kt - 1 - (int)blockIdx.x; t.j = t.i - 1 - tt; t.nch = (t.i - t.j) * CPB;
    } else if (MODE == REC_T || MODE == REC_W) {
      // full pairs hold hs x hs tiles; only the last pair can have a shorter (or no) second half
      const int tl = blockIdx.x + tt * gridDim.x;
      const int per = hs * hs;
      const int full = (Kt / (2 * hs)) * per;   // tiles of the complete pairs
      int pidx, a, b;
      if (tl < full) { pidx = tl / per; const int rem = tl - pidx * per; a = rem / hs; b = rem - a * hs; }
      else { pidx = Kt / (2 * hs); const int rem = tl - full; a = rem / hs; b = rem - a * hs; }
      const int base = pidx * 2 * hs;
      t.i = base + hs + a; t.j = base + b;
      if (MODE == REC_T) { t.c0 = t.j; t.nch = (hs - b) * CPB; }          // c = j .. base+hs-1
      else { t.c0 = t.i; t.nch = (a + 1) * CPB; }                          // c = i .. base+hs (descending)
    } else if (MODE == SYRK_UPD) {    // lower pairs incl. diagonal of the trailing triangle [ua, Kt); k-blocks ukb0 .. ukb0 + unkb - 1
      if (g.ujb == 0) {
        int a, b;
        tri_decode(blockIdx.x + tt * gridDim.x, a, b);
        t.i = g.ua + a; t.j = g.ua + b;
      } else {                        // a few block columns: column by column, Kt - j tiles each
        int tl = blockIdx.x + tt * gridDim.x, jj = g.ua;
        while (jj < g.ujb - 1 && tl >= Kt - jj) { tl -= Kt - jj; ++jj; }
        t.i = jj + tl; t.j = jj;
      }
      t.c0 = g.ukb0; t.nch = g.unkb * CPB;
    } else {                          // lower pairs incl. diagonal; k-blocks c = i .. last, the last one ragged
      tri_decode(blockIdx.x + tt * gridDim.x, t.i, t.j);
      t.nch = (last - t.i) * CPB + ragged_chunks;
    }
    return t;
  };

  if (threadIdx.x == 0) {
    for (int st = 0; st < STAGES; ++st) mbar_init(&full[st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (MODE == SYRK_UPD) {   // the chain launches of the look-ahead potrf may be programmatic dependents (common.cuh)
    pdl_launch_dependents();
    pdl_wait();
  }
  __syncthreads();
  int gq = 0;

  for (int mat = blockIdx.y; mat < g.batch; mat += gridDim.y) {
    if (guard_skips(g, mat)) continue;   // CTA-uniform; nothing of this matrix has touched the ring yet
    double* Am = g.A + (long)mat * g.strideA;
    double* A2m = g.A2 + (long)mat * g.strideA;
    const int dbase = mat * Kt * 2;

    // ---- producer (thread 0): walks the same (tile, chunk) sequence STAGES-1 chunks ahead of the consumers
    int p_tt = 0, p_qq = 0;
    Tile p_tile = tile_at(0);
    auto produce = [&](int slot) {
      const Tile t = p_tile;
      const int kb = p_qq / CPB, kcol = (p_qq % CPB) * KC;
      if (++p_qq == t.nch) { p_qq = 0; if (++p_tt < nmy) p_tile = tile_at(p_tt); }
      unsigned char* SA = ring + slot * 2 * OPB;
      unsigned char* SB = SA + OPB;
      unsigned long long* bar = &full[slot];
      mbar_expect_tx(bar, 2 * OPB);
      if (MODE == PANEL_ALL) {           // L(i,j) . W_jj              (B[n][k] = W_jj^T[n][k])
        tma_load_chunk(SA, &mapA, t.j * NB + kcol, t.i * NB, mat, bar);
        tma_load_chunk(SB, &mapD, kcol, 0, dbase + 2 * t.j + 1, bar);
      } else if (MODE == TRTRI_ROW) {    // W(i,cb) . Lh(cb,j),  cb = i, i-1, .., j+1   (B = upper tile (j,cb) of A)
        const int cb = t.i - kb;
        if (cb == t.i) tma_load_chunk(SA, &mapD, kcol, 0, dbase + 2 * t.i, bar);
        else tma_load_chunk(SA, &mapA, cb * NB + kcol, t.i * NB, mat, bar);
        tma_load_chunk(SB, &mapA, cb * NB + kcol, t.j * NB, mat, bar);
      } else if (MODE == REC_T) {        // L(i,cb) . W(cb,j),  cb = j, j+1, ..        (B = W(cb,j)^T: upper tile (j,cb) of A2)
        const int cb = t.c0 + kb;
        tma_load_chunk(SA, &mapA, cb * NB + kcol, t.i * NB, mat, bar);
        if (cb == t.j) tma_load_chunk(SB, &mapD, kcol, 0, dbase + 2 * t.j + 1, bar);
        else tma_load_chunk(SB, &mapA2, cb * NB + kcol, t.j * NB, mat, bar);
      } else if (MODE == REC_W) {        // W(i,cb) . T(cb,j),  cb = i, i-1, ..        (B = T(cb,j)^T: upper tile (j,cb) of A)
        const int cb = t.c0 - kb;
        if (cb == t.i) tma_load_chunk(SA, &mapD, kcol, 0, dbase + 2 * t.i, bar);
        else tma_load_chunk(SA, &mapA, cb * NB + kcol, t.i * NB, mat, bar);
        tma_load_chunk(SB, &mapA, cb * NB + kcol, t.j * NB, mat, bar);
      } else if (MODE == SYRK_UPD) {     // A(i,cb) . A(j,cb)^T,  cb = ukb0 ..         (panel tiles of the factor)
        const int cb = t.c0 + kb;
        tma_load_chunk(SA, &mapA, cb * NB + kcol, t.i * NB, mat, bar);
        tma_load_chunk(SB, &mapA, cb * NB + kcol, t.j * NB, mat, bar);
      } else {                           // W(cb,i)^T . W(cb,j),  cb = i .. last       (upper tiles (i,cb), (j,cb) of A2)
        const int cb = t.i + kb;
        if (cb == t.i) tma_load_chunk(SA, &mapD, kcol, 0, dbase + 2 * t.i + 1, bar);
        else tma_load_chunk(SA, &mapA2, cb * NB + kcol, t.i * NB, mat, bar);
        if (cb == t.j) tma_load_chunk(SB, &mapD, kcol, 0, dbase + 2 * t.j + 1, bar);
        else tma_load_chunk(SB, &mapA2, cb * NB + kcol, t.j * NB, mat, bar);
      }
    };

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    if (threadIdx.x == 0 && nmy > 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll 1
      for (int q = 0; q < STAGES - 1 && p_tt < nmy; ++q) produce((gq + q) % STAGES);
    }

    for (int tt = 0; tt < nmy; ++tt) {
      const Tile t = tile_at(tt);
      const int rows_valid = (t.i == last) ? vlast : NB;
      const int cols_valid = (t.j == last) ? vlast : NB;
      const bool diag_tile = (MODE == LAUUM || MODE == SYRK_UPD) && t.i == t.j;   // only the lower triangle is needed
      const bool active = (m0 < rows_valid) && (n0 < cols_valid) && !(diag_tile && m0 == 0 && n0 == NB / 2);
      for (int qq = 0; qq < t.nch; ++qq, ++gq) {
        const int slot = gq % STAGES;
        __syncthreads();                 // every warp is done with the previous chunk: its slot may be refilled
        if (threadIdx.x == 0 && p_tt < nmy) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          produce((gq + STAGES - 1) % STAGES);
        }
        // chunks of the first k-block that only meet structural zeros of the triangular diagonal inverse
        bool zero = false;
        if (qq < CPB && MODE != SYRK_UPD) {
          if (MODE == PANEL_ALL || MODE == REC_T) zero = (n0 == NB / 2 && qq < HALF);   // W_jj[k][n] = 0 for n > k
          else if (MODE == TRTRI_ROW || MODE == REC_W) zero = (m0 == 0 && qq >= HALF);  // W_ii[m][k] = 0 for k > m
          else zero = (m0 == NB / 2 && qq < HALF) || (diag_tile && n0 == NB / 2 && qq < HALF);   // W_ii[k][m] = 0 for m > k
        }
        mbar_wait(&full[slot], (gq / STAGES) & 1);
        if (active && !zero) {
          const unsigned char* SA = ring + slot * 2 * OPB;
          mma_chunk(SA + (m0 + lr) * 128, SA + OPB + (n0 + lr) * 128, off, acc);
        }
      }
      // ---- epilogue
      if (active) {
        double* C = Am + ((long)t.i * NB) * g.ld + (long)t.j * NB;        // lower tile (i,j) of A
        double* U = Am + ((long)t.j * NB) * g.ld + (long)t.i * NB;        // upper tile (j,i) of A
        double* U2 = A2m + ((long)t.j * NB) * g.ld + (long)t.i * NB;      // upper tile (j,i) of A2
        const bool mirror = !diag_tile || (m0 == NB / 2 && n0 == 0);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int row = m0 + 8 * a + r, col = n0 + 8 * b + c;
            if (MODE == PANEL_ALL || MODE == REC_T) {   // Lh(i,j)^T / T(i,j)^T -> upper tile (j,i) of A
              U[(long)col * g.ld + row] = acc[a][b][0];
              U[(long)(col + 1) * g.ld + row] = acc[a][b][1];
            } else if (MODE == SYRK_UPD) {   // A(i,j) -= sum
              double2* pp = reinterpret_cast<double2*>(C + (long)row * g.ld + col);
              double2 v = *pp;
              v.x -= acc[a][b][0];
              v.y -= acc[a][b][1];
              *pp = v;
            } else if (MODE == TRTRI_ROW || MODE == REC_W) {   // W(i,j) -> lower tile of A, W(i,j)^T -> upper tile of A2
              double2 v; v.x = -acc[a][b][0]; v.y = -acc[a][b][1];
              *reinterpret_cast<double2*>(C + (long)row * g.ld + col) = v;
              U2[(long)col * g.ld + row] = v.x;
              U2[(long)(col + 1) * g.ld + row] = v.y;
            } else {                          // Z(i,j) and its mirror -> A
              double2 v; v.x = acc[a][b][0]; v.y = acc[a][b][1];
              *reinterpret_cast<double2*>(C + (long)row * g.ld + col) = v;
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
      // the next tiles of this row read W(i,j) back through TMA: order this thread's global writes before later
      // async-proxy reads (the barrier at the top of the next chunk orders the CTA's threads among themselves)
      if (MODE == TRTRI_ROW) asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();   // all warps have consumed the last chunk before the next matrix's prologue refills the ring
  }
}

// ------------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// {k (contiguous), row, tile-or-matrix} view of `count` row-major [rows][cols] matrices `stride` doubles apart
int make_map(CUtensorMap* m, const double* base, long cols, long rows, long count, long stride) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) { set_last_error("cuTensorMapEncodeTiled is not available from this driver"); return -2; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)count};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 8, (cuuint64_t)stride * 8};
  cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)NB, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  // L2 promotion: a box row is ONE 128-byte line of a matrix row and the next k-chunk of the ring is the neighbouring line of
  // the same rows, so promoting the fetch to 256 bytes could bring it into L2 with the same DRAM burst -- measured: no
  // difference (110.05 / 110.28 / 110.04 ms per C4 sweep for 128 / 256 / none, profiles/r02_engine_ab.txt).  NMGP_TMA_L2: A/B.
  static const int l2 = getenv("NMGP_TMA_L2") ? atoi(getenv("NMGP_TMA_L2")) : 128;
  const CUtensorMapL2promotion promo = l2 == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                       : (l2 == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                    : (l2 == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return -2;
  }
  return 0;
}

struct MapSet {
  CUtensorMap mA, mA2, mD, mP;
};

// Tensor maps are pure functions of (pointers, layout).  A plan builds them once for its workspace (engine_maps_create,
// owned by the plan and released with it); a BlockBatch without them (unit entry points) gets a set built per call.  The
// kernels take the maps BY VALUE (__grid_constant__), so a set may be released as soon as the launches are queued.
int build_maps(const BlockBatch& b, MapSet* ms) {
  NMGP_TRY(make_map(&ms->mA, b.A, b.nP, b.nP, b.batch, b.strideA()));
  NMGP_TRY(make_map(&ms->mA2, b.A2 ? b.A2 : b.A, b.nP, b.nP, b.batch, b.strideA()));
  NMGP_TRY(make_map(&ms->mD, b.Dinv, NB, NB, (long)b.batch * b.Kt * 2, (long)NB * NB));
  NMGP_TRY(make_map(&ms->mP, b.Pbuf, NB, NB, (long)b.batch * b.Kt, (long)NB * NB));
  return 0;
}

struct MapRef {
  MapSet local;
  const MapSet* ms = nullptr;
  int init(const BlockBatch& b) {
    if (b.maps) { ms = static_cast<const MapSet*>(b.maps); return 0; }
    NMGP_TRY(build_maps(b, &local));
    ms = &local;
    return 0;
  }
};

// co-resident CTAs of a kernel on the current device (SMs x occupancy), cached per device and call site
template <typename K>
int resident_ctas(K kernel, int threads, size_t smem, std::atomic<int>* cache) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  const bool slot = dev >= 0 && dev < kMaxDevices;
  if (slot) { const int c = cache[dev].load(std::memory_order_relaxed); if (c > 0) return c; }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int v = per_sm * sm_count();
  if (slot) cache[dev].store(v, std::memory_order_relaxed);
  return v;
}

// dynamic scheduling (scheduler word present): one persistent CTA per resident slot; otherwise the static 2-D grid
template <int MODE>
int launch_ll(const MapSet& ms, const LLArgs& g0, int ntiles, cudaStream_t st, long* launches) {
  if (ntiles <= 0 || g0.batch <= 0) return 0;
  NMGP_SMEM_ATTR_PER_DEVICE(panel_gemm_kernel<MODE>, LL_SMEM);
  static std::atomic<int> slots[kMaxDevices];
  LLArgs g = g0;
  g.ntiles = ntiles;
  const long resident = resident_ctas(panel_gemm_kernel<MODE>, PG_THREADS, LL_SMEM, slots);
  if (g.sched) {
    const long nitems = (long)g.batch * ntiles;
    const long grid = resident < nitems ? resident : nitems;
    panel_gemm_kernel<MODE><<<(unsigned)grid, PG_THREADS, LL_SMEM, st>>>(ms.mA, ms.mD, ms.mP, g);
  } else {
    // tiles of one matrix are split over gx CTAs; ~`waves` waves of the resident slots (hardware-dynamic CTA assignment)
    static const int waves = getenv("NMGP_LL_WAVES") ? atoi(getenv("NMGP_LL_WAVES")) : 32;   // A/B timing: 8 -> 32: -1.5 % time
    const int gy = g.batch < 65535 ? g.batch : 65535;
    if (g.guard != 0 && gy != g.batch) { set_last_error("guarded inverse: more than 65535 matrices in one batch"); return -1; }
    int gx = (int)((resident * waves + gy - 1) / gy);
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    dim3 grid(gx, gy);
    panel_gemm_kernel<MODE><<<grid, PG_THREADS, LL_SMEM, st>>>(ms.mA, ms.mD, ms.mP, g);
  }
  NMGP_CUDA_TRY(cudaGetLastError());
  if (launches) ++*launches;
  return 0;
}

LLArgs make_ll(const BlockBatch& b) {
  LLArgs g;
  g.A = b.A; g.A2 = b.A2; g.Pbuf = b.Pbuf;
  g.strideA = b.strideA(); g.strideP = b.strideP();
  g.ld = b.nP; g.Kt = b.Kt; g.batch = b.batch; g.step = 0;
  g.n8 = (int)round_up(b.n, 8);
  g.ntiles = 0;
  g.guard = 0; g.guard_thr = 0.0; g.pivmin = b.pivmin; g.pivmax = b.pivmax;
  static const bool no_dyn = getenv("NMGP_LL_STATIC") != nullptr;   // A/B timing: the static 2-D grid
  g.sched = no_dyn ? nullptr : b.sched;
  return g;
}

}  // namespace

int engine_potrf_ll(const BlockBatch& b, cudaStream_t st, long* launches) {
  if (b.batch <= 0) return 0;
  if (!b.Pbuf) { set_last_error("engine_potrf_ll: no panel buffer"); return -1; }
  MapRef mr;
  NMGP_TRY(mr.init(b));
  const MapSet* ms = mr.ms;
  LLArgs g = make_ll(b);
  for (int k = 0; k < b.Kt; ++k) {
    g.step = k;
    if (k > 0) NMGP_TRY(launch_ll<LL_UPDATE>(*ms, g, b.Kt - k, st, launches));
    NMGP_TRY(engine_diag_step(b, k, st, launches));
    NMGP_TRY(launch_ll<LL_SOLVE>(*ms, g, b.Kt - k - 1, st, launches));
  }
  return 0;
}

namespace {
template <int MODE>
int launch_inv(const MapSet& ms, const LLArgs& g0, cudaStream_t st, long* launches, int hs = 0, bool pdl = false) {
  if (g0.batch <= 0) return 0;
  LLArgs g = g0;
  const int Kt = g.Kt;
  int ntiles;
  if (MODE == REC_T || MODE == REC_W) {   // level of half-size hs: complete pairs + the (shorter) second half of the last one
    const int full = Kt / (2 * hs), rest = Kt - full * 2 * hs;
    ntiles = full * hs * hs + (rest > hs ? (rest - hs) * hs : 0);
    g.step = hs;
  } else if (MODE == SYRK_UPD) {
    if (g.ujb == 0) {
      const int T = Kt - g.ua;
      ntiles = T * (T + 1) / 2;
    } else {
      ntiles = 0;
      for (int jj = g.ua; jj < g.ujb; ++jj) ntiles += Kt - jj;
    }
  } else {
    ntiles = MODE == PANEL_ALL ? Kt * (Kt - 1) / 2 : (MODE == TRTRI_ROW ? Kt - 1 : Kt * (Kt + 1) / 2);
  }
  if (ntiles <= 0) return 0;
  NMGP_SMEM_ATTR_PER_DEVICE(inverse_kernel<MODE>, LL_SMEM);
  g.ntiles = ntiles;
  const int gy = g.batch < 65535 ? g.batch : 65535;
  int gx;
  if (MODE == TRTRI_ROW) gx = Kt - 1;          // one CTA per row (its tiles depend on each other)
  else {
    gx = (sm_count() * 4 * 8 + gy - 1) / gy;   // ~8 waves of the resident slots (4 CTAs per SM)
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
  }
  dim3 grid(gx, gy);
  NMGP_CUDA_TRY(launch_kernel_pdl(inverse_kernel<MODE>, grid, dim3(THREADS), LL_SMEM, st, pdl, ms.mA, ms.mA2, ms.mD, g));
  NMGP_CUDA_TRY(cudaGetLastError());
  if (launches) ++*launches;
  return 0;
}
}  // namespace

// A <- inverse (both triangles) from the factor and the diagonal inverses left by engine_potrf_ll; stable for any size.
int engine_potri_ll_stable(const BlockBatch& b, cudaStream_t st, long* launches) {
  if (b.batch <= 0) return 0;
  if (!b.A2) { set_last_error("engine_potri_ll_stable: no second matrix buffer"); return -1; }
  MapRef mr;
  NMGP_TRY(mr.init(b));
  const MapSet* ms = mr.ms;
  LLArgs g = make_ll(b);
  NMGP_TRY(launch_inv<PANEL_ALL>(*ms, g, st, launches));
  NMGP_TRY(launch_inv<TRTRI_ROW>(*ms, g, st, launches));
  NMGP_TRY(launch_inv<LAUUM>(*ms, g, st, launches));
  return 0;
}

// Same result with the triangular inverse formed level by level (REC_T / REC_W, 2 log2(Kt) launches of independent
// long-K tiles) instead of row by row: the path for a few large matrices (n = 5000 ... 16 384 with batch < 8), where one
// CTA per block row leaves most of the GPU idle.  Works on the factor of either potrf engine (needs L in the lower tiles of
// A and W_kk / W_kk^T in Dinv); uses the upper tiles of A as scratch and A2 for the transposed inverse.
int engine_potri_ll_recursive(const BlockBatch& b, cudaStream_t st, long* launches) {
  if (b.batch <= 0) return 0;
  if (!b.A2 || !b.Pbuf) { set_last_error("engine_potri_ll_recursive: no second matrix buffer"); return -1; }
  MapRef mr;
  NMGP_TRY(mr.init(b));
  const MapSet* ms = mr.ms;
  LLArgs g = make_ll(b);
  for (int hs = 1; hs < b.Kt; hs *= 2) {
    NMGP_TRY(launch_inv<REC_T>(*ms, g, st, launches, hs));
    NMGP_TRY(launch_inv<REC_W>(*ms, g, st, launches, hs));
  }
  NMGP_TRY(launch_inv<LAUUM>(*ms, g, st, launches));
  return 0;
}

// threshold on (min pivot / max pivot)^2 below which a matrix leaves the Takahashi sweep: calibrated on the grid of
// profiles/r02_takahashi_stress.txt with a margin of ~4x in the noise variance (sweep error <= 2e-10 at the threshold)
static double takahashi_guard_threshold(int Kt) {
  static const char* ev = getenv("NMGP_TAKAHASHI_GUARD");   // A/B timing: 0 = no guard (raw sweep), x = threshold
  if (ev) return atof(ev);
  return Kt <= 10 ? 5e-4 : 2.5e-3;
}

int engine_potri_ll_guarded(const BlockBatch& b, cudaStream_t st, long* launches) {
  if (b.batch <= 0) return 0;
  const double thr = takahashi_guard_threshold(b.Kt);
  if (!(thr > 0.0) || !b.pivmin || !b.pivmax) return engine_potri_ll(b, st, launches);
  if (!b.Pbuf || !b.A2) { set_last_error("engine_potri_ll_guarded: no panel / second matrix buffer"); return -1; }
  MapRef mr;
  NMGP_TRY(mr.init(b));
  const MapSet* ms = mr.ms;
  LLArgs g = make_ll(b);
  g.guard_thr = thr;
  g.guard = 1;                                   // Takahashi sweep: the well-conditioned matrices
  for (int j = b.Kt - 1; j >= 0; --j) {
    g.step = j;
    NMGP_TRY(launch_ll<TK_PANEL>(*ms, g, b.Kt - 1 - j, st, launches));
    NMGP_TRY(launch_ll<TK_COL>(*ms, g, b.Kt - 1 - j, st, launches));
    NMGP_TRY(launch_ll<TK_DIAG>(*ms, g, 1, st, launches));
  }
  g.guard = 2;                                   // W^T W: the rest (normally none: three launches of CTAs that return at once)
  g.step = 0;
  NMGP_TRY(launch_inv<PANEL_ALL>(*ms, g, st, launches));
  NMGP_TRY(launch_inv<TRTRI_ROW>(*ms, g, st, launches));
  NMGP_TRY(launch_inv<LAUUM>(*ms, g, st, launches));
  return 0;
}

// Trailing update of the right-looking potrf on the TMA-ring kernel:  A(i,j) -= sum_{kb0 <= c < kb0+nkb} A(i,c) A(j,c)^T for
// the lower tiles i >= j, ja <= j < jb (jb = 0 or Kt: the whole trailing triangle; the diagonal tiles' strictly-upper quarter
// is not touched).  Replaces engine.cu's
// syrk_wide_kernel<true> (single-buffered loads, 27 TFLOP/s at n = 16 384) wherever the batch has TMA descriptors or the
// buffers to build them.
int engine_syrk_update_ll(const BlockBatch& b, int ja, int jb, int kb0, int nkb, cudaStream_t st, long* launches, bool pdl) {
  if (b.batch <= 0 || ja >= b.Kt || nkb <= 0 || (jb != 0 && jb <= ja)) return 0;
  MapRef mr;
  NMGP_TRY(mr.init(b));
  LLArgs g = make_ll(b);
  g.ua = ja; g.ujb = jb >= b.Kt ? 0 : jb; g.ukb0 = kb0; g.unkb = nkb;
  return launch_inv<SYRK_UPD>(*mr.ms, g, st, launches, 0, pdl);
}

int engine_maps_create(const BlockBatch& b, void** out) {
  *out = nullptr;
  if (!b.A || !b.Dinv || !b.Pbuf || b.batch <= 0) return 0;   // nothing the left-looking engine could run on
  MapSet* ms = new (std::nothrow) MapSet();
  if (!ms) { set_last_error("engine_maps_create: out of host memory"); return -1; }
  const int rc = build_maps(b, ms);
  if (rc != 0) { delete ms; return rc; }
  *out = ms;
  return 0;
}

void engine_maps_destroy(void* maps) { delete static_cast<MapSet*>(maps); }

int engine_potri_ll(const BlockBatch& b, cudaStream_t st, long* launches) {
  if (b.batch <= 0) return 0;
  if (!b.Pbuf) { set_last_error("engine_potri_ll: no panel buffer"); return -1; }
  MapRef mr;
  NMGP_TRY(mr.init(b));
  const MapSet* ms = mr.ms;
  LLArgs g = make_ll(b);
  for (int j = b.Kt - 1; j >= 0; --j) {
    g.step = j;
    NMGP_TRY(launch_ll<TK_PANEL>(*ms, g, b.Kt - 1 - j, st, launches));
    NMGP_TRY(launch_ll<TK_COL>(*ms, g, b.Kt - 1 - j, st, launches));
    NMGP_TRY(launch_ll<TK_DIAG>(*ms, g, 1, st, launches));
  }
  return 0;
}

}  // namespace nmgp
