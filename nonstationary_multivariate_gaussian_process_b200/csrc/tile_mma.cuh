// 64 x 64 tile products on DMMA.8x8x4 from padded shared-memory tiles (shared by engine.cu and diag.cu).
#pragma once
#include "engine.cuh"

namespace nmgp {
namespace tile {

constexpr int NB = kNB;          // 64
constexpr int LDS = NB + 4;      // shared tile stride (doubles): == 4 mod 16 -> conflict-free DMMA fragment loads
constexpr int TILE_THREADS = 128;

// straight copy of one NB x NB tile (row stride ld) into shared [NB][LDS]
__device__ __forceinline__ void load_tile(double* __restrict__ S, const double* __restrict__ src, int ld) {
  constexpr int V2 = NB / 2;  // double2 per row
#pragma unroll 4
  for (int idx = threadIdx.x; idx < NB * V2; idx += TILE_THREADS) {
    const int r = idx / V2, c2 = idx % V2;
    const double2 v = *reinterpret_cast<const double2*>(src + (long)r * ld + 2 * c2);
    *reinterpret_cast<double2*>(S + r * LDS + 2 * c2) = v;
  }
}

// Two tiles with ALL 32 loads of a thread in flight before the first shared-memory store: ONE global round trip instead of
// the eight of two load_tile calls (4 loads in flight each).  For the small launches on the critical chain of a large
// factorisation, where the round trips, not the bytes, are the time; costs 64 transient registers.
__device__ __forceinline__ void load_tiles2(double* __restrict__ SA, const double* __restrict__ srcA, int lda,
                                            double* __restrict__ SB, const double* __restrict__ srcB, int ldb) {
  constexpr int V2 = NB / 2, PER = NB * V2 / TILE_THREADS;   // 16 double2 per thread and tile
  double2 va[PER], vb[PER];
#pragma unroll
  for (int it = 0; it < PER; ++it) {
    const int idx = threadIdx.x + it * TILE_THREADS, r = idx / V2, c2 = idx % V2;
    va[it] = *reinterpret_cast<const double2*>(srcA + (long)r * lda + 2 * c2);
    vb[it] = *reinterpret_cast<const double2*>(srcB + (long)r * ldb + 2 * c2);
  }
#pragma unroll
  for (int it = 0; it < PER; ++it) {
    const int idx = threadIdx.x + it * TILE_THREADS, r = idx / V2, c2 = idx % V2;
    *reinterpret_cast<double2*>(SA + r * LDS + 2 * c2) = va[it];
    *reinterpret_cast<double2*>(SB + r * LDS + 2 * c2) = vb[it];
  }
}

// acc(32x32 per warp) += opA(32 x NB) * opB(32 x NB)^T.   KM: S[row][k],  MM: S[k][row].
template <bool A_KM, bool B_KM>
__device__ __forceinline__ void warp_mma(const double* __restrict__ SA, const double* __restrict__ SB, int m0, int n0,
                                         double (&acc)[4][4][2]) {
  const int lane = threadIdx.x & 31;
  const int lr = lane >> 2, lk = lane & 3;
  const double* pa = A_KM ? SA + (m0 + lr) * LDS + lk : SA + lk * LDS + m0 + lr;
  const double* pb = B_KM ? SB + (n0 + lr) * LDS + lk : SB + lk * LDS + n0 + lr;
  constexpr int a_sub = A_KM ? 8 * LDS : 8;  // next 8-row subtile
  constexpr int b_sub = B_KM ? 8 * LDS : 8;
  constexpr int a_k = A_KM ? 4 : 4 * LDS;    // next k-step of 4
  constexpr int b_k = B_KM ? 4 : 4 * LDS;
#pragma unroll 4
  for (int k = 0; k < NB; k += 4) {
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

}  // namespace tile
}  // namespace nmgp
