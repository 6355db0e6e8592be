// One-warp cyclic Jacobi eigen-decomposition of a small symmetric matrix (shared by models.cu and kron.cu).
#pragma once
#include "common.cuh"

namespace nmgp {

// Cyclic Jacobi eigen-decomposition of a symmetric M x M matrix held in shared memory, by one warp.
// On exit Bm holds the eigenvalues on its diagonal and V the eigenvectors in its columns.
__device__ inline void jacobi_eig_warp(double* Bm, double* V, int M, int lds) {
  const int lane = threadIdx.x & 31;
  for (int idx = lane; idx < M * M; idx += 32) V[(idx / M) * lds + idx % M] = (idx / M == idx % M) ? 1.0 : 0.0;
  __syncwarp();
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int idx = lane; idx < M * M; idx += 32) {
      const int r = idx / M, cc = idx % M;
      const double v = Bm[r * lds + cc];
      if (r != cc) off += v * v; else dg += v * v;
    }
    off = warp_sum(off);
    dg = warp_sum(dg);
    // Cyclic Jacobi converges quadratically, and rounding keeps sum(off^2) near 1e-31 sum(diag^2) forever: once it is below
    // 1e-22 one more sweep takes it to that floor, so that sweep is the last.  (The old test, 1e-40, could never be met and
    // cost all 30 sweeps: 230 us per call, the largest single item of a single-subject separable evaluation.)
    if (off <= 1e-33 * dg || off == 0.0) break;
    const bool final_sweep = off <= 1e-22 * dg;
    for (int p = 0; p < M - 1; ++p)
      for (int q = p + 1; q < M; ++q) {
        const double apq = Bm[p * lds + q];
        if (apq != 0.0) {  // uniform across the warp
          const double app = Bm[p * lds + p], aqq = Bm[q * lds + q];
          const double theta = (aqq - app) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
          __syncwarp();
          // columns p,q of B and V
          for (int r = lane; r < M; r += 32) {
            const double bp = Bm[r * lds + p], bq = Bm[r * lds + q];
            Bm[r * lds + p] = cs * bp - sn * bq;
            Bm[r * lds + q] = sn * bp + cs * bq;
            const double vp = V[r * lds + p], vq = V[r * lds + q];
            V[r * lds + p] = cs * vp - sn * vq;
            V[r * lds + q] = sn * vp + cs * vq;
          }
          __syncwarp();
          // rows p,q of B
          for (int r = lane; r < M; r += 32) {
            const double bp = Bm[p * lds + r], bq = Bm[q * lds + r];
            Bm[p * lds + r] = cs * bp - sn * bq;
            Bm[q * lds + r] = sn * bp + cs * bq;
          }
          __syncwarp();
        }
      }
    if (final_sweep) break;
  }
  __syncwarp();
}

}  // namespace nmgp
