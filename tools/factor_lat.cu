// Cycles of factor_tile (8 x 8 Cholesky + inverse by one warp, csrc/diag.cu) alone on an SM: the link of the 64-pivot chain.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/factor_lat tools/factor_lat.cu
#include "../nonstationary_multivariate_gaussian_process_b200/csrc/diag.cu"
namespace nmgp { void set_last_error(const std::string& m) { fprintf(stderr, "%s\n", m.c_str()); } }
using namespace nmgp;
__global__ void klat(long long* cyc, double* out, int mode) {
  extern __shared__ __align__(16) double sm[];
  double* M = sm; double* V = M + NB * MLD; double* rinv = V + 8 * 64;
  const int lane = threadIdx.x & 31;
  long long total = 0;
  int fsum = 0;
  for (int rep = 0; rep < 16; ++rep) {
    for (int i = lane; i < NB * MLD; i += 32) { const int r = i / MLD, c = i % MLD; M[i] = (r == c ? 3.0 : 0.0) + exp(-0.02 * (r - c) * (r - c)); }
    __syncwarp();
    const long long t0 = clock64();
    if (mode == 0) {
#pragma unroll 1
      for (int p = 0; p < 8; ++p) { fsum += factor_tile(M + (PB * p) * MLD + PB * p, V + p * 64, rinv + PB * p, lane, PB * p); __syncwarp(); }
    } else {
      // the look-ahead update + factor as the kernel chains them: tile p+1 -= L_{p+1,p} L_{p+1,p}^T (here: some tile), then factor
      const int r = lane >> 2, q = lane & 3;
#pragma unroll 1
      for (int p = 0; p < 8; ++p) {
        double* T = M + (PB * p) * MLD + PB * p;
        const double* Lr = M + (PB * p + r) * MLD + ((PB * p + 8) & 63);
        double2 c = *reinterpret_cast<const double2*>(T + r * MLD + 2 * q);
        const double a0 = 1e-3 * Lr[q], a1 = 1e-3 * Lr[4 + q];
        dmma884(c.x, c.y, -a0, a0);
        dmma884(c.x, c.y, -a1, a1);
        *reinterpret_cast<double2*>(T + r * MLD + 2 * q) = c;
        __syncwarp();
        fsum += factor_tile(T, V + p * 64, rinv + PB * p, lane, PB * p);
        __syncwarp();
      }
    }
    total += clock64() - t0;
  }
  if (lane == 0) { cyc[mode] = total / (16 * 8); out[mode] = rinv[5] + V[70] + fsum; }
}
int main() {
  long long* cyc; double* out;
  cudaMalloc(&cyc, 64); cudaMalloc(&out, 64);
  cudaFuncSetAttribute(klat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MMA_SMEM_BYTES);
  for (int mode = 0; mode < 2; ++mode) { klat<<<1, 32, MMA_SMEM_BYTES>>>(cyc, out, mode); cudaDeviceSynchronize(); }
  long long c[2]; double o[2];
  cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost); cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost);
  printf("factor_tile alone:                 %lld cycles per 8 x 8 tile   (check %.6f)\n", c[0], o[0]);
  printf("look-ahead update + factor_tile:   %lld cycles per tile         (check %.6f)\n", c[1], o[1]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
