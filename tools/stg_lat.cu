// Issue cost (cycles per warp instruction) of global stores from ONE SM, by access pattern: what does it cost the diagonal-block
// kernel to write W (8 x 8 tiles: 8 rows x 64 B per STG.128) and W^T (4 rows x 64 B per STG.64) against full 512-byte rows?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/stg_lat tools/stg_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 128;
__global__ void k(double* buf, long long* cyc, int pattern) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, q = lane & 3;
  double* base = buf + (size_t)warp * N * 4096;     // private region per warp
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) {
    double* p = base + (size_t)i * 4096;            // a fresh 32 KB-aligned tile region per instruction
    if (pattern == 0) *reinterpret_cast<double2*>(p + 2 * lane) = make_double2(1.0, 2.0);                    // 512 B contiguous
    else if (pattern == 1) *reinterpret_cast<double2*>(p + r * 64 + 2 * q) = make_double2(1.0, 2.0);        // 8 rows x 64 B
    else if (pattern == 2) p[(2 * q) * 64 + r] = 1.0;                                                         // 4 rows x 64 B
    else if (pattern == 3) p[lane] = 1.0;                                                                     // 256 B contiguous
    else if (pattern == 4) *reinterpret_cast<double2*>(p + (lane >> 3) * 64 + 2 * (lane & 7)) = make_double2(1.0, 2.0);  // 4 rows x 128 B
  }
  const long long t1 = clock64();
  __syncthreads();
  const long long t2 = clock64();
  if (lane == 0) { cyc[2 * warp] = t1 - t0; cyc[2 * warp + 1] = t2 - t0; }
}
int main() {
  double* buf; long long* cyc;
  cudaMalloc(&buf, (size_t)8 * N * 4096 * 8); cudaMalloc(&cyc, 64 * 8);
  const char* names[] = {"STG.128 512 B contiguous (4 full lines)", "STG.128 8 rows x 64 B (W tile)", "STG.64 4 rows x 64 B (W^T tile)",
                         "STG.64 256 B contiguous", "STG.128 4 rows x 128 B"};
  for (int warps : {1, 4, 8}) {
    for (int pat = 0; pat < 5; ++pat) {
      for (int rep = 0; rep < 2; ++rep) k<<<1, 32 * warps>>>(buf, cyc, pat);
      cudaDeviceSynchronize();
      long long c[16]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int w = 0; w < warps; ++w) mx = c[2 * w] > mx ? c[2 * w] : mx;
      const int bytes = (pat == 2 || pat == 3) ? 256 : 512;
      printf("%d warp(s)  %-42s %6.1f cycles per instruction per warp, %6.1f B/clk for the SM\n", warps, names[pat], (double)mx / N,
             (double)bytes * N * warps / mx);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
