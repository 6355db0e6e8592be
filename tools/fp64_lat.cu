// Dependent-issue latencies (cycles) of the instructions on the pivot chain of the 64 x 64 diagonal-block kernel, one warp
// alone on an SM: DFMA, DMUL, DADD, MUFU.RSQ64H (rsqrt.approx.ftz.f64), DMMA.8x8x4 (accumulator chain and A-operand chain),
// SHFL.IDX of a double, LDS.64 pointer chase, FFMA for reference.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/fp64_lat tools/fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 2048;
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void lat(double* out, long long* cyc, double seed, int* chase) {
  __shared__ int sm[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = chase[i];
  __syncwarp();
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001, z = 1e-9;
  long long t0, t1;
  int k = 0;
  // DFMA
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) x = fma(x, y, z);
  t1 = clock64(); cyc[k++] = t1 - t0;
  // DMUL
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clock64(); cyc[k++] = t1 - t0;
  // DADD
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) x = x + z;
  t1 = clock64(); cyc[k++] = t1 - t0;
  // MUFU.RSQ64H chain (approx only; value converges to 1)
  x = fabs(x) + 1.0;
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x));
  t1 = clock64(); cyc[k++] = t1 - t0;
  // DMMA accumulator chain
  double c0 = x, c1 = y;
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) dmma(c0, c1, z, z);
  t1 = clock64(); cyc[k++] = t1 - t0;
  // DMMA A-operand chain (result feeds the next A)
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) { double d0 = 0.0, d1 = 0.0; asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(c0), "d"(z)); c0 = d0; }
  t1 = clock64(); cyc[k++] = t1 - t0;
  // SHFL of a double
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) c0 = __shfl_sync(0xffffffffu, c0, (threadIdx.x + 1) & 31);
  t1 = clock64(); cyc[k++] = t1 - t0;
  // LDS chase
  int p = threadIdx.x;
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) p = sm[p];
  t1 = clock64(); cyc[k++] = t1 - t0;
  // FFMA
  float f = (float)x, g = 1.0001f, h = 1e-6f;
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) f = fmaf(f, g, h);
  t1 = clock64(); cyc[k++] = t1 - t0;
  // 8 independent DFMA chains (issue interval)
  double v[8];
  for (int j = 0; j < 8; ++j) v[j] = x + j;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fma(v[j], y, z);
  t1 = clock64(); cyc[k++] = t1 - t0;
  for (int j = 0; j < 8; ++j) x += v[j];
  // 8 independent DMMA chains
  double w[8][2];
  for (int j = 0; j < 8; ++j) w[j][0] = w[j][1] = x + j;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma(w[j][0], w[j][1], z, z);
  t1 = clock64(); cyc[k++] = t1 - t0;
  for (int j = 0; j < 8; ++j) x += w[j][0] + w[j][1];
  out[threadIdx.x] = x + c0 + c1 + p + f;
}
int main() {
  double* out; long long* cyc; int* chase;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&chase, 1024 * 4);
  int h[1024]; for (int i = 0; i < 1024; ++i) h[i] = (i * 33 + 32) & 1023;
  cudaMemcpy(chase, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; ++rep) lat<<<1, 32>>>(out, cyc, 1.5, chase);
  cudaDeviceSynchronize();
  long long c[16]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const char* names[] = {"DFMA dependent", "DMUL dependent", "DADD dependent", "MUFU.RSQ64H dependent", "DMMA.8x8x4 accumulator chain",
                         "DMMA.8x8x4 result -> A operand", "SHFL.IDX double (2 x 32 bit)", "LDS.32 pointer chase", "FFMA dependent",
                         "DFMA, 8 independent chains (per instruction)", "DMMA, 8 independent chains (per instruction)"};
  for (int i = 0; i < 11; ++i) printf("%-48s %7.1f cycles\n", names[i], (double)c[i] / (i >= 9 ? 8.0 * N : (double)N));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
