// Do DFMA (FP64 FMA pipe) and DMMA (mma.sync f64) share one execution resource on B200?
// Runs (a) DMMA only, (b) DFMA only, (c) warps alternating by parity DMMA / DFMA in the same CTA, (d) both instruction
// kinds interleaved in every warp, and prints the FP64 TFLOP/s of each part.  If (c)/(d) sum well above the single-kind
// peaks the two pipes are independent and FMA work can hide under tensor work.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// mode 0: all warps DMMA; 1: all DFMA; 2: even warps DMMA, odd warps DFMA; 3: every warp does both (R DFMA per DMMA)
template <int MODE, int R>
__global__ void k_mix(double* out, int iters, double s) {
  double c[8][2], f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; f[i] = threadIdx.x * 1e-3 + i; }
  const double a = s, b = 1.0 - s;
  const int warp = threadIdx.x >> 5;
  const bool do_mma = MODE == 0 || MODE == 3 || (MODE == 2 && (warp & 1) == 0);
  const bool do_fma = MODE == 1 || MODE == 3 || (MODE == 2 && (warp & 1) == 1);
  for (int it = 0; it < iters; ++it) {
    if (do_mma) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    if (do_fma) {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += c[i][0] + c[i][1] + f[i];
  if (r == 123.456) out[0] = r;
}

template <typename F>
static double time_ms(F f) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount, iters = 20000, threads = 512, grid = sms * 2;
  double* out; CK(cudaMalloc(&out, 8));
  const double warps = (double)grid * threads / 32;
  const double mma_fl = 2.0 * 256 * 8 * iters;      // per warp
  const double fma_fl = 2.0 * 32 * 8 * iters;       // per warp per R
  double ms;
  ms = time_ms([&] { k_mix<0, 1><<<grid, threads>>>(out, iters, 0.5); });
  printf("DMMA only             : %7.2f TF/s\n", mma_fl * warps / ms * 1e-9);
  ms = time_ms([&] { k_mix<1, 8><<<grid, threads>>>(out, iters, 0.5); });
  printf("DFMA only             : %7.2f TF/s\n", 8 * fma_fl * warps / ms * 1e-9);
  ms = time_ms([&] { k_mix<2, 8><<<grid, threads>>>(out, iters, 0.5); });
  printf("split warps (R=8)     : DMMA %7.2f + DFMA %7.2f = %7.2f TF/s\n", mma_fl * warps / 2 / ms * 1e-9,
         8 * fma_fl * warps / 2 / ms * 1e-9, (mma_fl + 8 * fma_fl) * warps / 2 / ms * 1e-9);
  ms = time_ms([&] { k_mix<3, 1><<<grid, threads>>>(out, iters, 0.5); });
  printf("interleaved 8:8 (R=1) : DMMA %7.2f + DFMA %7.2f = %7.2f TF/s\n", mma_fl * warps / ms * 1e-9,
         fma_fl * warps / ms * 1e-9, (mma_fl + fma_fl) * warps / ms * 1e-9);
  ms = time_ms([&] { k_mix<3, 4><<<grid, threads>>>(out, iters, 0.5); });
  printf("interleaved 8:32 (R=4): DMMA %7.2f + DFMA %7.2f = %7.2f TF/s\n", mma_fl * warps / ms * 1e-9,
         4 * fma_fl * warps / ms * 1e-9, (mma_fl + 4 * fma_fl) * warps / ms * 1e-9);
  ms = time_ms([&] { k_mix<3, 8><<<grid, threads>>>(out, iters, 0.5); });
  printf("interleaved 8:64 (R=8): DMMA %7.2f + DFMA %7.2f = %7.2f TF/s\n", mma_fl * warps / ms * 1e-9,
         8 * fma_fl * warps / ms * 1e-9, (mma_fl + 8 * fma_fl) * warps / ms * 1e-9);
  return 0;
}
