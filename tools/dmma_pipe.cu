// What limits a streaming FP64 DMMA tile kernel on B200?  A 64x64 output tile per CTA accumulated over a long K loop,
// operands streamed from global memory through a shared-memory ring -- the structure of panel_gemm_kernel -- with the
// ingredients switched on one at a time:
//   LOAD   0: operands stay in shared memory (no global traffic)   1: cp.async ring from global (L2-resident panel)
//   SYNC   0: no CTA barrier per chunk (only legal with LOAD=0)     1: __syncthreads per chunk
//   WARPS  4: 4 warps x (32x32)   8: 8 warps x (32x16)   (64x64 CTA tile either way)
//   KC     k-chunk per stage
// Prints TFLOP/s for each variant at the occupancy the kernel gets from its own register / shared-memory use.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NB = 64;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(double* smem, const double* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

template <int LOAD, int SYNC, int WARPS, int KC, int STAGES, int MINB, int NOMMA = 0, int PF = 0>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_tile(const double* __restrict__ A, const double* __restrict__ B,
                                                           double* __restrict__ C, int nchunks, int ld) {
  extern __shared__ __align__(16) double smem[];
  constexpr int LDK = KC + 4, OPSZ = NB * LDK, THREADS = WARPS * 32;
  constexpr int WN = (WARPS == 4) ? 4 : 2;   // 8-column sub-tiles per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (WARPS == 4) ? (warp >> 1) * 32 : (warp >> 2) * 32;
  const int n0 = (WARPS == 4) ? (warp & 1) * 32 : (warp & 3) * 16;
  const double* At = A + (long)blockIdx.x * NB * ld;   // row panel of this CTA
  const double* Bt = B + (long)(blockIdx.x % 7) * NB * ld;
  double acc[4][WN][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto issue = [&](int q) {
    double* S = smem + (q % STAGES) * 2 * OPSZ;
    constexpr int V = KC / 2, ITERS = NB * V / THREADS;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int idx = threadIdx.x + it * THREADS;
      const int r = idx / V, c2 = idx % V;
      cp_async16(S + r * LDK + 2 * c2, At + (long)r * ld + q * KC + 2 * c2);
      cp_async16(S + OPSZ + r * LDK + 2 * c2, Bt + (long)r * ld + q * KC + 2 * c2);
    }
  };
  if (LOAD) {
#pragma unroll
    for (int q = 0; q < STAGES - 1; ++q) { if (q < nchunks) issue(q); cp_async_commit(); }
  } else {
    for (int i = threadIdx.x; i < STAGES * 2 * OPSZ; i += THREADS) smem[i] = 1e-3 * (i % 97);
    __syncthreads();
  }
  const int lr = lane >> 2, lk = lane & 3;
  for (int q = 0; q < nchunks; ++q) {
    if (LOAD) cp_async_wait<STAGES - 2>();
    if (SYNC) __syncthreads();
    if (LOAD) { if (q + STAGES - 1 < nchunks) issue(q + STAGES - 1); cp_async_commit(); }
    if (PF > 0 && q + PF < nchunks && threadIdx.x < NB) {   // one 128-byte line of the A panel per thread, PF chunks ahead
      const double* pa = At + (long)threadIdx.x * ld + (q + PF) * KC;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
    }
    if (NOMMA) { if (q == nchunks - 1) acc[0][0][0] = smem[(q % STAGES) * 2 * OPSZ + threadIdx.x]; continue; }
    const double* S = smem + (q % STAGES) * 2 * OPSZ;
    const double* pa = S + (m0 + lr) * LDK + lk;
    const double* pb = S + OPSZ + (n0 + lr) * LDK + lk;
#pragma unroll
    for (int k = 0; k < KC; k += 4) {
      double a[4], b[WN];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = pa[i * 8 * LDK + k];
#pragma unroll
      for (int j = 0; j < WN; ++j) b[j] = pb[j * 8 * LDK + k];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  if (LOAD) cp_async_wait<0>();
  double* Ct = C + (long)blockIdx.x * NB * NB;
  const int r = lane >> 2, c = 2 * (lane & 3);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j) {
      double2 v; v.x = acc[i][j][0]; v.y = acc[i][j][1];
      *reinterpret_cast<double2*>(Ct + (long)(m0 + 8 * i + r) * NB + n0 + 8 * j + c) = v;
    }
}

template <int LOAD, int SYNC, int WARPS, int KC, int STAGES, int MINB, int NOMMA = 0, int PF = 0>
void run(const char* name, const double* A, const double* B, double* C, int ntiles, int K, int ld) {
  auto kern = k_tile<LOAD, SYNC, WARPS, KC, STAGES, MINB, NOMMA, PF>;
  const size_t smem = (size_t)STAGES * 2 * NB * (KC + 4) * sizeof(double);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int nchunks = K / KC;
  for (int w = 0; w < 2; ++w) kern<<<ntiles, WARPS * 32, smem>>>(A, B, C, nchunks, ld);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    kern<<<ntiles, WARPS * 32, smem>>>(A, B, C, nchunks, ld);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  printf("%-58s occ %d CTA/SM  %7.2f TF/s  %6.2f TB/s of A-panel reads  (%.3f ms)\n", name, occ,
         2.0 * NB * NB * (double)K * ntiles / best * 1e-9, (double)ntiles * NB * K * 8 / best * 1e-9, best);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  const int ntiles = sms * 3 * 16, K = 512, ld = K;   // every CTA streams a 64 x 512 row panel (+ a shared B panel)
  double *A, *B, *C;
  CK(cudaMalloc(&A, (size_t)ntiles * NB * ld * 8)); CK(cudaMalloc(&B, (size_t)8 * NB * ld * 8)); CK(cudaMalloc(&C, (size_t)ntiles * NB * NB * 8));
  CK(cudaMemset(A, 0, (size_t)ntiles * NB * ld * 8)); CK(cudaMemset(B, 0, (size_t)8 * NB * ld * 8));
  printf("tiles %d, K %d: A panel data %.1f MB (HBM-resident: > L2)\n", ntiles, K, (double)ntiles * NB * ld * 8 / 1e6);
  run<0, 0, 4, 16, 3, 3>("smem only, no barrier, 4 warps, 3 CTA/SM", A, B, C, ntiles, K, ld);
  run<0, 1, 4, 16, 3, 3>("smem only, barrier per 16-chunk, 4 warps", A, B, C, ntiles, K, ld);
  run<1, 1, 4, 16, 3, 3>("cp.async ring KC16 x3, barrier, 4 warps (current)", A, B, C, ntiles, K, ld);
  run<1, 1, 4, 16, 4, 2>("cp.async ring KC16 x4, barrier, 4 warps, 2 CTA/SM", A, B, C, ntiles, K, ld);
  run<1, 1, 4, 32, 2, 3>("cp.async ring KC32 x2, barrier, 4 warps", A, B, C, ntiles, K, ld);
  run<1, 1, 4, 8, 4, 3>("cp.async ring KC8 x4, barrier, 4 warps", A, B, C, ntiles, K, ld);
  run<0, 0, 8, 16, 3, 2>("smem only, no barrier, 8 warps (32x16), 2 CTA/SM", A, B, C, ntiles, K, ld);
  run<0, 1, 8, 16, 3, 2>("smem only, barrier, 8 warps (32x16), 2 CTA/SM", A, B, C, ntiles, K, ld);
  run<1, 1, 8, 16, 3, 2>("cp.async ring KC16 x3, barrier, 8 warps, 2 CTA/SM", A, B, C, ntiles, K, ld);
  run<1, 1, 8, 16, 3, 3>("cp.async ring KC16 x3, barrier, 8 warps, 3 CTA/SM", A, B, C, ntiles, K, ld);
  run<1, 1, 8, 32, 2, 3>("cp.async ring KC32 x2, barrier, 8 warps, 3 CTA/SM", A, B, C, ntiles, K, ld);
  // streaming only (no MMA): what HBM rate can the ring sustain by itself, and does an L2 prefetch ahead of it help?
  run<1, 1, 4, 16, 3, 3, 1, 0>("stream only: ring KC16 x3, 3 CTA/SM", A, B, C, ntiles, K, ld);
  run<1, 1, 4, 16, 3, 3, 1, 8>("stream only: ring KC16 x3 + L2 prefetch 8 chunks ahead", A, B, C, ntiles, K, ld);
  run<1, 1, 4, 16, 3, 3, 1, 16>("stream only: ring KC16 x3 + L2 prefetch 16 chunks ahead", A, B, C, ntiles, K, ld);
  run<1, 1, 4, 32, 2, 3, 1, 0>("stream only: ring KC32 x2, 3 CTA/SM", A, B, C, ntiles, K, ld);
  run<1, 1, 4, 16, 4, 2, 1, 0>("stream only: ring KC16 x4, 2 CTA/SM", A, B, C, ntiles, K, ld);
  run<1, 1, 4, 16, 3, 3, 0, 8>("MMA + ring KC16 x3 + L2 prefetch 8 ahead", A, B, C, ntiles, K, ld);
  return 0;
}
