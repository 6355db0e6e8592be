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
#include <cuda.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NB = 64;

__global__ void fill_kernel(double* p, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    p[i] = (double)((i * 37 + (i >> 9) * 11) % 101 - 50) / 64.0;
}

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

// ---------------------------------------------------------------------------------------------- TMA-fed variant
// Same tile / warp layout, but the ring is filled by cp.async.bulk.tensor (one elected thread, mbarrier complete_tx) into
// DENSE 128-byte rows with the hardware 128B swizzle; the DMMA fragments are read conflict-free by permuting which four
// k's a k-step uses: {2s, 2s+1, 2s+8, 2s+9} (any permutation of k is legal as long as A and B agree).
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

template <int STAGES, int MINB>
__global__ void __launch_bounds__(128, MINB) k_tile_tma(const __grid_constant__ CUtensorMap mapA,
                                                        const __grid_constant__ CUtensorMap mapB, double* __restrict__ C,
                                                        int nchunks) {
  constexpr int KC = 16, OPB = NB * KC * 8;   // 8192 bytes per operand chunk, dense
  extern __shared__ __align__(1024) unsigned char tsm[];
  unsigned long long* full = reinterpret_cast<unsigned long long*>(tsm + STAGES * 2 * OPB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = (warp >> 1) * 32, n0 = (warp & 1) * 32;
  const int lr = lane >> 2, lk = lane & 3;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  auto issue = [&](int q) {
    const int s = q % STAGES;
    mbar_expect_tx(&full[s], 2 * OPB);
    tma_load_2d(tsm + s * 2 * OPB, &mapA, q * KC, (int)blockIdx.x * NB, &full[s]);
    tma_load_2d(tsm + s * 2 * OPB + OPB, &mapB, q * KC, (int)(blockIdx.x % 7) * NB, &full[s]);
  };
  if (threadIdx.x == 0)
    for (int q = 0; q < STAGES - 1 && q < nchunks; ++q) issue(q);
  // per-lane swizzled offsets (bytes) of the four k-steps:  slot = (s + 4*(lk>>1)) ^ lr, half = lk & 1
  int off[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) off[s] = (((s + 4 * (lk >> 1)) ^ lr) << 4) + (lk & 1) * 8;
  for (int q = 0; q < nchunks; ++q) {
    __syncthreads();                                    // everyone is done with chunk q-1: its slot may be refilled
    if (threadIdx.x == 0 && q + STAGES - 1 < nchunks) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(q + STAGES - 1);
    }
    mbar_wait(&full[q % STAGES], (q / STAGES) & 1);
    const unsigned char* SA = tsm + (q % STAGES) * 2 * OPB + (m0 + lr) * 128;
    const unsigned char* SB = tsm + (q % STAGES) * 2 * OPB + OPB + (n0 + lr) * 128;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = *reinterpret_cast<const double*>(SA + i * 1024 + off[s]);
        b[i] = *reinterpret_cast<const double*>(SB + i * 1024 + off[s]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  double* Ct = C + (long)blockIdx.x * NB * NB;
  const int r = lane >> 2, c = 2 * (lane & 3);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double2 v; v.x = acc[i][j][0]; v.y = acc[i][j][1];
      *reinterpret_cast<double2*>(Ct + (long)(m0 + 8 * i + r) * NB + n0 + 8 * j + c) = v;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn enc, void* base, long rows, long cols) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 8};
  cuuint32_t box[2] = {16, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return m;
}

template <int STAGES, int MINB>
void run_tma(const char* name, EncodeTiledFn enc, double* A, double* B, double* C, int ntiles, int K) {
  CUtensorMap mA = make_map(enc, A, (long)ntiles * NB, K), mB = make_map(enc, B, 8L * NB, K);
  auto kern = k_tile_tma<STAGES, MINB>;
  const size_t smem = (size_t)STAGES * 2 * NB * 16 * 8 + STAGES * 8 + 1024;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int nchunks = K / 16;
  for (int w = 0; w < 2; ++w) kern<<<ntiles, 128, smem>>>(mA, mB, C, nchunks);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    kern<<<ntiles, 128, smem>>>(mA, mB, C, nchunks);
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
  fill_kernel<<<1024, 256>>>(A, (long)ntiles * NB * ld); fill_kernel<<<64, 256>>>(B, 8L * NB * ld);
  CK(cudaDeviceSynchronize());
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
  {
    EncodeTiledFn enc = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
    if (!enc) { printf("cuTensorMapEncodeTiled unavailable\n"); return 0; }
    // correctness of the swizzled fragment addressing: A = row index pattern, compare one tile against the cp.async kernel
    // correctness of the swizzled, k-permuted fragment addressing: same tiles as the cp.async kernel
    const size_t cn = (size_t)ntiles * NB * NB;
    double* h1 = (double*)malloc(cn * 8); double* h2 = (double*)malloc(cn * 8);
    run<1, 1, 4, 16, 3, 3>("cp.async ring KC16 x3 (reference result)", A, B, C, ntiles, K, ld);
    CK(cudaMemcpy(h1, C, cn * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemset(C, 0, cn * 8));
    run_tma<3, 3>("TMA (128B swizzle) + mbarrier ring KC16 x3, 4 warps", enc, A, B, C, ntiles, K);
    CK(cudaMemcpy(h2, C, cn * 8, cudaMemcpyDeviceToHost));
    double md = 0, mx = 0;
    for (size_t i = 0; i < cn; ++i) { double d = h1[i] - h2[i]; if (d < 0) d = -d; if (d > md) md = d; double a = h1[i] < 0 ? -h1[i] : h1[i]; if (a > mx) mx = a; }
    printf("TMA vs cp.async result: max |diff| = %.3e (max |C| = %.3e)\n", md, mx);
    run_tma<4, 3>("TMA (128B swizzle) + mbarrier ring KC16 x4, 4 warps", enc, A, B, C, ntiles, K);
    run_tma<6, 3>("TMA (128B swizzle) + mbarrier ring KC16 x6, 4 warps", enc, A, B, C, ntiles, K);
  }
  return 0;
}
