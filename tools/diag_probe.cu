// Phase timing (clock64) of diag64_kernel on ONE 64 x 64 block: where do the 33 us of a single-matrix diagonal step go?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DNMGP_DIAG_PROF -o tools/diag_probe tools/diag_probe.cu
#include "../nonstationary_multivariate_gaussian_process_b200/csrc/diag.cu"
#include <vector>
namespace nmgp { void set_last_error(const std::string& m) { fprintf(stderr, "%s\n", m.c_str()); } }
using namespace nmgp;
int main() {
  const int n = 64;
  std::vector<double> A(n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) A[i * n + j] = exp(-0.5 * (i - j) * (i - j) / 400.0) + (i == j ? 0.1 : 0.0);
  BlockBatch b; b.n = n; b.nP = n; b.Kt = 1; b.NB = 64; b.batch = 1;
  cudaMalloc(&b.A, n * n * 8); cudaMalloc(&b.Dinv, 2 * n * n * 8); cudaMalloc(&b.logdet, 8); cudaMalloc(&b.info, 4);
  for (int rep = 0; rep < 5; ++rep) {
    cudaMemcpy(b.A, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); engine_diag_step(b, 0, 0, nullptr, false); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long st[64]; cudaMemcpyFromSymbol(st, g_diag_prof, sizeof(st));
    double ld; int info; cudaMemcpy(&ld, b.logdet, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&info, b.info, 4, cudaMemcpyDeviceToHost);
    printf("rep %d: event %.1f us | cycles: load %lld  chol %lld  storeL+logdet %lld  inverse %lld  sync %lld  total %lld | logdet %.6f info %d\n",
           rep, ms * 1e3, st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3], st[5] - st[4], st[5] - st[0], ld, info);
  }
  // the tensor-pipe kernel's phases (NMGP_DIAG_MMA unset or 1): lap counters 6 / 7 / 8 accumulate over the 8 tile columns
  {
    long long z[64] = {0};
    cudaMemcpyToSymbol(g_diag_prof, z, sizeof(z));
    cudaMemcpy(b.A, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    engine_diag_step(b, 0, 0, nullptr, false);
    cudaDeviceSynchronize();
    long long st[64]; cudaMemcpyFromSymbol(st, g_diag_prof, sizeof(st));
    printf("mma kernel cycles: load %lld | lookahead+factor|trailing %lld  panel %lld  (unused %lld) | storeL %lld  - %lld  rows+stats %lld  end-barrier %lld | total %lld\n",
           st[1] - st[0], st[6], st[7], st[8], st[3] - st[2], st[4] - st[3], st[5] - st[4], st[9] - st[5], st[9] - st[0]);
  }
  // the same with a WARM instruction cache: one CTA loops over 8 matrices, the stamps are those of the last one
  {
    const int batch = 8;
    BlockBatch c = b; c.batch = batch;
    cudaMalloc(&c.A, (size_t)batch * n * n * 8); cudaMalloc(&c.Dinv, (size_t)batch * 2 * n * n * 8);
    cudaMalloc(&c.logdet, (size_t)batch * 8); cudaMalloc(&c.info, (size_t)batch * 4);
    for (int m = 0; m < batch; ++m) cudaMemcpy(c.A + (size_t)m * n * n, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    DiagArgs g;
    g.A = c.A; g.Dinv = c.Dinv; g.logdet = c.logdet; g.info = c.info; g.pivmin = nullptr; g.pivmax = nullptr;
    g.strideA = c.strideA(); g.strideD = c.strideD(); g.ld = c.nP; g.batch = batch; g.step = 0;
    cudaFuncSetAttribute(diag64_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MMA_SMEM_BYTES);
    diag64_mma_kernel<<<1, MMA_THREADS, MMA_SMEM_BYTES>>>(g);
    cudaDeviceSynchronize();
    long long st[64]; cudaMemcpyFromSymbol(st, g_diag_prof, sizeof(st));
    printf("mma kernel, warm (8th matrix of one CTA): load %lld | lookahead+factor|trailing %lld  panel %lld  (unused %lld) | storeL %lld  - %lld  rows+stats %lld  end-barrier %lld | total %lld  (%s)\n",
           st[1] - st[0], st[6], st[7], st[8], st[3] - st[2], st[4] - st[3], st[5] - st[4], st[9] - st[5], st[9] - st[0], cudaGetErrorString(cudaGetLastError()));
    printf("   factor warp at p = 3 (tile column 4): look-ahead tile %lld  load+pivots %lld  V %lld  stores %lld  barrier %lld\n", st[10] - st[15], st[11] - st[10], st[12] - st[11], st[13] - st[12], st[14] - st[13]);
    printf("   per tile column p = -1..6: factor|trailing phase");
    for (int i = 0; i < 8; ++i) printf(" %lld", st[32 + i]);
    printf("   panel phase");
    for (int i = 0; i < 8; ++i) printf(" %lld", st[44 + i]);
    printf("\n");
    cudaMemcpy(c.A, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    for (int m = 0; m < batch; ++m) cudaMemcpy(c.A + (size_t)m * n * n, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    for (int m = 0; m < batch; ++m) cudaMemcpy(c.A + (size_t)m * n * n, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(diag64_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    diag64_kernel<false, true><<<1, 2 * THREADS, SMEM_BYTES>>>(g);
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(st, g_diag_prof, sizeof(st));
    printf("old pipelined kernel, warm: load %lld  chol+inverse %lld  total %lld  (%s)\n", st[1] - st[0], st[3] - st[1], st[5] - st[0], cudaGetErrorString(cudaGetLastError()));
    cudaFree(c.A); cudaFree(c.Dinv); cudaFree(c.logdet); cudaFree(c.info);
  }
  // throughput shape: many independent blocks
  for (int batch : {592, 4096, 16384}) {
    BlockBatch c = b; c.batch = batch;
    cudaMalloc(&c.A, (size_t)batch * n * n * 8); cudaMalloc(&c.Dinv, (size_t)batch * 2 * n * n * 8);
    cudaMalloc(&c.logdet, (size_t)batch * 8); cudaMalloc(&c.info, (size_t)batch * 4);
    for (int m = 0; m < batch; ++m) cudaMemcpy(c.A + (size_t)m * n * n, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
      for (int m = 0; m < batch; m += 1024) cudaMemcpy(c.A + (size_t)m * n * n, A.data(), n * n * 8, cudaMemcpyHostToDevice);
      cudaEventRecord(e0); engine_diag_step(c, 0, 0, nullptr, false); cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("batch %d rep %d: %.1f us = %.1f ns per block, %.2f us per block per SM\n", batch, rep, ms * 1e3, ms * 1e6 / batch, ms * 1e3 / batch * 148);
    }
    cudaFree(c.A); cudaFree(c.Dinv); cudaFree(c.logdet); cudaFree(c.info);
  }
  return 0;
}
