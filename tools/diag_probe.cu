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
    long long st[16]; cudaMemcpyFromSymbol(st, g_diag_prof, sizeof(st));
    double ld; int info; cudaMemcpy(&ld, b.logdet, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&info, b.info, 4, cudaMemcpyDeviceToHost);
    printf("rep %d: event %.1f us | cycles: load %lld  chol %lld  storeL+logdet %lld  inverse %lld  sync %lld  total %lld | logdet %.6f info %d\n",
           rep, ms * 1e3, st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3], st[5] - st[4], st[5] - st[0], ld, info);
  }
  return 0;
}
