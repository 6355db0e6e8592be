// Live FP64 tensor-pipe probe: the roofline denominator of bench.py, measured in the same process, on the same device and
// under the same clocks as the numbers it is compared with (MEASURED_PEAKS.json carries no FP64 entry).  Register-only
// DMMA.8x8x4 (mma.sync.m8n8k4.f64, 8 independent accumulators per warp, 512 threads x 2 CTAs per SM: the shape that
// sustains the highest rate in tools/fp64_peak.cu) back to back for at least `min_seconds`, timed with CUDA events.
#include "../../include/nmgp_b200.h"

#include "common.cuh"

namespace nmgp {
namespace {
__global__ void __launch_bounds__(512, 2) dmma_probe_kernel(double* out, int iters, double s) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  const double a = s, b = 1.0 - s;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += c[i][0] + c[i][1];
  if (r == 123.456) out[0] = r;   // never true: keeps the loop alive
}
}  // namespace
}  // namespace nmgp

extern "C" int nmgp_fp64_dmma_probe(double min_seconds, double* tflops_out, double* seconds_out, void* stream) {
  using namespace nmgp;
  if (!tflops_out || !(min_seconds >= 0.0) || min_seconds > 10.0) { set_last_error("nmgp_fp64_dmma_probe: bad arguments"); return NMGP_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  double* out = nullptr;
  NMGP_CUDA_TRY(cudaMalloc(&out, sizeof(double)));
  const int grid = sm_count() * 2, threads = 512, iters = 20000;
  const double flop_per_launch = 2.0 * 256 * 8 * (double)iters * ((double)grid * threads / 32);   // 8x8x4 = 256 FMA per warp-MMA
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int rc = 0;
  dmma_probe_kernel<<<grid, threads, 0, st>>>(out, iters, 0.5);   // warm-up (module load, clocks)
  if (cudaGetLastError() != cudaSuccess) { set_last_error("nmgp_fp64_dmma_probe: launch failed"); rc = NMGP_ECUDA; }
  double total_ms = 0.0;
  long launches = 0;
  while (rc == 0 && total_ms < min_seconds * 1e3) {
    cudaEventRecord(e0, st);
    for (int k = 0; k < 4; ++k) dmma_probe_kernel<<<grid, threads, 0, st>>>(out, iters, 0.5);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) { set_last_error("nmgp_fp64_dmma_probe: kernel failed"); rc = NMGP_ECUDA; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    total_ms += ms;
    launches += 4;
  }
  if (rc == 0 && launches == 0) {   // min_seconds == 0: one timed batch
    cudaEventRecord(e0, st);
    dmma_probe_kernel<<<grid, threads, 0, st>>>(out, iters, 0.5);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    total_ms = ms;
    launches = 1;
  }
  if (rc == 0) {
    *tflops_out = flop_per_launch * (double)launches / (total_ms * 1e-3) * 1e-12;
    if (seconds_out) *seconds_out = total_ms * 1e-3;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return rc;
}
