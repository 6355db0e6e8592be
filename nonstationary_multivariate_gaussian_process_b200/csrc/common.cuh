// Shared helpers for the NMGP B200 kernels (sm_100a, FP64).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>

namespace nmgp {

// -------------------------------------------------------------------------------- errors
void set_last_error(const std::string& msg);  // api.cu

#define NMGP_CUDA_TRY(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::nmgp::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " +    \
                             __FILE__ + ":" + std::to_string(__LINE__));                      \
      return -2;                                                                              \
    }                                                                                         \
  } while (0)

#define NMGP_TRY(expr)            \
  do {                            \
    int _r = (expr);              \
    if (_r != 0) return _r;       \
  } while (0)

// Function attributes (dynamic shared memory above 48 KB) belong to the (function, device) pair, not to the process: a
// plan on cuda:1 after one on cuda:0 must set them again.  One flag per device and call site; a lost race only repeats
// the (idempotent) call.  `kernel` may be a parenthesised template-id.
constexpr int kMaxDevices = 64;
#define NMGP_SMEM_ATTR_PER_DEVICE(kernel, bytes)                                                                  \
  do {                                                                                                            \
    static std::atomic<size_t> _nmgp_done[::nmgp::kMaxDevices];                                                   \
    int _dev = 0;                                                                                                 \
    NMGP_CUDA_TRY(cudaGetDevice(&_dev));                                                                          \
    const bool _slot = _dev >= 0 && _dev < ::nmgp::kMaxDevices;                                                   \
    if (!_slot || _nmgp_done[_dev].load(std::memory_order_acquire) < (size_t)(bytes)) {                           \
      NMGP_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));     \
      if (_slot) _nmgp_done[_dev].store((size_t)(bytes), std::memory_order_release);                              \
    }                                                                                                             \
  } while (0)

// Programmatic dependent launch (sm_90+): a kernel launched with launch_kernel_pdl(.., pdl = true) may be scheduled while
// the previous kernel of its stream still runs; it must call pdl_wait() before touching anything that kernel (or any earlier
// one) wrote.  pdl_launch_dependents() lets the NEXT kernel be scheduled early.  Both are no-ops for ordinary launches.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// number of SMs of the current device (148 on B200; never hard-coded in grid sizing)
inline int sm_count() {
  static std::atomic<int> cached[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < kMaxDevices) {
    const int c = cached[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (dev >= 0 && dev < kMaxDevices) cached[dev].store(n, std::memory_order_relaxed);
  return n;
}

constexpr double kJitter = 1e-6;  // Utility/settings.py:3

__host__ __device__ inline int tril_size(int M) { return M * (M + 1) / 2; }
__host__ __device__ inline long round_up(long a, long b) { return (a + b - 1) / b * b; }

// -------------------------------------------------------------------------------- FP64 tensor core
// One DMMA.8x8x4 (mma.sync m8n8k4, f64): D(8x8) += A(8x4, row) * B(4x8, col).
// Lane l holds  A[l>>2][l&3],  B[l&3][l>>2],  C[l>>2][2*(l&3) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// -------------------------------------------------------------------------------- covariance functions
// squared distance exactly as Utility/kernels.py:13-20 forms it: (x_i^2 + x_j^2) - 2.0*(x_i*x_j), no FMA contraction
__device__ __forceinline__ double ref_sqdist(double xi, double xj) {
  const double s = __dadd_rn(__dmul_rn(xi, xi), __dmul_rn(xj, xj));
  return __dsub_rn(s, __dmul_rn(2.0, __dmul_rn(xi, xj)));
}

// Gibbs kernel value without jitter and its log-derivative factor w.r.t. tilde_l_i
__device__ __forceinline__ void gibbs_pair(double xi, double xj, double li, double lj, double sij, double& k0,
                                           double& cfac) {
  const double d = ref_sqdist(xi, xj);
  const double li2 = __dmul_rn(li, li);
  const double A = __dadd_rn(li2, __dmul_rn(lj, lj));
  const double Bm = __dmul_rn(li, lj);
  const double root = sqrt(__ddiv_rn(__dmul_rn(2.0, Bm), A));
  const double e = exp(__ddiv_rn(-d, A));
  k0 = __dmul_rn(__dmul_rn(sij, root), e);            // (C*sqrt(2B/A))*exp(-d/A), kernels.py:72
  const double q = li2 / A;
  cfac = 0.5 - q + 2.0 * d * q / A;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum for blockDim.x <= 1024; result valid in every thread.
__device__ __forceinline__ double block_sum(double v, double* scratch /* >= 33 doubles */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < nw ? scratch[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

}  // namespace nmgp
