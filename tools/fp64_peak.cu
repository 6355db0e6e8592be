// FP64 pipe microbenchmark for B200 (sm_100a): DFMA vs mma.sync f64 shapes.
// Prints TFLOP/s per variant; the best sustained figure is the FP64 roofline denominator used in DESIGN.md.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double s) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  double a = s, b = 1.0 - s;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += acc[i];
  if (r == 123.456) out[0] = r;
}

template <int ILP>
__global__ void k_dmma884(double* out, int iters, double s) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = s, b = 1.0 - s;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += c[i][0] + c[i][1];
  if (r == 123.456) out[0] = r;
}

template <int ILP>
__global__ void k_dmma1684(double* out, int iters, double s) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double a0 = s, a1 = s * 0.5, b = 1.0 - s;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b));
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (r == 123.456) out[0] = r;
}

template <int ILP>
__global__ void k_dmma1688(double* out, int iters, double s) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double a0 = s, a1 = s * 0.5, a2 = s * 0.25, a3 = s * 0.125, b0 = 1.0 - s, b1 = 0.5 - s;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (r == 123.456) out[0] = r;
}

template <int ILP>
__global__ void k_dmma16816(double* out, int iters, double s) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = s * (i + 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = 1.0 - s * i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (r == 123.456) out[0] = r;
}

// SMEM-fed m8n8k4: a warp computes a 32x32 tile from As[32][KP], Bs[32][KP] (stride KP = 20 doubles, conflict-free).
__global__ void k_dmma884_smem(double* out, int iters) {
  constexpr int KT = 16, KP = 20;
  extern __shared__ double sm_dyn[];
  double (*As)[32 * KP] = reinterpret_cast<double (*)[32 * KP]>(sm_dyn);
  double (*Bs)[32 * KP] = reinterpret_cast<double (*)[32 * KP]>(sm_dyn + 8 * 32 * KP);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < 32 * KP; i += 32) { As[warp][i] = 1e-3 * i; Bs[warp][i] = 1e-4 * i; }
  __syncwarp();
  double c[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[i][j][0] = 0; c[i][j][1] = 0; }
  const double* Ap = &As[warp][(lane >> 2) * KP + (lane & 3)];
  const double* Bp = &Bs[warp][(lane >> 2) * KP + (lane & 3)];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < KT; k += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = Ap[i * 8 * KP + k]; b[i] = Bp[i * 8 * KP + k]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) r += c[i][j][0] + c[i][j][1];
  if (r == 123.456) out[0] = r;
}

static void row(const char* what, int threads, int cps, double flop, double ms) {
  if (ms < 0) printf("threads %4d cta/sm %d  %-16s launch failed (too many registers for this CTA size): skipped\n", threads, cps, what);
  else printf("threads %4d cta/sm %d  %-16s %7.2f TF/s\n", threads, cps, what, flop / ms * 1e-9);
}

// best-of-`reps` duration of f() in ms, or a negative value when the launch itself fails (e.g. 1024 threads x more than 64
// registers do not fit an SM): an unlaunched kernel must not be timed -- round 1's table printed 3e5..1e6 "TF/s" for those.
template <typename F>
static double time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f();
  if (cudaGetLastError() != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return -1.0; }
  f(); f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  printf("device %s sms %d\n", p.name, sms);
  double* out; CK(cudaMalloc(&out, 8));
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int cps : {1, 2}) {
      if (threads * cps > 2048) continue;
      int grid = sms * cps;
      double warps = (double)grid * threads / 32;
      double ms;
      ms = time_ms([&] { k_dfma<8><<<grid, threads>>>(out, iters, 0.5); }, 5);
      row("DFMA ilp8", threads, cps, 2.0 * 8 * iters * grid * threads, ms);
      ms = time_ms([&] { k_dmma884<8><<<grid, threads>>>(out, iters, 0.5); }, 5);
      row("DMMA m8n8k4 x8", threads, cps, 2.0 * 256 * 8 * iters * warps, ms);
      ms = time_ms([&] { k_dmma1684<8><<<grid, threads>>>(out, iters, 0.5); }, 5);
      row("DMMA m16n8k4 x8", threads, cps, 2.0 * 512 * 8 * iters * warps, ms);
      ms = time_ms([&] { k_dmma1688<8><<<grid, threads>>>(out, iters, 0.5); }, 5);
      row("DMMA m16n8k8 x8", threads, cps, 2.0 * 1024 * 8 * iters * warps, ms);
      ms = time_ms([&] { k_dmma16816<4><<<grid, threads>>>(out, iters, 0.5); }, 5);
      row("DMMA m16n8k16 x4", threads, cps, 2.0 * 2048 * 4 * iters * warps, ms);
    }
  }
  {
    CK(cudaFuncSetAttribute(k_dmma884_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 32 * 20 * 8));
    int grid = sms, threads = 256;
    double warps = (double)grid * threads / 32;
    double ms = time_ms([&] { k_dmma884_smem<<<grid, threads, 16 * 32 * 20 * 8>>>(out, 4000); }, 5);
    printf("SMEM-fed m8n8k4 32x32 warp tile, 8 warps/SM: %7.2f TF/s\n", 2.0 * 256 * 64 * 4000 * warps / ms * 1e-9);
    grid = sms * 2;
    warps = (double)grid * threads / 32;
    ms = time_ms([&] { k_dmma884_smem<<<grid, threads, 16 * 32 * 20 * 8>>>(out, 4000); }, 5);
    printf("SMEM-fed m8n8k4 32x32 warp tile, 16 warps/SM: %7.2f TF/s\n", 2.0 * 256 * 64 * 4000 * warps / ms * 1e-9);
  }
  // sustained: 3 s of the best variant, report clocks from nvidia-smi separately
  {
    int grid = sms * 2, threads = 512;
    double warps = (double)grid * threads / 32;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0;
    for (; n < 200; ++n) k_dmma884<8><<<grid, threads>>>(out, iters, 0.5);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("sustained DMMA m8n8k4 (%.1f s): %7.2f TF/s\n", ms * 1e-3, 2.0 * 256 * 8 * iters * warps * n / ms * 1e-9);
    CK(cudaEventRecord(e0));
    for (n = 0; n < 200; ++n) k_dfma<8><<<grid, threads>>>(out, iters, 0.5);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("sustained DFMA (%.1f s): %7.2f TF/s\n", ms * 1e-3, 2.0 * 8 * iters * (double)grid * threads * n / ms * 1e-9);
  }
  return 0;
}
