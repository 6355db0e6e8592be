// C ABI of libnmgp_b200.so (include/nmgp_b200.h): plan = cached loop-invariant state + workspace; one call =
// one batched log-posterior + gradient evaluation.  No torch types, no exceptions across the boundary.
#include "../../include/nmgp_b200.h"

#include <cmath>
#include <new>
#include <vector>

#include "common.cuh"
#include "engine.cuh"
#include "models.cuh"

namespace nmgp {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
}  // namespace nmgp

using namespace nmgp;

struct nmgp_plan {
  int model = 0, S = 0, N = 0, M = 0, T = 0, P = 0;
  int nmat = 1;   // matrices per subject (1 nonseparable, M separable/stationary)
  int n = 0;      // dimension of each matrix
  int nP = 0, Kt = 0;
  int chunk = 0;  // subjects per pass
  int nprior = 0; // GP-prior covariance matrices per subject (0 stationary, 2 otherwise)
  int nv1 = 1;    // right-hand sides for prior 1
  int engine_mode = 0;  // 0 auto, 1 force right-looking tile tasks, 2 force left-looking (tests / A-B timing)
  int stable_inverse = 0;  // 1: W^T W inverse even where the Takahashi sweep would be allowed (engine mode 3)
  double hyper[NMGP_NHYPER] = {0};
  HyperConst hc{};
  std::vector<void*> allocs;
  size_t dev_bytes = 0;
  long last_launches = 0;
  // persistent
  double *x = nullptr, *Y = nullptr;
  int* indx = nullptr;                      // [S][N] output index of every observation (Hadamard objectives only)
  double *Wp0 = nullptr, *Wp1 = nullptr;    // [S][N][N] lower Cholesky factors of the prior covariances
  double *hld0 = nullptr, *hld1 = nullptr;  // [S] half log-determinants
  // chunk workspace
  BlockBatch bb;
  void* maps = nullptr;      // TMA descriptors of the workspace (engine_maps_create); rebuilt when A2 appears
  double* A2_own = nullptr;  // second matrix buffer, allocated on first use by an inverse that needs it (not in `allocs`)
  Scratch w{};
  // staging for the host-buffer call
  double *pars_d = nullptr, *vals_d = nullptr, *grad_d = nullptr;
  int* info_d = nullptr;
  // nmgp_hyper_grad: traces of the prior precisions (depend on x and the hyper-parameters only; formed on the first call)
  double *trI0 = nullptr, *trB0 = nullptr, *trI1 = nullptr, *trB1 = nullptr;  // [S]
  double *hq0 = nullptr, *hq1 = nullptr;                                       // [chunk][prior_quad_blocks(N)][4]
  bool traces_ready = false;
  double* trace_scratch = nullptr;   // kept between calls: a tied-hyper-prior loop re-forms the traces every iteration
  size_t trace_cap = 0;              // subjects per pass through trace_scratch
  // CUDA-graph replay of launch-bound evaluations (single-chunk plans with few launches; see evaluate_replay)
  struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    const double* pars = nullptr;
    double *vals = nullptr, *grad = nullptr;
    int* info = nullptr;
    long launches = 0;
    int warm = 0;        // evaluations of this variant launched directly so far (lazy module loads, helper streams)
    int captures = 0;    // re-captures after a pointer change; capped, then the variant launches directly for good
  };
  GraphSlot gslot[2];    // [grad requested]
  cudaStream_t gcap = nullptr;   // capture stream (torch's current stream is usually the legacy stream, which cannot capture)
  int graph_mode = 0;    // 0 auto, 1 never
  long graph_replays = 0;
  // lazily allocated scratch of the prediction entry points (not part of `allocs`)
  double* pred_scratch = nullptr;
  size_t pred_scratch_doubles = 0;
  // side stream: the GP-prior triangular solves are independent of the factorisation and run beside it
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // second side stream: the two priors' solves are independent chains of small kernels (single-subject plans: the longest
  // chain of the whole evaluation), so each prior gets its own stream
  cudaStream_t side2 = nullptr;
  cudaEvent_t ev_join2 = nullptr;
  // host-buffer call: copy streams and per-chunk events, so that the H2D of chunk c+1 and the D2H of chunk c-1 run
  // under the evaluation of chunk c
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev_start = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_done;
};

// optional per-chunk synchronisation of evaluate(): wait for before[c] on the compute stream, record after[c] on it
struct ChunkSync {
  cudaEvent_t* before = nullptr;
  cudaEvent_t* after = nullptr;
};

namespace {

template <typename Tp>
int dev_alloc(nmgp_plan* pl, Tp** out, size_t count) {
  *out = nullptr;
  if (count == 0) count = 1;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, count * sizeof(Tp));
  if (e != cudaSuccess) {
    set_last_error(std::string("cudaMalloc of ") + std::to_string(count * sizeof(Tp)) + " bytes failed: " +
                   cudaGetErrorString(e));
    cudaGetLastError();
    return NMGP_ENOMEM;
  }
  pl->allocs.push_back(p);
  pl->dev_bytes += count * sizeof(Tp);
  *out = reinterpret_cast<Tp*>(p);
  return 0;
}

HyperConst make_hyper_const(int model, const double* hy, int prior) {
  HyperConst h{};
  if (model >= NMGP_HADAMARD) model = model == NMGP_HADAMARD ? NMGP_SEPARABLE : (model == NMGP_HADAMARD_SVC ? NMGP_NONSEPARABLE : NMGP_STATIONARY);
  h.prior = prior ? 1 : 0;
  h.half_log2pi = std::log(std::sqrt(2.0 * M_PI));
  double a = 1, b = 1, c = 10;
  if (model == NMGP_STATIONARY) {
    const float loc = (float)hy[0], sc = (float)hy[1];     // Normal(mu_tilde_l, sigma_tilde_l), logpos.py:446
    h.s_loc = (double)loc;
    h.s_var = (double)(sc * sc);
    h.s_logscale = (double)logf(sc);
    a = hy[2]; b = hy[3]; c = hy[4];
  } else if (model == NMGP_SEPARABLE) {
    h.mu0 = hy[0]; h.mu1 = hy[3];
    a = hy[6]; b = hy[7]; c = hy[8];
  } else {
    h.mu0 = hy[0]; h.mu1 = hy[3];
    a = hy[6]; b = hy[7];
  }
  const float c32 = (float)c;                               // Normal(0, c), logpos.py:283,450
  h.n_loc = 0.0;
  h.n_var = (double)(c32 * c32);
  h.n_logscale = (double)logf(c32);
  h.ig_a = a; h.ig_b = b;
  h.ig_alogb = a * std::log(b);
  h.ig_lgamma = std::lgamma(a);
  return h;
}

size_t per_subject_bytes(const nmgp_plan* pl) {
  const size_t N = pl->N, M = pl->M, n = pl->n, nm = pl->nmat, nP = pl->nP, Kt = pl->Kt;
  const size_t MT = pl->model == NMGP_NONSEPARABLE ? padded_M(pl->M) : 0;
  const bool had = pl->model >= NMGP_HADAMARD;
  size_t d = had ? 6 * N * 16 : 0;     // row factors (both layouts) and the four gradient tables of the Hadamard objectives
  if (had) d += 3 * N * (size_t)pl->nv1;
  d += 2 * nm * nP * nP;        // A, A2
  d += 3 * nm * Kt * kNB * kNB; // Dinv (W and W^T), Pbuf
  d += nm;                      // logdet
  d += 2 * N + 1;               // ell, sig, s2
  d += pl->model == NMGP_NONSEPARABLE ? n * MT : M * M;  // Lst
  d += 2 * N * N;               // Kx, CK
  d += 2 * nm * n;              // alpha, yv
  d += pl->model == NMGP_NONSEPARABLE ? 2 * n * MT : 2 * N * M;  // Wout, Vout
  d += 2 * N + M + M * M + 3 * N * MT;  // gl, gs, lam, Vec, Sa, Ca, Ua
  d += 3 * (N + N * (size_t)pl->nv1);  // R, Z, G
  return d * sizeof(double) + nm * sizeof(int);
}

int alloc_workspace(nmgp_plan* pl) {
  const size_t cs = pl->chunk, N = pl->N, M = pl->M, n = pl->n, nm = pl->nmat;
  const bool svc = pl->model == NMGP_NONSEPARABLE;
  const bool had = pl->model >= NMGP_HADAMARD;
  const size_t MT = svc ? padded_M(pl->M) : 0;
  BlockBatch& b = pl->bb;
  b.n = pl->n; b.nP = pl->nP; b.Kt = pl->Kt; b.NB = kNB; b.batch = (int)(cs * nm);
  NMGP_TRY(dev_alloc(pl, &b.A, cs * nm * (size_t)pl->nP * pl->nP));
  b.A2 = nullptr;   // allocated by the first evaluation that takes an inverse path with a W^T W stage: ensure_second_buffer()
  NMGP_TRY(dev_alloc(pl, &b.Dinv, cs * nm * (size_t)pl->Kt * 2 * kNB * kNB));
  NMGP_TRY(dev_alloc(pl, &b.Pbuf, cs * nm * (size_t)pl->Kt * kNB * kNB));
  // the left-looking engine multiplies padding / not-yet-written upper tiles by exact zeros: they must be finite
  NMGP_CUDA_TRY(cudaMemsetAsync(b.A, 0, cs * nm * (size_t)pl->nP * pl->nP * sizeof(double), 0));
  NMGP_CUDA_TRY(cudaMemsetAsync(b.Pbuf, 0, cs * nm * (size_t)pl->Kt * kNB * kNB * sizeof(double), 0));
  NMGP_CUDA_TRY(cudaStreamSynchronize(0));
  NMGP_TRY(dev_alloc(pl, &b.logdet, cs * nm));
  NMGP_TRY(dev_alloc(pl, &b.info, cs * nm));
  NMGP_TRY(dev_alloc(pl, &b.pivmin, cs * nm));
  NMGP_TRY(dev_alloc(pl, &b.pivmax, cs * nm));
  NMGP_TRY(dev_alloc(pl, &b.sched, 2));
  NMGP_CUDA_TRY(cudaMemsetAsync(b.sched, 0, 2 * sizeof(int), 0));
  NMGP_TRY(engine_maps_create(b, &pl->maps));
  b.maps = pl->maps;
  Scratch& w = pl->w;
  NMGP_TRY(dev_alloc(pl, &w.ell, cs * N));
  NMGP_TRY(dev_alloc(pl, &w.sig, cs * N));
  NMGP_TRY(dev_alloc(pl, &w.s2, cs));
  NMGP_TRY(dev_alloc(pl, &w.Lst, svc ? cs * n * MT : (had ? cs * N * 16 : cs * M * M)));
  NMGP_TRY(dev_alloc(pl, &w.Kx, cs * N * N));
  NMGP_TRY(dev_alloc(pl, &w.CK, cs * N * N));
  NMGP_TRY(dev_alloc(pl, &w.alpha, cs * nm * n));
  NMGP_TRY(dev_alloc(pl, &w.yv, (svc || had) ? 1 : cs * nm * n));
  NMGP_TRY(dev_alloc(pl, &w.Wout, svc ? cs * n * MT : (had ? cs * N * 16 : cs * N * M)));
  NMGP_TRY(dev_alloc(pl, &w.Vout, svc ? cs * n * MT : (had ? cs * N * 16 : cs * N * M)));
  NMGP_TRY(dev_alloc(pl, &w.Sa, svc ? cs * N * MT : (had ? cs * N * 16 : 1)));
  NMGP_TRY(dev_alloc(pl, &w.Ca, svc ? cs * N * MT : (had ? cs * N * 16 : 1)));
  NMGP_TRY(dev_alloc(pl, &w.Ua, svc ? cs * N * MT : (had ? cs * N * 16 : 1)));
  NMGP_TRY(dev_alloc(pl, &w.gl, cs * N));
  NMGP_TRY(dev_alloc(pl, &w.gs, cs * N));
  NMGP_TRY(dev_alloc(pl, &w.lam, cs * M));
  NMGP_TRY(dev_alloc(pl, &w.Vec, cs * M * M));
  const size_t nv1 = pl->nv1;
  NMGP_TRY(dev_alloc(pl, &w.R0, cs * N));
  NMGP_TRY(dev_alloc(pl, &w.R1, cs * N * nv1));
  NMGP_TRY(dev_alloc(pl, &w.Z0, cs * N));
  NMGP_TRY(dev_alloc(pl, &w.Z1, cs * N * nv1));
  NMGP_TRY(dev_alloc(pl, &w.G0, cs * N));
  NMGP_TRY(dev_alloc(pl, &w.G1, cs * N * nv1));
  return 0;
}

// Factor the GP-prior covariance alpha^2 exp(-0.5 d/beta^2) + 1e-6 I of every subject once (the reference rebuilds and
// re-factors it on every call, logpos.py:271-281, 357-365): L = chol(Sigma_p) and 0.5 log det.
int factor_prior(nmgp_plan* pl, double alpha, double beta, double* Wp, double* hld, cudaStream_t st) {
  BlockBatch pb = pl->bb;  // reuse the likelihood workspace
  pb.maps = nullptr;       // different block layout: the plan's TMA descriptors do not describe it
  pb.n = pl->N;
  pb.nP = padded_dim(pl->N);
  pb.Kt = pb.nP / kNB;
  const size_t capA = (size_t)pl->bb.batch * pl->bb.strideA();
  const size_t capD = (size_t)pl->bb.batch * pl->bb.strideD();
  size_t cap = capA / (size_t)pb.strideA();
  if (capD / (size_t)pb.strideD() < cap) cap = capD / (size_t)pb.strideD();
  if ((size_t)pl->bb.batch < cap) cap = pl->bb.batch;   // logdet / info arrays
  if (cap > 65535) cap = 65535;
  if (cap == 0) { set_last_error("factor_prior: workspace too small"); return NMGP_ENOMEM; }
  for (int s0 = 0; s0 < pl->S; s0 += (int)cap) {
    const int cs = pl->S - s0 < (int)cap ? pl->S - s0 : (int)cap;
    pb.batch = cs;
    NMGP_TRY(launch_prior_cov_blocks(pl->x + (size_t)s0 * pl->N, cs, pl->N, alpha, beta, pb, st, nullptr));
    NMGP_TRY(engine_potrf(pb, st, nullptr, /*stable_panel=*/true));
    NMGP_TRY(launch_extract_factor(pb, pl->N, Wp + (size_t)s0 * pl->N * pl->N, hld + s0, st, nullptr));
  }
  return 0;
}

// torch.optim.Adam's update, element-wise over all subjects (same operation order as torch's single-tensor path:
// exp_avg.lerp_(grad, 1-beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2);
// denom = exp_avg_sq.sqrt() / sqrt(bias_correction2) + eps; param.addcdiv_(exp_avg, denom, -lr / bias_correction1))
__global__ void __launch_bounds__(256) adam_kernel(double* __restrict__ pars, const double* __restrict__ grad,
                                                   double* __restrict__ m, double* __restrict__ v,
                                                   const int* __restrict__ info, const unsigned char* __restrict__ frozen,
                                                   long S, long P, double beta1, double beta2, double eps,
                                                   double step_size, double sqrt_bc2) {
  const long total = S * P;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long s = idx / P, j = idx - s * P;
    if ((info && info[s] != 0) || (frozen && frozen[j])) continue;
    const double g = grad[idx];
    const double mi = m[idx] + (g - m[idx]) * (1.0 - beta1);
    const double vi = v[idx] * beta2 + (1.0 - beta2) * g * g;
    m[idx] = mi;
    v[idx] = vi;
    const double denom = sqrt(vi) / sqrt_bc2 + eps;
    pars[idx] -= step_size * (mi / denom);
  }
}

// ---- Hamiltonian Monte Carlo pieces (the drivers' HMC loops call an external sampler with potential_func = nlogpos_obj*:
// Separable_model.py:209-210, Nonseparable_model_mpiKAISER.py:267-270): leapfrog updates and the accept step for all subjects
__global__ void __launch_bounds__(256) hmc_kick_kernel(double* __restrict__ p, const double* __restrict__ grad,
                                                       const int* __restrict__ info, long S, long P, double step) {
  const long total = S * P;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    if (info && info[idx / P] != 0) continue;
    p[idx] -= step * grad[idx];
  }
}

__global__ void __launch_bounds__(256) hmc_drift_kernel(double* __restrict__ q, const double* __restrict__ p, long total,
                                                        double eps) {
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x)
    q[idx] += eps * p[idx];
}

// one CTA per subject: dH = (U + |p0|^2/2) - (U' + |p1|^2/2); accept iff the proposal is valid and log u < dH
__global__ void __launch_bounds__(256) hmc_accept_kernel(double* __restrict__ q, const double* __restrict__ q_prop,
                                                         double* __restrict__ grad, const double* __restrict__ grad_prop,
                                                         double* __restrict__ U, const double* __restrict__ vals_prop,
                                                         const double* __restrict__ p0, const double* __restrict__ p1,
                                                         const int* __restrict__ failed, const double* __restrict__ log_u,
                                                         int* __restrict__ accepted, long P) {
  __shared__ double scratch[40];
  __shared__ int acc_s;
  const long s = blockIdx.x;
  double k0 = 0.0, k1 = 0.0;
  for (long j = threadIdx.x; j < P; j += blockDim.x) {
    const double a = p0[s * P + j], b = p1[s * P + j];
    k0 += a * a;
    k1 += b * b;
  }
  const double K0 = 0.5 * block_sum(k0, scratch);
  const double K1 = 0.5 * block_sum(k1, scratch);
  if (threadIdx.x == 0) {
    const double Up = vals_prop[s * NMGP_NVALS];
    const double dH = (U[s] + K0) - (Up + K1);
    const bool ok = (!failed || failed[s] == 0) && Up == Up && dH == dH && log_u[s] < dH;
    acc_s = ok ? 1 : 0;
    accepted[s] = acc_s;
    if (ok) U[s] = Up;
  }
  __syncthreads();
  if (acc_s) {
    for (long j = threadIdx.x; j < P; j += blockDim.x) {
      q[s * P + j] = q_prop[s * P + j];
      grad[s * P + j] = grad_prop[s * P + j];
    }
  }
}

__global__ void pack_kernel(const double* __restrict__ src, int n, double* __restrict__ dst, long strideA, int ld) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  const int b = blockIdx.z;
  if (q >= ld) return;
  double v;
  if (p < n && q < n) v = (q <= p) ? src[((long)b * n + p) * n + q] : src[((long)b * n + q) * n + p];
  else v = (p == q) ? 1.0 : 0.0;
  dst[(long)b * strideA + (long)p * ld + q] = v;
}

__global__ void unpack_kernel(const double* __restrict__ src, long strideA, int ld, int n, double* __restrict__ dst,
                              int lower_only) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  const int b = blockIdx.z;
  if (q >= n) return;
  if (lower_only && q > p) return;
  dst[((long)b * n + p) * n + q] = src[(long)b * strideA + (long)p * ld + q];
}

int potrf_unit(double* A, int n, int batch, double* logdet, int* info, int invert, cudaStream_t st) {
  if (n <= 0 || batch < 0 || !A) { set_last_error("potrf: bad arguments"); return NMGP_EINVAL; }
  if (batch == 0) return 0;
  BlockBatch b;
  b.n = n; b.nP = padded_dim(n); b.Kt = b.nP / kNB; b.NB = kNB;
  long step = (long)((8ull << 30) / ((size_t)b.strideA() * sizeof(double)));
  if (step < 1) step = 1;
  if (step > 4096) step = 4096;
  if (step > batch) step = batch;
  double *ws = nullptr, *dinv = nullptr, *ld = nullptr;
  int* inf = nullptr;
  NMGP_CUDA_TRY(cudaMalloc(&ws, (size_t)step * b.strideA() * sizeof(double)));
  cudaError_t e1 = cudaMalloc(&dinv, (size_t)step * b.strideD() * sizeof(double));
  cudaError_t e2 = cudaMalloc(&ld, (size_t)step * sizeof(double));
  cudaError_t e3 = cudaMalloc(&inf, (size_t)step * sizeof(int));
  int rc = 0;
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { set_last_error("potrf: out of memory"); rc = NMGP_ENOMEM; }
  for (int b0 = 0; rc == 0 && b0 < batch; b0 += (int)step) {
    const int cs = batch - b0 < step ? batch - b0 : (int)step;
    b.A = ws; b.Dinv = dinv; b.logdet = ld; b.info = inf; b.batch = cs;
    dim3 grid((b.nP + 127) / 128, b.nP, cs);
    pack_kernel<<<grid, 128, 0, st>>>(A + (size_t)b0 * n * n, n, ws, b.strideA(), b.nP);
    rc = engine_potrf(b, st, nullptr);
    if (rc == 0 && invert) rc = engine_potri(b, st, nullptr);
    if (rc != 0) break;
    dim3 g2((n + 127) / 128, n, cs);
    unpack_kernel<<<g2, 128, 0, st>>>(ws, b.strideA(), b.nP, n, A + (size_t)b0 * n * n, invert ? 0 : 1);
    if (logdet) cudaMemcpyAsync(logdet + b0, ld, cs * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (info) cudaMemcpyAsync(info + b0, inf, cs * sizeof(int), cudaMemcpyDeviceToDevice, st);
    if (cudaGetLastError() != cudaSuccess) { set_last_error("potrf: launch failed"); rc = NMGP_ECUDA; }
  }
  cudaStreamSynchronize(st);
  cudaFree(ws); cudaFree(dinv); cudaFree(ld); cudaFree(inf);
  return rc;
}

}  // namespace

// =================================================================================================== C ABI
extern "C" {

const char* nmgp_last_error(void) { return g_last_error.c_str(); }

int nmgp_n_params(int model, int N, int M) {
  if (N <= 0 || M <= 0) return NMGP_EINVAL;
  const int T = M * (M + 1) / 2;
  switch (model) {
    case NMGP_STATIONARY: return T + 3;
    case NMGP_SEPARABLE: return 2 * N + T + 1;
    case NMGP_NONSEPARABLE: return N + N * T + 1;
    case NMGP_HADAMARD: return 2 * N + T + 1;       // vec2pars(pars, N, M), logpos.py:479
    case NMGP_HADAMARD_SVC: return N + N * T + 1;   // vec2pars_hadamard_SVC, logpos.py:60-72
    case NMGP_HADAMARD_S: return T + 3;             // vec2pars_S, logpos.py:657
    default: return NMGP_EINVAL;
  }
}

int nmgp_plan_destroy(nmgp_plan* pl) {
  if (!pl) return 0;
  engine_maps_destroy(pl->maps);
  if (pl->A2_own) cudaFree(pl->A2_own);
  for (void* p : pl->allocs) cudaFree(p);
  if (pl->pred_scratch) cudaFree(pl->pred_scratch);
  if (pl->trace_scratch) cudaFree(pl->trace_scratch);
  for (auto& g : pl->gslot)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (pl->gcap) cudaStreamDestroy(pl->gcap);
  if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
  if (pl->ev_join) cudaEventDestroy(pl->ev_join);
  if (pl->side) cudaStreamDestroy(pl->side);
  if (pl->ev_join2) cudaEventDestroy(pl->ev_join2);
  if (pl->side2) cudaStreamDestroy(pl->side2);
  if (pl->ev_start) cudaEventDestroy(pl->ev_start);
  for (cudaEvent_t e : pl->ev_in) cudaEventDestroy(e);
  for (cudaEvent_t e : pl->ev_done) cudaEventDestroy(e);
  if (pl->copy_in) cudaStreamDestroy(pl->copy_in);
  if (pl->copy_out) cudaStreamDestroy(pl->copy_out);
  delete pl;
  return 0;
}

static int plan_create_impl(nmgp_plan** out, int model, int S, int N, int M, const double* x_dev, const double* Y_dev,
                            const int* indx_dev, const double* hyper, int prior_flag, size_t workspace_limit_bytes,
                            void* stream);

int nmgp_plan_create(nmgp_plan** out, int model, int S, int N, int M, const double* x_dev, const double* Y_dev,
                     const double* hyper, int prior_flag, size_t workspace_limit_bytes, void* stream) {
  if (!out) return NMGP_EINVAL;
  *out = nullptr;
  if (model < 0 || model > 2 || S < 0 || N <= 0 || M <= 0 || M > 16 || !hyper || (S > 0 && (!x_dev || !Y_dev))) {
    set_last_error("nmgp_plan_create: bad arguments (need 0<=model<=2, S>=0, N>=1, 1<=M<=16, non-null x/Y/hyper)");
    return NMGP_EINVAL;
  }
  return plan_create_impl(out, model, S, N, M, x_dev, Y_dev, nullptr, hyper, prior_flag, workspace_limit_bytes, stream);
}

int nmgp_plan_create_hadamard(nmgp_plan** out, int model, int S, int N, int M, const double* x_dev, const int* indx_dev,
                              const double* y_dev, const double* hyper, int prior_flag, size_t workspace_limit_bytes,
                              void* stream) {
  if (!out) return NMGP_EINVAL;
  *out = nullptr;
  if (model < NMGP_HADAMARD || model > NMGP_HADAMARD_S || S < 0 || N <= 0 || M <= 0 || M > 16 || !hyper ||
      (S > 0 && (!x_dev || !indx_dev || !y_dev))) {
    set_last_error("nmgp_plan_create_hadamard: bad arguments (need 3<=model<=5, S>=0, N>=1, 1<=M<=16, non-null x/indx/y/hyper)");
    return NMGP_EINVAL;
  }
  return plan_create_impl(out, model, S, N, M, x_dev, y_dev, indx_dev, hyper, prior_flag, workspace_limit_bytes, stream);
}

static int plan_create_impl(nmgp_plan** out, int model, int S, int N, int M, const double* x_dev, const double* Y_dev,
                            const int* indx_dev, const double* hyper, int prior_flag, size_t workspace_limit_bytes,
                            void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  nmgp_plan* pl = new (std::nothrow) nmgp_plan();
  if (!pl) return NMGP_ENOMEM;
  pl->model = model; pl->S = S; pl->N = N; pl->M = M;
  pl->T = M * (M + 1) / 2;
  pl->P = nmgp_n_params(model, N, M);
  std::memcpy(pl->hyper, hyper, sizeof(pl->hyper));
  pl->hc = make_hyper_const(model, hyper, prior_flag);
  const bool svc = model == NMGP_NONSEPARABLE;
  const bool had = model >= NMGP_HADAMARD;          // one N x N matrix per subject, one observation per row
  const int ycols = had ? 1 : M;
  pl->nmat = (svc || had) ? 1 : M;
  pl->n = svc ? N * M : N;
  pl->nP = padded_dim(pl->n);
  pl->Kt = pl->nP / kNB;
  pl->nprior = (model == NMGP_STATIONARY || model == NMGP_HADAMARD_S) ? 0 : 2;
  pl->nv1 = (svc || model == NMGP_HADAMARD_SVC) ? pl->T : 1;
  int rc = 0;
  do {
    if (S == 0) { pl->chunk = 0; break; }
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { set_last_error("cudaMemGetInfo failed (no CUDA device?)"); rc = NMGP_ECUDA; break; }
    // persistent state
    const size_t SN = (size_t)S * N;
    if ((rc = dev_alloc(pl, &pl->x, SN))) break;
    if ((rc = dev_alloc(pl, &pl->Y, SN * ycols))) break;
    if (cudaMemcpyAsync(pl->x, x_dev, SN * sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(pl->Y, Y_dev, SN * ycols * sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_last_error("nmgp_plan_create: copying x/Y failed"); rc = NMGP_ECUDA; break;
    }
    if (had) {
      if ((rc = dev_alloc(pl, &pl->indx, SN))) break;
      if (cudaMemcpyAsync(pl->indx, indx_dev, SN * sizeof(int), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        set_last_error("nmgp_plan_create_hadamard: copying indx failed"); rc = NMGP_ECUDA; break;
      }
    }
    if (pl->nprior) {
      if ((rc = dev_alloc(pl, &pl->Wp0, SN * N))) break;
      if ((rc = dev_alloc(pl, &pl->Wp1, SN * N))) break;
      if ((rc = dev_alloc(pl, &pl->hld0, (size_t)S))) break;
      if ((rc = dev_alloc(pl, &pl->hld1, (size_t)S))) break;
    }
    if ((rc = dev_alloc(pl, &pl->pars_d, (size_t)S * pl->P))) break;
    if ((rc = dev_alloc(pl, &pl->grad_d, (size_t)S * pl->P))) break;
    if ((rc = dev_alloc(pl, &pl->vals_d, (size_t)S * NMGP_NVALS))) break;
    if ((rc = dev_alloc(pl, &pl->info_d, (size_t)S))) break;
    // chunking
    size_t limit = workspace_limit_bytes;
    if (limit == 0) {
      const size_t used = pl->dev_bytes;
      const size_t avail = free_b > used ? free_b - used : 0;
      limit = (size_t)(0.7 * (double)avail);
      const size_t cap = 96ull << 30;
      if (limit > cap) limit = cap;
    }
    const size_t per = per_subject_bytes(pl);
    long chunk = (long)(limit / per);
    if (chunk < 1) chunk = 1;
    if (chunk > S) chunk = S;
    const long maxc = 65535 / pl->nmat;
    if (chunk > maxc) chunk = maxc;
    // Nothing is gained beyond ~4000 matrices per launch (9+ waves of CTAs); equal chunks of at most that size keep
    // the workspace small and let the host-buffer call pipeline its copies against the evaluation.
    static const long kMaxMat = getenv("NMGP_CHUNK_MATRICES") ? atol(getenv("NMGP_CHUNK_MATRICES")) : 4096;   // A/B timing
    if (chunk * pl->nmat > kMaxMat) {
      const long per_chunk = kMaxMat / pl->nmat > 0 ? kMaxMat / pl->nmat : 1;
      const long nch = (S + per_chunk - 1) / per_chunk;
      chunk = (S + nch - 1) / nch;
    }
    pl->chunk = (int)chunk;
    if ((rc = alloc_workspace(pl))) break;
    if (pl->nprior) {
      const double a0 = hyper[1], b0 = hyper[2], a1 = hyper[4], b1 = hyper[5];
      if ((rc = factor_prior(pl, a0, b0, pl->Wp0, pl->hld0, st))) break;
      if ((rc = factor_prior(pl, a1, b1, pl->Wp1, pl->hld1, st))) break;
    }
    {
      const int nch = (S + pl->chunk - 1) / pl->chunk;
      bool ok = cudaStreamCreateWithFlags(&pl->copy_in, cudaStreamNonBlocking) == cudaSuccess &&
                cudaStreamCreateWithFlags(&pl->copy_out, cudaStreamNonBlocking) == cudaSuccess &&
                cudaEventCreateWithFlags(&pl->ev_start, cudaEventDisableTiming) == cudaSuccess;
      for (int c = 0; ok && c < nch; ++c) {
        cudaEvent_t e1 = nullptr, e2 = nullptr;
        ok = cudaEventCreateWithFlags(&e1, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&e2, cudaEventDisableTiming) == cudaSuccess;
        if (e1) pl->ev_in.push_back(e1);
        if (e2) pl->ev_done.push_back(e2);
      }
      if (!ok) { set_last_error("nmgp_plan_create: creating the copy streams failed"); rc = NMGP_ECUDA; break; }
    }
    if (pl->nprior) {
      if (cudaStreamCreateWithFlags(&pl->side, cudaStreamNonBlocking) != cudaSuccess ||
          cudaStreamCreateWithFlags(&pl->side2, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&pl->ev_join2, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&pl->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        set_last_error("nmgp_plan_create: creating the side stream failed"); rc = NMGP_ECUDA; break;
      }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) {
      set_last_error(std::string("nmgp_plan_create: ") + cudaGetErrorString(cudaGetLastError())); rc = NMGP_ECUDA; break;
    }
  } while (0);
  if (rc != 0) { nmgp_plan_destroy(pl); return rc; }
  *out = pl;
  return 0;
}

// Engine selection shared by the evaluation and the prediction path.
static bool use_left_looking(const nmgp_plan* pl, const BlockBatch& b) {
  return b.Pbuf != nullptr && (pl->engine_mode == 2 || ((pl->engine_mode == 0 || pl->engine_mode == 4) && prefer_left_looking(b)));
}
static int run_potrf(nmgp_plan* pl, const BlockBatch& b, cudaStream_t st, long* launches) {
  return use_left_looking(pl, b) ? engine_potrf_ll(b, st, launches) : engine_potrf(b, st, launches);
}
// Inverse from the factor.  Three formulations, chosen by regime (measured on B200, profiles/r01_inverse_ab.txt):
//  * Takahashi sweep (engine_potri_ll, behind the conditioning guard of engine_potri_ll_guarded: an ill-conditioned matrix
//    -- small noise variance against the signal -- is routed to the W^T W path on the device): fewest bytes per matrix; best
//    for thousands of mid-size matrices (10 000 x n = 600:
//    19.5 ms vs 21.2 ms per 3334).  It propagates the error of the trailing inverse block into every new block column,
//    multiplied by |L(B,j) L_jj^-1| -- for smooth GP covariances that factor exceeds 1 and the error grows geometrically
//    with the number of block columns (exact at Kt = 16, 5e-7 at Kt = 32, garbage at Kt = 64) -- so only up to
//    kTakahashiMaxBlocks block columns.
//  * level-synchronous recursive triangular inverse + W^T W (engine_potri_ll_recursive): backward stable, 2 log2(Kt) + 1
//    launches of independent long-K tiles; everything else (n = 5000 single: 26 TFLOP/s instead of 10.9 for the tile tasks;
//    n = 16 384: 34 TFLOP/s single and batched).
//  * right-looking tile tasks (engine_potri): unit entry points without the second matrix buffer, and engine mode 1.
constexpr int kTakahashiMinBatch = 1024;
// The W^T W inverses keep W^T in a second matrix buffer.  It is allocated when an evaluation first takes such a path (never
// inside a graph capture: the first evaluation of every variant launches directly) and stays with the plan.
static int ensure_second_buffer(nmgp_plan* pl, BlockBatch& b, cudaStream_t st) {
  if (!pl->bb.A2) {
    const size_t bytes = (size_t)pl->bb.batch * pl->bb.strideA() * sizeof(double);
    if (cudaMalloc(&pl->A2_own, bytes) != cudaSuccess) {
      cudaGetLastError();
      pl->A2_own = nullptr;
      set_last_error("inverse: cudaMalloc of " + std::to_string(bytes) + " bytes for the second matrix buffer failed");
      return NMGP_ENOMEM;
    }
    // multiplied by exact zeros before it is written: must be finite.  Stream-ordered on the evaluation's own stream (a plain
    // cudaMemset runs on the legacy stream, which a non-blocking caller stream does not wait for).
    NMGP_CUDA_TRY(cudaMemsetAsync(pl->A2_own, 0, bytes, st));
    pl->dev_bytes += bytes;
    pl->bb.A2 = pl->A2_own;
    engine_maps_destroy(pl->maps);
    pl->maps = nullptr;
    pl->bb.maps = nullptr;
    NMGP_TRY(engine_maps_create(pl->bb, &pl->maps));
    pl->bb.maps = pl->maps;
  }
  b.A2 = pl->bb.A2;
  b.maps = pl->bb.maps;
  return 0;
}
static int run_potri(nmgp_plan* pl, BlockBatch& b, cudaStream_t st, long* launches) {
  const bool ll = use_left_looking(pl, b);
  const bool has_p = b.Pbuf != nullptr;
  switch (pl->engine_mode) {
    case 1: return engine_potri(b, st, launches);
    case 2:
      if (!ll) return engine_potri(b, st, launches);
      NMGP_TRY(ensure_second_buffer(pl, b, st));
      if (pl->stable_inverse || b.Kt > kTakahashiMaxBlocks) return engine_potri_ll_stable(b, st, launches);
      return engine_potri_ll_guarded(b, st, launches);
    case 4:
      if (!has_p) return engine_potri(b, st, launches);
      NMGP_TRY(ensure_second_buffer(pl, b, st));
      return engine_potri_ll_recursive(b, st, launches);
    default: break;
  }
  if (ll && b.Kt <= kTakahashiMaxBlocks && b.batch >= kTakahashiMinBatch) {
    NMGP_TRY(ensure_second_buffer(pl, b, st));
    return engine_potri_ll_guarded(b, st, launches);
  }
  if (has_p && b.Kt > 1) {
    NMGP_TRY(ensure_second_buffer(pl, b, st));
    return engine_potri_ll_recursive(b, st, launches);
  }
  return engine_potri(b, st, launches);
}

// digamma(x), x > 0: recurrence up to x >= 6, then the asymptotic series (error < 1e-15 there)
static double digamma_pos(double x) {
  double r = 0.0;
  while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  return r + std::log(x) - 0.5 / x -
         f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132 - f * (691.0 / 32760 - f / 12))))));
}

static int ensure_prior_traces(nmgp_plan* pl, cudaStream_t st) {
  if (pl->traces_ready || !pl->nprior || pl->S == 0) return 0;
  const size_t S = pl->S, N = pl->N;
  if (!pl->trI0) {
    NMGP_TRY(dev_alloc(pl, &pl->trI0, S));
    NMGP_TRY(dev_alloc(pl, &pl->trB0, S));
    NMGP_TRY(dev_alloc(pl, &pl->trI1, S));
    NMGP_TRY(dev_alloc(pl, &pl->trB1, S));
  }
  const size_t per = 3 * N * N + 2 * N;                    // doubles of scratch per subject (prior_traces)
  if (!pl->trace_scratch) {
    size_t cap = ((size_t)512 << 20) / (per * sizeof(double));
    if (cap < 1) cap = 1;
    if (cap > S) cap = S;
    if (cap > 65535) cap = 65535;
    if (cudaMalloc(&pl->trace_scratch, cap * per * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      pl->trace_scratch = nullptr;
      set_last_error("nmgp_hyper_grad: out of device memory for the trace scratch");
      return NMGP_ENOMEM;
    }
    pl->trace_cap = cap;
    pl->dev_bytes += cap * per * sizeof(double);
  }
  const size_t cap = pl->trace_cap;
  double* scratch = pl->trace_scratch;
  // stream-ordered on `st`: no host synchronisation, the scratch stays with the plan
  for (size_t s0 = 0; s0 < S; s0 += cap) {
    const int cs = (int)(S - s0 < cap ? S - s0 : cap);
    NMGP_TRY(prior_traces(pl->x + s0 * N, pl->Wp0 + s0 * N * N, cs, (int)N, pl->hyper[1], pl->hyper[2], scratch, pl->trI0 + s0,
                          pl->trB0 + s0, st, nullptr));
    NMGP_TRY(prior_traces(pl->x + s0 * N, pl->Wp1 + s0 * N * N, cs, (int)N, pl->hyper[4], pl->hyper[5], scratch, pl->trI1 + s0,
                          pl->trB1 + s0, st, nullptr));
  }
  pl->traces_ready = true;
  return 0;
}

static int hyper_setup(nmgp_plan* pl, cudaStream_t st, HyperRaw* h) {
  for (int k = 0; k < NMGP_NHYPER; ++k) h->hy[k] = pl->hyper[k];
  h->a = pl->hc.ig_a;
  h->b = pl->hc.ig_b;
  h->digamma_a = digamma_pos(h->a);
  h->prior = pl->hc.prior;
  if (h->prior) NMGP_TRY(ensure_prior_traces(pl, st));
  if (pl->nprior && !pl->hq0) {
    NMGP_TRY(dev_alloc(pl, &pl->hq0, (size_t)pl->chunk * prior_quad_blocks(pl->N) * 4));
    NMGP_TRY(dev_alloc(pl, &pl->hq1, (size_t)pl->chunk * prior_quad_blocks(pl->N) * 4));
  }
  return 0;
}

// the hyper-parameter gradient of chunk [s0, s0 + cs) once Z0/Z1/G0/G1 and s2 of that chunk are in the scratch
static int hyper_chunk(nmgp_plan* pl, const HyperRaw& h, int s0, int cs, const double* ps, double* hgrad, cudaStream_t st,
                       long* launches) {
  const int N = pl->N;
  const double* xs = pl->x + (size_t)s0 * N;
  const Scratch& w = pl->w;
  double* hq0 = pl->hq0;
  double* hq1 = pl->hq1;
  if (pl->nprior && h.prior) {
    NMGP_TRY(launch_prior_quad(xs, w.Z0, w.G0, cs, N, 1, pl->hyper[1], pl->hyper[2], hq0, st, launches));
    NMGP_TRY(launch_prior_quad(xs, w.Z1, w.G1, cs, N, pl->nv1, pl->hyper[4], pl->hyper[5], hq1, st, launches));
  }
  return launch_hyper_finish(pl->model, cs, N, pl->M, pl->P, ps, h, w.s2, hq0, hq1,
                             pl->trI0 ? pl->trI0 + s0 : nullptr, pl->trB0 ? pl->trB0 + s0 : nullptr,
                             pl->trI1 ? pl->trI1 + s0 : nullptr, pl->trB1 ? pl->trB1 + s0 : nullptr, pl->nv1,
                             hgrad + (size_t)s0 * NMGP_NHYPER, st, launches);
}

static int evaluate(nmgp_plan* pl, const double* pars, double* vals, double* grad, int* info, cudaStream_t st,
                    float* phase_ms, const ChunkSync* sync = nullptr, double* hgrad = nullptr) {
  long launches = 0;
  const int N = pl->N, M = pl->M, P = pl->P;
  HyperRaw hraw{};
  if (hgrad) NMGP_TRY(hyper_setup(pl, st, &hraw));
  cudaEvent_t ev[NMGP_NPHASES + 1];
  if (phase_ms) {
    for (int i = 0; i <= NMGP_NPHASES; ++i) NMGP_CUDA_TRY(cudaEventCreate(&ev[i]));
    for (int i = 0; i < NMGP_NPHASES; ++i) phase_ms[i] = 0.f;
  }
#define NMGP_MARK(i) do { if (phase_ms) NMGP_CUDA_TRY(cudaEventRecord(ev[i], st)); } while (0)
  for (int s0 = 0; s0 < pl->S; s0 += pl->chunk) {
    const int cs = pl->S - s0 < pl->chunk ? pl->S - s0 : pl->chunk;
    BlockBatch b = pl->bb;
    b.batch = cs * pl->nmat;
    const double* xs = pl->x + (size_t)s0 * N;
    const bool had = pl->model >= NMGP_HADAMARD;
    const double* Ys = pl->Y + (size_t)s0 * N * (had ? 1 : M);
    const int* ixs = had ? pl->indx + (size_t)s0 * N : nullptr;
    const double* ps = pars + (size_t)s0 * P;
    double* vs = vals + (size_t)s0 * NMGP_NVALS;
    double* gs = grad ? grad + (size_t)s0 * P : nullptr;
    int* is = info + s0;
    const int ci = s0 / pl->chunk;
    if (sync && sync->before) NMGP_CUDA_TRY(cudaStreamWaitEvent(st, sync->before[ci], 0));
    NMGP_MARK(0);
    if (had) {
      NMGP_TRY(had_forward(pl->model - NMGP_HADAMARD, cs, N, M, xs, ixs, ps, P, pl->hc, pl->w, b, st, &launches));
    } else if (pl->model == NMGP_NONSEPARABLE) {
      NMGP_TRY(svc_forward(cs, N, M, xs, ps, P, pl->hc, pl->w, b, st, &launches));
    } else {
      NMGP_TRY(sep_forward(pl->model, cs, N, M, xs, Ys, ps, P, pl->hc, pl->w, b, st, &launches));
    }
    NMGP_MARK(1);
    // The prior solves only need the residuals written by the forward pass.  In the normal call they run on the plan's
    // side stream beside the factorisation (fork/join with events, so the call stays stream-ordered on `st`); the
    // profiling call keeps them on `st` so that every phase is timed alone.
    const bool overlap = pl->nprior && !phase_ms && pl->side && pl->side2;
    cudaStream_t ps_st = overlap ? pl->side : st;     // prior 0, then the hyper-parameter gradient
    cudaStream_t ps_st1 = overlap ? pl->side2 : st;   // prior 1
    auto prior_solves = [&]() -> int {
      // the hyper-parameter gradient needs only the prep outputs and these solves: it rides on the side stream too,
      // beside the factorisation
      if (!pl->nprior) return hgrad ? hyper_chunk(pl, hraw, s0, cs, ps, hgrad, ps_st, &launches) : 0;
      const double* L0 = pl->Wp0 + (size_t)s0 * N * N;
      const double* L1 = pl->Wp1 + (size_t)s0 * N * N;
      const bool back = grad || hgrad;
      NMGP_TRY(launch_prior_solve(L0, pl->w.R0, pl->w.Z0, cs, N, 1, 0, ps_st, &launches));
      if (back) NMGP_TRY(launch_prior_solve(L0, pl->w.Z0, pl->w.G0, cs, N, 1, 1, ps_st, &launches));
      NMGP_TRY(launch_prior_solve(L1, pl->w.R1, pl->w.Z1, cs, N, pl->nv1, 0, ps_st1, &launches));
      if (back) NMGP_TRY(launch_prior_solve(L1, pl->w.Z1, pl->w.G1, cs, N, pl->nv1, 1, ps_st1, &launches));
      if (overlap) {
        NMGP_CUDA_TRY(cudaEventRecord(pl->ev_join2, pl->side2));
        NMGP_CUDA_TRY(cudaStreamWaitEvent(pl->side, pl->ev_join2, 0));
      }
      if (hgrad) NMGP_TRY(hyper_chunk(pl, hraw, s0, cs, ps, hgrad, ps_st, &launches));
      return 0;
    };
    if (overlap) {
      NMGP_CUDA_TRY(cudaEventRecord(pl->ev_fork, st));
      NMGP_CUDA_TRY(cudaStreamWaitEvent(pl->side, pl->ev_fork, 0));
      NMGP_CUDA_TRY(cudaStreamWaitEvent(pl->side2, pl->ev_fork, 0));
      NMGP_TRY(prior_solves());
      NMGP_CUDA_TRY(cudaEventRecord(pl->ev_join, pl->side));
    }
    NMGP_TRY(run_potrf(pl, b, st, &launches));
    NMGP_MARK(2);
    NMGP_TRY(run_potri(pl, b, st, &launches));
    NMGP_MARK(3);
    if (overlap) NMGP_CUDA_TRY(cudaStreamWaitEvent(st, pl->ev_join, 0));
    else NMGP_TRY(prior_solves());
    NMGP_MARK(4);
    const double* h0 = pl->hld0 ? pl->hld0 + s0 : nullptr;
    const double* h1 = pl->hld1 ? pl->hld1 + s0 : nullptr;
    if (had) {
      NMGP_TRY(had_backward(pl->model - NMGP_HADAMARD, cs, N, M, Ys, ixs, ps, P, pl->hc, pl->w, b, h0, h1, vs, gs, is, st,
                            &launches));
    } else if (pl->model == NMGP_NONSEPARABLE) {
      NMGP_TRY(svc_backward(cs, N, M, Ys, ps, P, pl->hc, pl->w, b, h0, h1, vs, gs, is, st, &launches));
    } else {
      NMGP_TRY(sep_backward(pl->model, cs, N, M, ps, P, pl->hc, pl->w, b, h0, h1, vs, gs, is, st, &launches));
    }
    NMGP_MARK(5);
    if (sync && sync->after) NMGP_CUDA_TRY(cudaEventRecord(sync->after[ci], st));
    if (phase_ms) {
      NMGP_CUDA_TRY(cudaEventSynchronize(ev[NMGP_NPHASES]));
      for (int i = 0; i < NMGP_NPHASES; ++i) {
        float ms = 0.f;
        NMGP_CUDA_TRY(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
        phase_ms[i] += ms;
      }
    }
  }
#undef NMGP_MARK
  if (phase_ms)
    for (int i = 0; i <= NMGP_NPHASES; ++i) cudaEventDestroy(ev[i]);
  pl->last_launches = launches;
  return 0;
}

static void drop_graphs(nmgp_plan* pl) {
  for (auto& g : pl->gslot) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
    g.warm = 0;
    g.captures = 0;
  }
}

// Launch-bound plans (one chunk, a few dozen to a few hundred small kernels: the drivers' one-subject-per-process MAP / HMC
// loops, C1 / C2 of BASELINE.json) replay the evaluation as ONE CUDA graph: the kernel sequence of evaluate() -- including the
// fork to the side stream and the look-ahead helper stream -- is captured once per (pars, vals, grad, info) pointer tuple on
// the plan's own capture stream and launched with cudaGraphLaunch on the caller's stream.  The first evaluation of each
// variant always launches directly (lazy module loading and the helper streams are not capturable); a failed capture or a
// caller whose buffers keep moving falls back to direct launches of the same kernels -- never to anything else.
// Cap of 256 kernels: measured (profiles/r01_latency_graph.txt) the replay gains 15-25 % up to ~60 kernels and nothing on the
// chain-bound large-matrix evaluations (n = 5000: 355 kernels, 8.34 vs 8.39 ms; n = 16 384: 1072 kernels, 144 vs 148 ms),
// whose look-ahead relies on stream priorities.
static long graph_max_launches() {
  static const long v = getenv("NMGP_GRAPH_MAX_LAUNCHES") ? atol(getenv("NMGP_GRAPH_MAX_LAUNCHES")) : 256;   // A/B timing
  return v;
}
constexpr int kGraphMaxCaptures = 4;
static int evaluate_replay(nmgp_plan* pl, const double* pars, double* vals, double* grad, int* info, cudaStream_t st) {
  nmgp_plan::GraphSlot& g = pl->gslot[grad ? 1 : 0];
  const bool eligible = pl->graph_mode == 0 && pl->S > 0 && pl->S <= pl->chunk && g.captures <= kGraphMaxCaptures &&
                        (g.warm == 0 || g.launches <= graph_max_launches());
  if (!eligible || g.warm == 0) {
    const int rc = evaluate(pl, pars, vals, grad, info, st, nullptr);
    if (rc == 0) { g.warm++; g.launches = pl->last_launches; }
    return rc;
  }
  if (g.exec && (g.pars != pars || g.vals != vals || g.grad != grad || g.info != info)) {
    cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
  }
  if (!g.exec) {
    if (++g.captures > kGraphMaxCaptures) return evaluate(pl, pars, vals, grad, info, st, nullptr);
    if (!pl->gcap && cudaStreamCreateWithFlags(&pl->gcap, cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      pl->graph_mode = 1;
      return evaluate(pl, pars, vals, grad, info, st, nullptr);
    }
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(pl->gcap, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      const int rc = evaluate(pl, pars, vals, grad, info, pl->gcap, nullptr);
      const cudaError_t ec = cudaStreamEndCapture(pl->gcap, &graph);
      ok = rc == 0 && ec == cudaSuccess && graph != nullptr;
    }
    if (ok) ok = cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      cudaGetLastError();
      g.exec = nullptr;
      pl->graph_mode = 1;   // this plan launches directly from now on
      return evaluate(pl, pars, vals, grad, info, st, nullptr);
    }
    g.pars = pars; g.vals = vals; g.grad = grad; g.info = info;
    g.launches = pl->last_launches;
  }
  NMGP_CUDA_TRY(cudaGraphLaunch(g.exec, st));
  pl->last_launches = g.launches;
  pl->graph_replays++;
  return 0;
}

int nmgp_logpost_grad(nmgp_plan* pl, const double* pars, double* vals, double* grad, int* info, void* stream) {
  if (!pl || (pl->S > 0 && (!pars || !vals || !info))) { set_last_error("nmgp_logpost_grad: null argument"); return NMGP_EINVAL; }
  return evaluate_replay(pl, pars, vals, grad, info, (cudaStream_t)stream);
}

int nmgp_logpost_grad_profile(nmgp_plan* pl, const double* pars, double* vals, double* grad, int* info, float* phase_ms,
                              void* stream) {
  if (!pl || !phase_ms || (pl->S > 0 && (!pars || !vals || !info))) { set_last_error("nmgp_logpost_grad_profile: null argument"); return NMGP_EINVAL; }
  return evaluate(pl, pars, vals, grad, info, (cudaStream_t)stream, phase_ms);
}

int nmgp_logpost_grad_host(nmgp_plan* pl, const double* pars_h, double* vals_h, double* grad_h, int* info_h,
                           void* stream) {
  if (!pl || (pl->S > 0 && (!pars_h || !vals_h))) { set_last_error("nmgp_logpost_grad_host: null argument"); return NMGP_EINVAL; }
  if (pl->S == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = pl->P;
  const int nch = (pl->S + pl->chunk - 1) / pl->chunk;
  if (nch == 1 && pl->graph_mode == 0) {
    // one chunk: nothing to pipeline -- copy in, evaluate (graph replay when launch-bound), copy out, all on `st`
    const size_t S = pl->S;
    NMGP_CUDA_TRY(cudaMemcpyAsync(pl->pars_d, pars_h, S * P * sizeof(double), cudaMemcpyHostToDevice, st));
    NMGP_TRY(evaluate_replay(pl, pl->pars_d, pl->vals_d, grad_h ? pl->grad_d : nullptr, pl->info_d, st));
    NMGP_CUDA_TRY(cudaMemcpyAsync(vals_h, pl->vals_d, S * NMGP_NVALS * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (grad_h) NMGP_CUDA_TRY(cudaMemcpyAsync(grad_h, pl->grad_d, S * P * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (info_h) NMGP_CUDA_TRY(cudaMemcpyAsync(info_h, pl->info_d, S * sizeof(int), cudaMemcpyDeviceToHost, st));
    NMGP_CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
  }
  // fork the copy streams off `st`, so the call stays ordered after whatever the caller queued there
  NMGP_CUDA_TRY(cudaEventRecord(pl->ev_start, st));
  NMGP_CUDA_TRY(cudaStreamWaitEvent(pl->copy_in, pl->ev_start, 0));
  NMGP_CUDA_TRY(cudaStreamWaitEvent(pl->copy_out, pl->ev_start, 0));
  for (int c = 0; c < nch; ++c) {   // all H2D copies are queued up front: chunk c+1 arrives while chunk c is evaluated
    const size_t s0 = (size_t)c * pl->chunk;
    const size_t cs = (size_t)pl->S - s0 < (size_t)pl->chunk ? (size_t)pl->S - s0 : (size_t)pl->chunk;
    NMGP_CUDA_TRY(cudaMemcpyAsync(pl->pars_d + s0 * P, pars_h + s0 * P, cs * P * sizeof(double), cudaMemcpyHostToDevice,
                                  pl->copy_in));
    NMGP_CUDA_TRY(cudaEventRecord(pl->ev_in[c], pl->copy_in));
  }
  ChunkSync sync;
  sync.before = pl->ev_in.data();
  sync.after = pl->ev_done.data();
  NMGP_TRY(evaluate(pl, pl->pars_d, pl->vals_d, grad_h ? pl->grad_d : nullptr, pl->info_d, st, nullptr, &sync));
  for (int c = 0; c < nch; ++c) {   // D2H of chunk c as soon as its evaluation is done, under the evaluation of c+1
    const size_t s0 = (size_t)c * pl->chunk;
    const size_t cs = (size_t)pl->S - s0 < (size_t)pl->chunk ? (size_t)pl->S - s0 : (size_t)pl->chunk;
    NMGP_CUDA_TRY(cudaStreamWaitEvent(pl->copy_out, pl->ev_done[c], 0));
    NMGP_CUDA_TRY(cudaMemcpyAsync(vals_h + s0 * NMGP_NVALS, pl->vals_d + s0 * NMGP_NVALS, cs * NMGP_NVALS * sizeof(double),
                                  cudaMemcpyDeviceToHost, pl->copy_out));
    if (grad_h)
      NMGP_CUDA_TRY(cudaMemcpyAsync(grad_h + s0 * P, pl->grad_d + s0 * P, cs * P * sizeof(double), cudaMemcpyDeviceToHost,
                                    pl->copy_out));
    if (info_h)
      NMGP_CUDA_TRY(cudaMemcpyAsync(info_h + s0, pl->info_d + s0, cs * sizeof(int), cudaMemcpyDeviceToHost, pl->copy_out));
  }
  NMGP_CUDA_TRY(cudaStreamSynchronize(pl->copy_out));
  NMGP_CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

int nmgp_hyper_grad(nmgp_plan* pl, const double* pars, double* hgrad, void* stream) {
  if (!pl || (pl->S > 0 && (!pars || !hgrad))) { set_last_error("nmgp_hyper_grad: null argument"); return NMGP_EINVAL; }
  if (pl->model >= NMGP_HADAMARD) { set_last_error("nmgp_hyper_grad: not available for the Hadamard objectives"); return NMGP_EINVAL; }
  if (pl->S == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int N = pl->N, M = pl->M, P = pl->P;
  long launches = 0;
  HyperRaw h{};
  NMGP_TRY(hyper_setup(pl, st, &h));
  for (int s0 = 0; s0 < pl->S; s0 += pl->chunk) {
    const int cs = pl->S - s0 < pl->chunk ? pl->S - s0 : pl->chunk;
    const double* ps = pars + (size_t)s0 * P;
    if (pl->model == NMGP_NONSEPARABLE) NMGP_TRY(launch_svc_prep(cs, N, M, ps, P, pl->hc, pl->w, st, &launches));
    else NMGP_TRY(launch_sep_prep(pl->model, cs, N, M, pl->Y + (size_t)s0 * N * M, ps, P, pl->hc, pl->w, st, &launches));
    if (pl->nprior && h.prior) {
      const double* L0 = pl->Wp0 + (size_t)s0 * N * N;
      const double* L1 = pl->Wp1 + (size_t)s0 * N * N;
      NMGP_TRY(launch_prior_solve(L0, pl->w.R0, pl->w.Z0, cs, N, 1, 0, st, &launches));
      NMGP_TRY(launch_prior_solve(L1, pl->w.R1, pl->w.Z1, cs, N, pl->nv1, 0, st, &launches));
      NMGP_TRY(launch_prior_solve(L0, pl->w.Z0, pl->w.G0, cs, N, 1, 1, st, &launches));
      NMGP_TRY(launch_prior_solve(L1, pl->w.Z1, pl->w.G1, cs, N, pl->nv1, 1, st, &launches));
    }
    NMGP_TRY(hyper_chunk(pl, h, s0, cs, ps, hgrad, st, &launches));
  }
  pl->last_launches = launches;
  return 0;
}

int nmgp_sweep_reduce(const double* vals, const double* hgrad, const int* info, long S, double* out17, void* stream) {
  if (S < 0 || !out17 || (S > 0 && (!vals || !info))) { set_last_error("nmgp_sweep_reduce: bad arguments"); return NMGP_EINVAL; }
  return launch_sweep_reduce(vals, hgrad, info, S, out17, (cudaStream_t)stream);
}

int nmgp_plan_set_hyper(nmgp_plan* pl, const double* hyper, void* stream) {
  if (!pl || !hyper) { set_last_error("nmgp_plan_set_hyper: null argument"); return NMGP_EINVAL; }
  for (int k = 0; k < NMGP_NHYPER; ++k)
    if (!(hyper[k] == hyper[k])) { set_last_error("nmgp_plan_set_hyper: NaN hyper-parameter"); return NMGP_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  const bool refactor0 = pl->nprior && (hyper[1] != pl->hyper[1] || hyper[2] != pl->hyper[2]);
  const bool refactor1 = pl->nprior && (hyper[4] != pl->hyper[4] || hyper[5] != pl->hyper[5]);
  std::memcpy(pl->hyper, hyper, sizeof(pl->hyper));
  pl->hc = make_hyper_const(pl->model, hyper, pl->hc.prior);
  drop_graphs(pl);   // the hyper-parameter constants are kernel arguments baked into the captured graphs
  if (pl->S == 0) return 0;
  // only the covariances whose (alpha, beta) moved are factored again; the means enter through the residuals
  if (refactor0) NMGP_TRY(factor_prior(pl, hyper[1], hyper[2], pl->Wp0, pl->hld0, st));
  if (refactor1) NMGP_TRY(factor_prior(pl, hyper[4], hyper[5], pl->Wp1, pl->hld1, st));
  if (refactor0 || refactor1) pl->traces_ready = false;
  return 0;
}

int nmgp_logpost_grad_hyper(nmgp_plan* pl, const double* pars, double* vals, double* grad, double* hgrad, int* info,
                            void* stream) {
  if (!pl || (pl->S > 0 && (!pars || !vals || !info || !hgrad))) { set_last_error("nmgp_logpost_grad_hyper: null argument"); return NMGP_EINVAL; }
  if (pl->model >= NMGP_HADAMARD) { set_last_error("nmgp_logpost_grad_hyper: not available for the Hadamard objectives"); return NMGP_EINVAL; }
  return evaluate(pl, pars, vals, grad, info, (cudaStream_t)stream, nullptr, nullptr, hgrad);
}

int nmgp_adam_step(double* pars, const double* grad, double* m, double* v, const int* info, const unsigned char* frozen,
                   long S, long P, double lr, double beta1, double beta2, double eps, long step, void* stream) {
  if (!pars || !grad || !m || !v || S < 0 || P <= 0 || step < 1) { set_last_error("nmgp_adam_step: bad arguments"); return NMGP_EINVAL; }
  if (S == 0) return 0;
  const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
  const long total = S * P;
  long blocks = (total + 255) / 256;
  if (blocks > (long)sm_count() * 16) blocks = (long)sm_count() * 16;
  adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(pars, grad, m, v, info, frozen, S, P, beta1, beta2, eps,
                                                             lr / bc1, std::sqrt(bc2));
  NMGP_CUDA_TRY(cudaGetLastError());
  return 0;
}

static int ensure_pred_scratch(nmgp_plan* pl, size_t per_subject, int cs) {
  const size_t cap = (2ull << 30) / sizeof(double);
  size_t want = per_subject * (size_t)cs;
  if (want > cap) want = cap > per_subject ? cap : per_subject;
  if (pl->pred_scratch_doubles >= want) return 0;
  if (pl->pred_scratch) { cudaFree(pl->pred_scratch); pl->pred_scratch = nullptr; pl->pred_scratch_doubles = 0; }
  if (cudaMalloc(&pl->pred_scratch, want * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    set_last_error("prediction scratch: cudaMalloc of " + std::to_string(want * sizeof(double)) + " bytes failed");
    return NMGP_ENOMEM;
  }
  pl->pred_scratch_doubles = want;
  return 0;
}

int nmgp_predict_prior_moments(nmgp_plan* pl, const double* pars, const double* xstar, int G, double* mu_l, double* s2_l,
                               double* mu_uL, double* s2_uL, void* stream) {
  if (!pl || (pl->model != NMGP_SEPARABLE && pl->model != NMGP_NONSEPARABLE) || G < 0 || (pl->S > 0 && G > 0 && (!pars || !xstar || !mu_l || !s2_l || !mu_uL || !s2_uL))) {
    set_last_error("nmgp_predict_prior_moments: needs a separable or nonseparable plan and non-null buffers");
    return NMGP_EINVAL;
  }
  if (pl->S == 0 || G == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int N = pl->N, M = pl->M, P = pl->P, nv1 = pl->nv1;
  long launches = 0;
  const size_t per = predict_prior_scratch_per_subject(N, G);
  NMGP_TRY(ensure_pred_scratch(pl, per, pl->chunk));
  const long SB = (long)(pl->pred_scratch_doubles / per);
  for (int s0 = 0; s0 < pl->S; s0 += pl->chunk) {
    const int cs = pl->S - s0 < pl->chunk ? pl->S - s0 : pl->chunk;
    const double* ps = pars + (size_t)s0 * P;
    if (pl->model == NMGP_NONSEPARABLE) {
      NMGP_TRY(launch_svc_prep(cs, N, M, ps, P, pl->hc, pl->w, st, &launches));
    } else {   // the separable prep also rotates the observations; only its residuals R0 / R1 are used here
      BlockBatch b = pl->bb;
      b.batch = cs * pl->nmat;
      NMGP_TRY(sep_forward(pl->model, cs, N, M, pl->x + (size_t)s0 * N, pl->Y + (size_t)s0 * N * M, ps, P, pl->hc, pl->w, b, st,
                           &launches));
    }
    const double* L0 = pl->Wp0 + (size_t)s0 * N * N;
    const double* L1 = pl->Wp1 + (size_t)s0 * N * N;
    NMGP_TRY(launch_prior_solve(L0, pl->w.R0, pl->w.Z0, cs, N, 1, 0, st, &launches));
    NMGP_TRY(launch_prior_solve(L1, pl->w.R1, pl->w.Z1, cs, N, nv1, 0, st, &launches));
    for (int c0 = 0; c0 < cs; c0 += (int)SB) {
      const int sb = cs - c0 < SB ? cs - c0 : (int)SB;
      const size_t g0 = (size_t)(s0 + c0);
      const double* xs = pl->x + g0 * N;
      const double* xq = xstar + g0 * G;
      NMGP_TRY(launch_predict_prior(xs, L0 + (size_t)c0 * N * N, pl->w.Z0 + (size_t)c0 * N, sb, N, 1, xq, G, pl->hyper[1],
                                    pl->hyper[2], pl->hyper[0], pl->pred_scratch, mu_l + g0 * G, s2_l + g0 * G, st, &launches));
      NMGP_TRY(launch_predict_prior(xs, L1 + (size_t)c0 * N * N, pl->w.Z1 + (size_t)c0 * N * nv1, sb, N, nv1, xq, G,
                                    pl->hyper[4], pl->hyper[5], pl->hyper[3], pl->pred_scratch, mu_uL + g0 * G * nv1,
                                    s2_uL + g0 * G, st, &launches));
    }
  }
  pl->last_launches = launches;
  return 0;
}

int nmgp_predict_moments_sep(nmgp_plan* pl, const double* pars, const double* xstar, int G, int n_sample,
                             const double* tl_star, const double* ts_star, double* mu_f, double* quad, int* info,
                             void* stream) {
  if (!pl || (pl->model != NMGP_STATIONARY && pl->model != NMGP_SEPARABLE) || G < 0 || n_sample < 0 ||
      (pl->S > 0 && G > 0 && n_sample > 0 && (!pars || !xstar || !tl_star || !ts_star || !mu_f || !quad))) {
    set_last_error("nmgp_predict_moments_sep: needs a stationary or separable plan and non-null buffers");
    return NMGP_EINVAL;
  }
  if (pl->S == 0 || G == 0 || n_sample == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int N = pl->N, M = pl->M, P = pl->P;
  const size_t C = (size_t)G * n_sample;
  long launches = 0;
  const size_t per = predict_sep_scratch_per_subject(N, M, (long)C);
  NMGP_TRY(ensure_pred_scratch(pl, per, pl->chunk));
  for (int s0 = 0; s0 < pl->S; s0 += pl->chunk) {
    const int cs = pl->S - s0 < pl->chunk ? pl->S - s0 : pl->chunk;
    BlockBatch b = pl->bb;
    b.batch = cs * pl->nmat;
    const double* xs = pl->x + (size_t)s0 * N;
    NMGP_TRY(sep_forward(pl->model, cs, N, M, xs, pl->Y + (size_t)s0 * N * M, pars + (size_t)s0 * P, P, pl->hc, pl->w, b, st,
                         &launches));
    NMGP_TRY(run_potrf(pl, b, st, &launches));
    NMGP_TRY(run_potri(pl, b, st, &launches));
    if (info) NMGP_TRY(launch_reduce_info(b.info, cs, M, info + s0, st, &launches));
    NMGP_TRY(predict_moments_sep_chunk(cs, N, M, xs, pl->w, b, xstar + (size_t)s0 * G, tl_star + (size_t)s0 * C,
                                       ts_star + (size_t)s0 * C, G, n_sample, pl->pred_scratch, pl->pred_scratch_doubles,
                                       mu_f + (size_t)s0 * C * M, quad + (size_t)s0 * C * M, st, &launches));
  }
  pl->last_launches = launches;
  return 0;
}

int nmgp_predict_moments(nmgp_plan* pl, const double* pars, const double* xstar, int G, int n_sample, const double* tl_star,
                         const double* uL_star, int flags, double* mu_f, double* s2_y, int* info, void* stream) {
  if (!pl || pl->model != NMGP_NONSEPARABLE || G < 0 || n_sample < 0 ||
      (pl->S > 0 && G > 0 && n_sample > 0 && (!pars || !xstar || !tl_star || !uL_star || !mu_f || !s2_y))) {
    set_last_error("nmgp_predict_moments: needs a nonseparable plan and non-null buffers");
    return NMGP_EINVAL;
  }
  if (pl->S == 0 || G == 0 || n_sample == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int N = pl->N, M = pl->M, P = pl->P, T = pl->T;
  const size_t C = (size_t)G * n_sample;
  long launches = 0;
  const size_t per = predict_scratch_per_subject(N, M, (long)C);
  NMGP_TRY(ensure_pred_scratch(pl, per, pl->chunk));
  for (int s0 = 0; s0 < pl->S; s0 += pl->chunk) {
    const int cs = pl->S - s0 < pl->chunk ? pl->S - s0 : pl->chunk;
    BlockBatch b = pl->bb;
    b.batch = cs;
    const double* xs = pl->x + (size_t)s0 * N;
    const double* Ys = pl->Y + (size_t)s0 * N * M;
    NMGP_TRY(svc_forward(cs, N, M, xs, pars + (size_t)s0 * P, P, pl->hc, pl->w, b, st, &launches));
    NMGP_TRY(run_potrf(pl, b, st, &launches));
    NMGP_TRY(run_potri(pl, b, st, &launches));
    if (info) NMGP_CUDA_TRY(cudaMemcpyAsync(info + s0, b.info, (size_t)cs * sizeof(int), cudaMemcpyDeviceToDevice, st));
    NMGP_TRY(predict_moments_chunk(cs, N, M, xs, Ys, pl->w, b, xstar + (size_t)s0 * G, tl_star + (size_t)s0 * C,
                                   uL_star + (size_t)s0 * C * T, G, n_sample, (flags & NMGP_PRED_RAW_FACTOR) ? 1 : 0,
                                   pl->pred_scratch, pl->pred_scratch_doubles,
                                   mu_f + (size_t)s0 * C * M, s2_y + (size_t)s0 * C * M, st, &launches));
  }
  pl->last_launches = launches;
  return 0;
}

int nmgp_hmc_kick(double* p, const double* grad, const int* info, long S, long P, double step, void* stream) {
  if (!p || !grad || S < 0 || P <= 0) { set_last_error("nmgp_hmc_kick: bad arguments"); return NMGP_EINVAL; }
  if (S == 0) return 0;
  long blocks = (S * P + 255) / 256;
  if (blocks > (long)sm_count() * 16) blocks = (long)sm_count() * 16;
  hmc_kick_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, grad, info, S, P, step);
  NMGP_CUDA_TRY(cudaGetLastError());
  return 0;
}

int nmgp_hmc_drift(double* q, const double* p, long S, long P, double eps, void* stream) {
  if (!q || !p || S < 0 || P <= 0) { set_last_error("nmgp_hmc_drift: bad arguments"); return NMGP_EINVAL; }
  if (S == 0) return 0;
  long blocks = (S * P + 255) / 256;
  if (blocks > (long)sm_count() * 16) blocks = (long)sm_count() * 16;
  hmc_drift_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(q, p, S * P, eps);
  NMGP_CUDA_TRY(cudaGetLastError());
  return 0;
}

int nmgp_hmc_accept(double* q, const double* q_prop, double* grad, const double* grad_prop, double* U, const double* vals_prop,
                    const double* p0, const double* p1, const int* failed, const double* log_u, int* accepted, long S,
                    long P, void* stream) {
  if (!q || !q_prop || !grad || !grad_prop || !U || !vals_prop || !p0 || !p1 || !log_u || !accepted || S < 0 || P <= 0) {
    set_last_error("nmgp_hmc_accept: bad arguments");
    return NMGP_EINVAL;
  }
  if (S == 0) return 0;
  hmc_accept_kernel<<<(unsigned)S, 256, 0, (cudaStream_t)stream>>>(q, q_prop, grad, grad_prop, U, vals_prop, p0, p1, failed,
                                                                   log_u, accepted, P);
  NMGP_CUDA_TRY(cudaGetLastError());
  return 0;
}

int nmgp_plan_set_engine(nmgp_plan* pl, int mode) {
  if (!pl || mode < 0 || mode > 4) return NMGP_EINVAL;
  drop_graphs(pl);                          // the captured kernel sequence belongs to the old engine choice
  pl->engine_mode = mode == 3 ? 2 : mode;   // 4: automatic potrf, level-synchronous recursive inverse
  pl->stable_inverse = mode == 3;
  return 0;
}

int nmgp_plan_set_graph(nmgp_plan* pl, int mode) {
  if (!pl || mode < 0 || mode > 1) { set_last_error("nmgp_plan_set_graph: mode must be 0 (auto) or 1 (never)"); return NMGP_EINVAL; }
  pl->graph_mode = mode;
  return 0;
}
long nmgp_plan_graph_replays(const nmgp_plan* pl) { return pl ? pl->graph_replays : 0; }
long nmgp_plan_last_launches(const nmgp_plan* pl) { return pl ? pl->last_launches : 0; }
size_t nmgp_plan_device_bytes(const nmgp_plan* pl) { return pl ? pl->dev_bytes : 0; }
int nmgp_plan_chunk(const nmgp_plan* pl) { return pl ? pl->chunk : 0; }
int nmgp_plan_block(const nmgp_plan* pl) { return pl ? pl->bb.NB : 0; }

int nmgp_rbf_cov(const double* x1, int N1, const double* x2, int N2, double alpha, double beta, double* out, void* stream) {
  if (!x1 || !out || N1 < 0 || (x2 && N2 < 0)) { set_last_error("nmgp_rbf_cov: bad arguments"); return NMGP_EINVAL; }
  return launch_rbf_cov(x1, N1, x2, N2, alpha, beta, out, (cudaStream_t)stream);
}

int nmgp_gibbs_cov(const double* x1, const double* sigma1, const double* ell1, int N1, const double* x2,
                   const double* sigma2, const double* ell2, int N2, double* out, void* stream) {
  if (!x1 || !ell1 || !out || N1 < 0 || (x2 && (!ell2 || N2 < 0))) { set_last_error("nmgp_gibbs_cov: bad arguments"); return NMGP_EINVAL; }
  return launch_gibbs_cov(x1, sigma1, ell1, N1, x2, sigma2, ell2, N2, out, (cudaStream_t)stream);
}

int nmgp_nonseparable_cov(const double* x, const double* pars, int batch, int N, int M, double* out, void* stream) {
  if (!x || !pars || !out || batch < 0 || N <= 0 || M <= 0 || M > 16) { set_last_error("nmgp_nonseparable_cov: bad arguments"); return NMGP_EINVAL; }
  return launch_nonseparable_cov_reference_order(x, pars, batch, N, M, out, (cudaStream_t)stream);
}

int nmgp_pairwise_sqdist(const double* x1, int N1, const double* x2, int N2, double* out, void* stream) {
  if (!x1 || !out || N1 < 0 || (x2 && N2 < 0)) { set_last_error("nmgp_pairwise_sqdist: bad arguments"); return NMGP_EINVAL; }
  return launch_pairwise_sqdist(x1, N1, x2, N2, out, (cudaStream_t)stream);
}

int nmgp_kron(const double* t1, int h1, int w1, const double* t2, int h2, int w2, double* out, void* stream) {
  if (!t1 || !t2 || !out || h1 < 0 || w1 < 0 || h2 < 0 || w2 < 0) { set_last_error("nmgp_kron: bad arguments"); return NMGP_EINVAL; }
  return launch_kron(t1, h1, w1, t2, h2, w2, out, (cudaStream_t)stream);
}

int nmgp_kron_mv(const double* B, int M1, int M2, const double* K, int N1, int N2, const double* y, double* out,
                 double* scratch, void* stream) {
  if (!B || !K || !y || !out || !scratch || M1 < 0 || M2 < 0 || N1 < 0 || N2 < 0) { set_last_error("nmgp_kron_mv: bad arguments"); return NMGP_EINVAL; }
  return launch_kron_mv(B, M1, M2, K, N1, N2, y, out, scratch, (cudaStream_t)stream);
}

int nmgp_gram(const double* L, int R, int M, double* out, void* stream) {
  if (!L || !out || R < 0 || M < 0) { set_last_error("nmgp_gram: bad arguments"); return NMGP_EINVAL; }
  return launch_gram(L, R, M, out, (cudaStream_t)stream);
}

int nmgp_sym_eig(const double* B, int M, double* lam, double* V, void* stream) {
  if (!B || !lam || !V) { set_last_error("nmgp_sym_eig: null argument"); return NMGP_EINVAL; }
  return launch_sym_eig(B, M, lam, V, (cudaStream_t)stream);
}

int nmgp_kron_eig_solve(const double* B, int M, const double* K, int N, double sigma2, double* inv_out, const double* r,
                        double* out2, int* info, void* stream) {
  if (!B || !K || (!inv_out && !out2)) { set_last_error("nmgp_kron_eig_solve: null argument"); return NMGP_EINVAL; }
  return kron_eig_solve(B, M, K, N, sigma2, inv_out, r, out2, info, (cudaStream_t)stream);
}

int nmgp_potrf_batched(double* A, int n, int batch, double* logdet, int* info, void* stream) {
  return potrf_unit(A, n, batch, logdet, info, 0, (cudaStream_t)stream);
}

int nmgp_potrf_potri_batched(double* A, int n, int batch, double* logdet, int* info, void* stream) {
  return potrf_unit(A, n, batch, logdet, info, 1, (cudaStream_t)stream);
}

}  // extern "C"
