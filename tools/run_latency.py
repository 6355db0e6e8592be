"""Single-plan evaluation latency with and without CUDA-graph replay.  usage: python tools/run_latency.py MODEL N M S [reps]
device = nmgp_logpost_grad into fixed buffers (what map_fit / hmc_sample do), host = nmgp_logpost_grad_host (what the
reference-signature shim calls), shim = logpos.nlogpos_obj*(pars, Y, x, **hyper) + .backward() on a CPU leaf tensor."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nonstationary_multivariate_gaussian_process_b200 import _lib, logpos, synth
from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan

model, N, M, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 200
hyper = {"nonseparable": {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0,
                          "a": 1e-2, "b": 1e-2},
         "separable": {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_tilde_sigma": 0.0, "alpha_tilde_sigma": 1.0,
                       "beta_tilde_sigma": 1.0, "a": 1e-2, "b": 1e-2, "c": 0.1},
         "stationary": {"mu_tilde_l": 0.0, "sigma_tilde_l": 10.0, "a": 1e-6, "b": 1e-6, "c": 1.0}}[model]
base = [synth.sample_subject(N, M, s)[:2] + (synth.start_point(model, N, M, s, 0.02),) for s in range(S)]
xs, Ys, ps = (np.stack([b[k] for b in base]) for k in range(3))
out = {"model": model, "N": N, "M": M, "S": S, "reps": reps}
for graph in (False, True):
    plan = LogPosteriorPlan(model, xs, Ys, hyper)
    plan.set_graph(graph)
    p = torch.from_numpy(ps).cuda()
    buf = (torch.empty((S, _lib.NVALS), dtype=torch.float64, device="cuda"), torch.empty_like(p),
           torch.empty((S,), dtype=torch.int32, device="cuda"))
    for _ in range(5):
        plan.value_and_grad(p, out=buf)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        plan.value_and_grad(p, out=buf)
    torch.cuda.synchronize(); dev = (time.perf_counter() - t0) / reps * 1e3
    ph = torch.from_numpy(ps)
    for _ in range(5):
        plan.value_and_grad_host(ph)
    t0 = time.perf_counter()
    for _ in range(reps):
        plan.value_and_grad_host(ph)
    host = (time.perf_counter() - t0) / reps * 1e3
    tag = "graph" if graph else "direct"
    out[f"device_ms_{tag}"] = dev
    out[f"host_ms_{tag}"] = host
    out[f"replays_{tag}"] = plan.graph_replays
    out["launches"] = plan.last_launches
if S == 1:
    fn = {"stationary": logpos.nlogpos_obj_S, "separable": logpos.nlogpos_obj, "nonseparable": logpos.nlogpos_obj_SVC}[model]
    Yt, xt = torch.from_numpy(Ys[0]), torch.from_numpy(xs[0])
    leaf = torch.from_numpy(ps[0]).clone().requires_grad_(True)
    def call():
        leaf.grad = None
        fn(leaf, Yt, xt, **hyper).backward()
    for _ in range(5):
        call()
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    out["shim_value_and_backward_ms"] = (time.perf_counter() - t0) / reps * 1e3
print(json.dumps(out))
