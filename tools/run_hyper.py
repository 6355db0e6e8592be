"""Time the hyper-parameter gradient on the GPU box.  usage: python tools/run_hyper.py MODEL N M S [reps]
Reports: first call (forms the cached prior traces), nmgp_hyper_grad alone, the plain evaluation, the fused evaluation
(nmgp_logpost_grad_hyper) and nmgp_plan_set_hyper (re-factoring both prior covariances)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nonstationary_multivariate_gaussian_process_b200 import synth
from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan

model, N, M, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
hyper = {"nonseparable": {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0,
                          "a": 1e-2, "b": 1e-2},
         "separable": {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_tilde_sigma": 0.0, "alpha_tilde_sigma": 1.0,
                       "beta_tilde_sigma": 1.0, "a": 1e-2, "b": 1e-2, "c": 0.1},
         "stationary": {"mu_tilde_l": 0.0, "sigma_tilde_l": 10.0, "a": 1e-6, "b": 1e-6, "c": 1.0}}[model]
base = [synth.sample_subject(N, M, s)[:2] + (synth.start_point(model, N, M, s, 0.02),) for s in range(min(S, 8))]
xs, Ys, ps = (np.stack([base[s % len(base)][k] for s in range(S)]) for k in range(3))
plan = LogPosteriorPlan(model, xs, Ys, hyper)
p = torch.from_numpy(ps).cuda()


def timed(fn):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


torch.cuda.synchronize(); t0 = time.perf_counter(); plan.hyper_grad(p); torch.cuda.synchronize()
first = (time.perf_counter() - t0) * 1e3
for _ in range(2):
    plan.value_and_grad(p); plan.value_grad_and_hyper_grad(p)
out = {"model": model, "N": N, "M": M, "S": S, "first_hyper_grad_ms": first,
       "hyper_grad_ms": timed(lambda: plan.hyper_grad(p)), "hyper_grad_launches": plan.last_launches,
       "eval_ms": timed(lambda: plan.value_and_grad(p)), "eval_launches": plan.last_launches,
       "fused_ms": timed(lambda: plan.value_grad_and_hyper_grad(p)), "fused_launches": plan.last_launches}
h2 = dict(hyper)
if model != "stationary":
    h2["beta_tilde_l"] = 0.9
    h2["beta_L" if model == "nonseparable" else "beta_tilde_sigma"] = 0.9
else:
    h2["sigma_tilde_l"] = 5.0
flip = [h2, hyper]
out["set_hyper_ms"] = timed(lambda: plan.set_hyper(flip[0]) or flip.reverse())
print(json.dumps(out))
