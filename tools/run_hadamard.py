"""Time the irregularly sampled ("Hadamard") objectives on the GPU box.  usage: python tools/run_hadamard.py MODEL N M S [reps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nonstationary_multivariate_gaussian_process_b200 import synth
from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan

model, N, M, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
hyper = {"hadamard": {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_tilde_sigma": 0.0, "alpha_tilde_sigma": 1.0,
                      "beta_tilde_sigma": 1.0, "a": 1e-2, "b": 1e-2, "c": 0.1},
         "hadamard_svc": {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0,
                          "a": 1e-2, "b": 1e-2},
         "hadamard_s": {"mu_tilde_l": 0.0, "sigma_tilde_l": 10.0, "a": 1e-6, "b": 1e-6, "c": 1.0}}[model]
base = [synth.hadamard_case(model, N, M, s) for s in range(min(S, 8))]
xs, ixs, ys, ps = (np.stack([base[s % len(base)][k] for s in range(S)]) for k in range(4))
plan = LogPosteriorPlan(model, xs, ys, hyper, indx=ixs, M=M)
p = torch.from_numpy(ps).cuda()
for _ in range(3):
    plan.value_and_grad(p)
ts = []
for _ in range(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); v, g, i = plan.value_and_grad(p); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
ph, _, _, _ = plan.profile(p)
print(json.dumps({"model": model, "N": N, "M": M, "S": S, "ms_per_eval_batch": ms, "evals_per_s": S / ms * 1e3, "launches": plan.last_launches,
                  "phase_ms": ph, "info_bad": int((i != 0).sum())}))
