"""Run one configuration through the plan API and print per-phase timings (GPU box).
usage: python tools/run_config.py MODEL N M S [reps] [auto|left|right|left_stable]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nonstationary_multivariate_gaussian_process_b200 import synth
from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
from bench import HYPER, algorithmic_flops

model, N, M, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
engine = sys.argv[6] if len(sys.argv) > 6 else "auto"
rng = np.random.RandomState(0)
xs, ps = [], []
for s in range(S):
    x, _, _, _ = synth.truth(N, M, s)
    xs.append(x); ps.append(synth.start_point(model, N, M, s, 0.02))
Y = rng.standard_normal((S, N, M))
t0 = time.time()
plan = LogPosteriorPlan(model, np.stack(xs), Y, HYPER[model])
torch.cuda.synchronize()
print(f"plan: {time.time()-t0:.2f}s chunk={plan.chunk} dev_bytes={plan.device_bytes/2**30:.2f} GiB")
plan.set_engine(engine)
p = torch.from_numpy(np.stack(ps)).cuda()
for _ in range(2):
    plan.value_and_grad(p)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    v, g, i = plan.value_and_grad(p)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
ph, _, _, _ = plan.profile(p)
f1, f2 = algorithmic_flops(model, N, M)
print(json.dumps({"model": model, "engine": engine, "N": N, "M": M, "S": S, "ms_per_eval_batch": ms, "evals_per_s": S / ms * 1e3,
                  "launches": plan.last_launches, "phase_ms": ph,
                  "potrf_tflops": S * f1 / ph["potrf"] / 1e9, "potri_tflops": S * f2 / ph["potri"] / 1e9,
                  "info_bad": int((i != 0).sum()), "neglogpost0": float(v[0, 0])}))
