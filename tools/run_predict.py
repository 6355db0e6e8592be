"""Time the posterior-prediction path on the GPU box (CUDA events).
usage: python tools/run_predict.py N M S G NS [reps]
Prints the milliseconds of nmgp_predict_prior_moments and nmgp_predict_moments for S subjects, G new inputs and NS samples
each (the driver's shape is G = 201, NS = 100: Nonseparable_model.py:377), and the DMMA rate of the quadratic-form kernel
(T * 2 * N^2 * G * NS flop per subject)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nonstationary_multivariate_gaussian_process_b200 import synth
from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
from bench import HYPER

N, M, S, G, NS = (int(v) for v in sys.argv[1:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
T = M * (M + 1) // 2
rng = np.random.RandomState(0)
xs, ps = [], []
for s in range(S):
    x, _, _, _ = synth.truth(N, M, s)
    xs.append(x); ps.append(synth.start_point("nonseparable", N, M, s, 0.02))
Y = rng.standard_normal((S, N, M))
plan = LogPosteriorPlan("nonseparable", np.stack(xs), Y, HYPER["nonseparable"])
p = torch.from_numpy(np.stack(ps)).cuda()
grids = torch.linspace(0, 1, G, dtype=torch.float64).cuda()
tl = torch.from_numpy(-3.0 + 0.3 * rng.standard_normal((S, G, NS))).cuda()
ul = torch.from_numpy(0.3 * rng.standard_normal((S, G, NS, T))).cuda()


def timed(fn):
    """median of per-call CUDA-event times after warm-up (the first call allocates the scratch and loads the kernels)"""
    for _ in range(3):
        fn()
    ts = []
    for _ in range(max(reps, 5)):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


ms_prior, _ = timed(lambda: plan.predict_prior_moments(p, grids))
l_prior = plan.last_launches
ms_mom, out = timed(lambda: plan.predict_moments(p, grids, tl, ul))
l_mom = plan.last_launches
ms_eval, _ = timed(lambda: plan.value_and_grad(p))
flop = S * T * 2.0 * N * N * G * NS
print(json.dumps({"N": N, "M": M, "S": S, "G": G, "n_sample": NS, "ms_prior_moments": ms_prior, "ms_predict_moments": ms_mom,
                  "ms_value_and_grad": ms_eval, "launches": [l_prior, l_mom],
                  "quad_form_tflops_lower_bound": flop / (ms_mom * 1e-3) / 1e12, "info_bad": int((out[2] != 0).sum()),
                  "finite": bool(torch.isfinite(out[0]).all() and torch.isfinite(out[1]).all())}))
