"""Print the per-component parity of the CUDA path against every golden vector (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import golden_cases, load_golden, rel_err
from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan

print(f"{'case':42s} {'total':>9s} {'loglik':>9s} {'priors(max)':>11s} {'grad(2-norm)':>12s} {'grad(max abs/|g|inf)':>20s}")
for name in golden_cases():
    g = load_golden(name)
    plan = LogPosteriorPlan(g["model"], g["x"], g["Y"], g["hyper"], prior=g["prior"])
    vals, grad, info = plan.value_and_grad_host(torch.from_numpy(g["pars"]))
    vals, grad = vals.numpy()[0], grad.numpy()[0]
    ref = g["vals"]
    e0 = rel_err(vals[0], ref[0])
    e1 = rel_err(vals[1], ref[1]) if len(ref) > 1 else float('nan')
    ep = max([rel_err(vals[k], ref[k]) for k in range(2, len(ref))], default=float('nan'))
    eg = rel_err(grad, g["grad"])
    em = np.abs(grad - g["grad"]).max() / np.abs(g["grad"]).max()
    print(f"{name:42s} {e0:9.2e} {e1:9.2e} {ep:11.2e} {eg:12.2e} {em:20.2e}  info={int(info[0])}")
