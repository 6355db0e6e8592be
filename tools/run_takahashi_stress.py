"""NMGP_TAKAHASHI_GUARD=0 python tools/run_takahashi_stress.py  -> the RAW sweep;  without the variable: the guarded one.
Accuracy of the two inverse formulations on ill-conditioned covariances (GPU box): Takahashi sweep ('left') vs the
backward-stable W^T W inverse ('left_stable'), both against the CPU oracle (the reference's LU inverse + autograd), for a
grid of matrix sizes (block columns Kt) and noise variances.  Evidence for api.cu:run_potri's choice and for
tests/test_gpu_edge_cases.py::test_takahashi_sweep_*.   usage: python tools/run_takahashi_stress.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import rel_err
from nonstationary_multivariate_gaussian_process_b200 import synth
from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
from oracle import nmgp_oracle as O

HYPER = {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0, "a": 1e-2, "b": 1e-2}
print(f"{'N':>4s} {'M':>2s} {'n':>5s} {'Kt':>3s} {'log_s2':>7s} | {'val left':>9s} {'val stab':>9s} | {'grad left':>9s} {'grad stab':>9s} | left-vs-stable grad")
for N, M in [(100, 6), (96, 8), (128, 8)]:
    for log_s2 in (-4.0, -6.0, -8.0, -10.0, -11.5):
        x, Y, _ = synth.sample_subject(N, M, 40)
        p = synth.start_point("nonseparable", N, M, 40, 0.02)
        p[-1] = log_s2
        ov, og = O.value_and_grad("nonseparable", p, Y, x, Prior=False, **HYPER)
        res = {}
        for engine in ("left", "left_stable"):
            plan = LogPosteriorPlan("nonseparable", x, Y, HYPER, prior=False)
            plan.set_engine(engine)
            vals, grad, info = plan.value_and_grad_host(torch.from_numpy(p))
            plan.close()
            res[engine] = (float(vals[0, 0]), grad.numpy()[0], int(info[0]))
        n = N * M
        print(f"{N:4d} {M:2d} {n:5d} {(n + 63) // 64:3d} {log_s2:7.1f} | {rel_err(res['left'][0], float(ov[0])):9.1e} "
              f"{rel_err(res['left_stable'][0], float(ov[0])):9.1e} | {rel_err(res['left'][1], og.numpy()):9.1e} "
              f"{rel_err(res['left_stable'][1], og.numpy()):9.1e} | {rel_err(res['left'][1], res['left_stable'][1]):9.1e}"
              f"  info={res['left'][2]},{res['left_stable'][2]}", flush=True)
