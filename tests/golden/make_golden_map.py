"""Golden MAP traces: the drivers' Adam loops run with the UNMODIFIED reference objective (build container only).

    python tests/golden/make_golden_map.py

C1 of BASELINE.json -- Stationary_Model/Stationary_model.py:106-131: 1000 Adam steps (lr 0.1) on [tilde_l, uL_vec,
tilde_sigma2_err] with tilde_sigma fixed at 0, hyper-parameters of :79, start point of the non-empirical branch (:99-101,
tilde_l = -3, tilde_sigma2_err = log 0.1; uL_vec from the seeded synthetic start instead of torch.rand), on SIM_code/sim.py
style data (M=2, N=50).  Recorded: -log posterior and log-likelihood of every step, the final parameters (MAP.dat).
Also a 300-step separable loop (Separable_model.py:149-231 shape of the loop, lr 0.01) at N=40, M=3.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")
torch.symeig = lambda A, eigenvectors=False, upper=True: torch.linalg.eigh(A, UPLO="U" if upper else "L")
from Utility import logpos  # noqa: E402  (the reference)

from nonstationary_multivariate_gaussian_process_b200 import synth  # noqa: E402


def run(objective, leaves, fixed_builder, Y, x, hyper, steps, lr):
    opt = torch.optim.Adam([l for l in leaves if l.requires_grad], lr=lr)
    trace = np.zeros((steps, 2))
    for i in range(steps):
        opt.zero_grad()
        pars = fixed_builder(leaves)
        out = objective(pars, Y, x, verbose=True, **hyper)
        out[0].backward()
        opt.step()
        trace[i] = [float(out[0]), float(out[1])]
    return trace, fixed_builder(leaves).detach().numpy()


def main():
    torch.set_num_threads(8)
    # ---- C1
    N, M, seed = 50, 2, 0
    x, Y, _ = synth.sample_subject(N, M, seed)
    p0 = synth.start_point("stationary", N, M, seed, 0.0)
    p0[0], p0[1], p0[-1] = -3.0, 0.0, float(np.log(np.float32(0.1)))
    hyper = {"mu_tilde_l": 0, "sigma_tilde_l": 1., "a": 1., "b": 1., "c": 10.}
    leaves = [torch.tensor(p0[0:1], requires_grad=True), torch.zeros(1, dtype=torch.float64),
              torch.tensor(p0[2:5], requires_grad=True), torch.tensor(p0[-1:], requires_grad=True)]
    trace, pf = run(logpos.nlogpos_obj_S, leaves, lambda ls: torch.cat([l.reshape(-1) for l in ls]), torch.from_numpy(Y),
                    torch.from_numpy(x), hyper, 1000, 0.1)
    np.savez_compressed(os.path.join(HERE, "map_stationary_N50_M2_s0.npz"), model="stationary", N=N, M=M, x=x, Y=Y, pars0=p0,
                        hyper=json.dumps(hyper), steps=1000, lr=0.1, trace=trace, pars_final=pf,
                        torch_version=torch.__version__, threads=torch.get_num_threads())
    print("stationary", trace[0], trace[-1], pf)
    # ---- separable
    N, M, seed = 40, 3, 1
    x, Y, _ = synth.sample_subject(N, M, seed)
    p0 = synth.start_point("separable", N, M, seed, 0.0)
    hyper = {"mu_tilde_l": 0., "alpha_tilde_l": 10., "beta_tilde_l": 1., "mu_tilde_sigma": 0., "alpha_tilde_sigma": 1.,
             "beta_tilde_sigma": 1., "a": 1e-2, "b": 1e-2, "c": 0.1}                # Separable_model_mpisim.py:296-297
    T = M * (M + 1) // 2
    leaves = [torch.tensor(p0[:N], requires_grad=True), torch.tensor(p0[N:2 * N], requires_grad=True),
              torch.tensor(p0[2 * N:2 * N + T], requires_grad=True), torch.tensor(p0[-1:], requires_grad=True)]
    trace, pf = run(logpos.nlogpos_obj, leaves, lambda ls: torch.cat([l.reshape(-1) for l in ls]), torch.from_numpy(Y),
                    torch.from_numpy(x), hyper, 300, 0.01)
    np.savez_compressed(os.path.join(HERE, "map_separable_N40_M3_s1.npz"), model="separable", N=N, M=M, x=x, Y=Y, pars0=p0,
                        hyper=json.dumps(hyper), steps=300, lr=0.01, trace=trace, pars_final=pf,
                        torch_version=torch.__version__, threads=torch.get_num_threads())
    print("separable", trace[0], trace[-1])


if __name__ == "__main__":
    main()
