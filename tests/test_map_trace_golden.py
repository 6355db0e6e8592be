"""The drivers' MAP loops end to end against traces recorded with the UNMODIFIED reference (tests/golden/make_golden_map.py):
C1 of BASELINE.json -- Stationary_model.py:106-131, 1000 Adam steps, lr 0.1, M=2, N=50 -- and a 300-step separable loop.
CPU: the oracle reproduces the traces.  GPU: the reference-signature shim (autograd on CPU leaves, exactly the drivers' loop)
and the device-resident `map_fit` reproduce them."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

CASES = ["map_stationary_N50_M2_s0", "map_separable_N40_M3_s1"]
# (trace tolerance, final-parameter tolerance): the separable objective carries the GP priors' conditioning floor
TOL = {"stationary": (1e-9, 1e-7), "separable": (1e-6, 1e-4)}


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["model"], d["hyper"] = str(d["model"]), json.loads(str(d["hyper"]))
    for k in ("N", "M", "steps"):
        d[k] = int(d[k])
    d["lr"] = float(d["lr"])
    return d


def leaves_of(d):
    p0, N, M = d["pars0"], d["N"], d["M"]
    T = M * (M + 1) // 2
    if d["model"] == "stationary":      # tilde_sigma fixed at 0 and not optimised (Stationary_model.py:88,116)
        return [torch.tensor(p0[0:1], requires_grad=True), torch.zeros(1, dtype=torch.float64),
                torch.tensor(p0[2:2 + T], requires_grad=True), torch.tensor(p0[-1:], requires_grad=True)]
    return [torch.tensor(p0[:N], requires_grad=True), torch.tensor(p0[N:2 * N], requires_grad=True),
            torch.tensor(p0[2 * N:2 * N + T], requires_grad=True), torch.tensor(p0[-1:], requires_grad=True)]


def adam_loop(objective, d):
    leaves = leaves_of(d)
    Y, x = torch.from_numpy(d["Y"]), torch.from_numpy(d["x"])
    opt = torch.optim.Adam([l for l in leaves if l.requires_grad], lr=d["lr"])
    trace = np.zeros((d["steps"], 2))
    for i in range(d["steps"]):
        opt.zero_grad()
        out = objective(torch.cat([l.reshape(-1) for l in leaves]), Y, x, verbose=True, **d["hyper"])
        out[0].backward()
        opt.step()
        trace[i] = [float(out[0]), float(out[1])]
    return trace, torch.cat([l.detach().reshape(-1) for l in leaves]).numpy()


def check(d, trace, pf, what):
    tt, tp = TOL[d["model"]]
    et = np.max(np.abs(trace - d["trace"]) / np.maximum(np.abs(d["trace"]), 1.0))
    ep = np.max(np.abs(pf - d["pars_final"]))
    assert et < tt and ep < tp, (what, et, ep)
    assert trace[-1, 0] < trace[0, 0]


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_reference_map_trace(name):
    from oracle import nmgp_oracle as O
    d = load(name)

    def objective(pars, Y, x, verbose=True, **hyper):
        out = O._MODELS[d["model"]](pars, Y, x, **hyper)
        return (-out[0],) + tuple(o.detach() for o in out[1:])
    trace, pf = adam_loop(objective, d)
    check(d, trace, pf, "oracle")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_shim_reproduces_the_reference_map_trace(name, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import logpos
    d = load(name)
    fn = {"stationary": logpos.nlogpos_obj_S, "separable": logpos.nlogpos_obj}[d["model"]]
    trace, pf = adam_loop(fn, d)
    check(d, trace, pf, "shim")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_resident_map_fit_reproduces_the_reference_map_trace(name, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    d = load(name)
    plan = LogPosteriorPlan(d["model"], d["x"], d["Y"], d["hyper"])
    frozen = np.zeros(plan.P, dtype=bool)
    p0 = d["pars0"].copy()
    if d["model"] == "stationary":
        frozen[1] = True
        p0[1] = 0.0
    pf, tr, info = plan.map_fit(p0[None, :], steps=d["steps"], lr=d["lr"], frozen=frozen)
    assert int(info.abs().sum()) == 0
    check(d, tr[:, 0, :2].cpu().numpy(), pf[0].cpu().numpy(), "map_fit")
    assert plan.graph_replays >= d["steps"] - 2          # every iteration after the first two is one graph replay
