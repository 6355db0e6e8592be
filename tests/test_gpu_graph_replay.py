"""GPU: CUDA-graph replay of launch-bound evaluations (nmgp_plan_set_graph) -- the same kernels captured once and replayed:
results must equal the directly launched evaluation bit for bit, for the device path with fixed buffers, for the host-buffer
call, across hyper-parameter changes, and for a subject whose factorisation fails."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HYPER = {
    "stationary": {"mu_tilde_l": 0.0, "sigma_tilde_l": 1.0, "a": 1.0, "b": 1.0, "c": 10.0},
    "separable": {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_tilde_sigma": 0.0,
                  "alpha_tilde_sigma": 1.0, "beta_tilde_sigma": 1.0, "a": 1e-2, "b": 1e-2, "c": 0.1},
    "nonseparable": {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0,
                     "beta_L": 1.0, "a": 1e-2, "b": 1e-2},
}
# (model, N, M, S): C1 and C2 of BASELINE.json, the drivers' one-subject nonseparable shape (n = 600: look-ahead potrf with
# its helper stream inside the capture), a small batch
SHAPES = [("stationary", 50, 2, 1), ("separable", 200, 5, 1), ("nonseparable", 100, 6, 1), ("nonseparable", 40, 3, 4)]


def _case(model, N, M, S, reps):
    from nonstationary_multivariate_gaussian_process_b200 import synth
    xs, Ys = zip(*[synth.sample_subject(N, M, s)[:2] for s in range(S)])
    pars = [np.stack([synth.start_point(model, N, M, s, 0.02 + 0.01 * r) for s in range(S)]) for r in range(reps)]
    return np.stack(xs), np.stack(Ys), pars


@pytest.mark.parametrize("model,N,M,S", SHAPES)
def test_device_path_replay_is_bit_identical(model, N, M, S, cuda_device):
    import torch
    from nonstationary_multivariate_gaussian_process_b200 import _lib
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    x, Y, pars = _case(model, N, M, S, 6)
    direct = LogPosteriorPlan(model, x, Y, HYPER[model])
    direct.set_graph(False)
    plan = LogPosteriorPlan(model, x, Y, HYPER[model])
    p = torch.empty((S, plan.P), dtype=torch.float64, device=cuda_device)
    buf = (torch.empty((S, _lib.NVALS), dtype=torch.float64, device=cuda_device), torch.empty_like(p),
           torch.empty((S,), dtype=torch.int32, device=cuda_device))
    for need_grad in (True, False):
        for r, pr in enumerate(pars):
            p.copy_(torch.from_numpy(pr))
            v, g, i = plan.value_and_grad(p, need_grad=need_grad, out=buf)
            v0, g0, i0 = direct.value_and_grad(p, need_grad=need_grad)
            assert torch.equal(v, v0) and torch.equal(i, i0) and int(i.abs().sum()) == 0, (model, r)
            if need_grad:
                assert torch.equal(g, g0), (model, r)
    assert plan.graph_replays >= 8 and direct.graph_replays == 0     # first call of each variant is direct, second captures
    assert plan.last_launches == direct.last_launches


@pytest.mark.parametrize("model,N,M,S", SHAPES[:3])
def test_host_path_replay_is_bit_identical(model, N, M, S, cuda_device):
    import torch
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    x, Y, pars = _case(model, N, M, S, 5)
    direct = LogPosteriorPlan(model, x, Y, HYPER[model])
    direct.set_graph(False)
    plan = LogPosteriorPlan(model, x, Y, HYPER[model])
    for pr in pars:
        v, g, i = plan.value_and_grad_host(torch.from_numpy(pr))
        v0, g0, i0 = direct.value_and_grad_host(torch.from_numpy(pr))
        assert torch.equal(v, v0) and torch.equal(g, g0) and torch.equal(i, i0)
    assert plan.graph_replays >= 3


def test_moving_buffers_fall_back_to_direct_launches(cuda_device):
    """A caller whose output buffers change on every call gets a few re-captures, then plain launches; results unchanged."""
    import torch
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    x, Y, pars = _case("separable", 30, 3, 2, 1)
    plan = LogPosteriorPlan("separable", x, Y, HYPER["separable"])
    p = torch.from_numpy(pars[0]).to(cuda_device)
    keep = [plan.value_and_grad(p) for _ in range(12)]       # all outputs alive: every call gets new pointers
    for v, g, i in keep[1:]:
        assert torch.equal(v, keep[0][0]) and torch.equal(g, keep[0][1])
    assert plan.graph_replays <= 5


def test_replay_reports_a_failed_subject_and_recovers(cuda_device):
    import torch
    from nonstationary_multivariate_gaussian_process_b200 import _lib
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    x, Y, pars = _case("nonseparable", 40, 3, 4, 1)
    plan = LogPosteriorPlan("nonseparable", x, Y, HYPER["nonseparable"])
    p = torch.from_numpy(pars[0]).to(cuda_device)
    buf = (torch.empty((4, _lib.NVALS), dtype=torch.float64, device=cuda_device), torch.empty_like(p),
           torch.empty((4,), dtype=torch.int32, device=cuda_device))
    good = [t.clone() for t in plan.value_and_grad(p, out=buf)]
    plan.value_and_grad(p, out=buf)
    bad = p.clone()
    bad[2, :] = float("nan")
    pb = p.clone()
    p.copy_(bad)
    v, g, i = plan.value_and_grad(p, out=buf)
    assert int(i[2]) != 0 and bool(torch.isnan(v[2, 0])) and int(i[[0, 1, 3]].abs().sum()) == 0
    assert torch.equal(v[[0, 1, 3]], good[0][[0, 1, 3]])
    p.copy_(pb)
    v, g, i = plan.value_and_grad(p, out=buf)
    assert torch.equal(v, good[0]) and torch.equal(g, good[1]) and int(i.abs().sum()) == 0
    assert plan.graph_replays >= 3


def test_set_hyper_invalidates_the_captured_graph(cuda_device):
    import torch
    from nonstationary_multivariate_gaussian_process_b200 import _lib
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    x, Y, pars = _case("separable", 64, 3, 1, 1)
    plan = LogPosteriorPlan("separable", x, Y, HYPER["separable"])
    p = torch.from_numpy(pars[0]).to(cuda_device)
    buf = (torch.empty((1, _lib.NVALS), dtype=torch.float64, device=cuda_device), torch.empty_like(p),
           torch.empty((1,), dtype=torch.int32, device=cuda_device))
    for _ in range(3):
        plan.value_and_grad(p, out=buf)
    new = dict(HYPER["separable"], mu_tilde_l=0.3, beta_tilde_sigma=0.8, c=0.2)
    plan.set_hyper(new)
    fresh = LogPosteriorPlan("separable", x, Y, new)
    fresh.set_graph(False)
    v0, g0, _ = fresh.value_and_grad(p)
    for _ in range(3):
        v, g, i = plan.value_and_grad(p, out=buf)
        assert torch.equal(v, v0) and torch.equal(g, g0)


def test_map_fit_and_hmc_use_replays_and_match_direct_launches(cuda_device):
    import torch
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    x, Y, pars = _case("separable", 40, 3, 2, 1)
    out = []
    for graph in (True, False):
        plan = LogPosteriorPlan("separable", x, Y, HYPER["separable"])
        plan.set_graph(graph)
        pf, trace, info = plan.map_fit(pars[0], steps=12, lr=0.05)
        smp, acc, U = plan.hmc_sample(pf, sample_size=4, step_size=1e-3, num_steps_in_leap=3, seed=5)
        out.append((pf, trace, smp, acc, U, plan.graph_replays))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    assert torch.equal(out[0][2], out[1][2]) and torch.equal(out[0][4], out[1][4])
    assert out[0][5] >= 15 and out[1][5] == 0
