"""Edge shapes through the C ABI against the CPU oracle: one time point, one output, the maximum number of outputs (16),
matrix dimensions on and next to the 64-block boundary, empty plans, a single new input / sample in the predictors."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

HYPER = {
    "nonseparable": {"mu_tilde_l": -1.0, "alpha_tilde_l": 2.0, "beta_tilde_l": 0.05, "mu_L": 0.1, "alpha_L": 1.5, "beta_L": 0.05,
                     "a": 1.0, "b": 1.0},
    "separable": {"mu_tilde_l": -1.0, "alpha_tilde_l": 2.0, "beta_tilde_l": 0.05, "mu_tilde_sigma": 0.2, "alpha_tilde_sigma": 1.0,
                  "beta_tilde_sigma": 0.05, "a": 1.0, "b": 1.0, "c": 10.0},
    "stationary": {"mu_tilde_l": 0.0, "sigma_tilde_l": 1.0, "a": 1.0, "b": 1.0, "c": 10.0},
}
SHAPES = [
    ("nonseparable", 1, 1), ("nonseparable", 2, 16), ("nonseparable", 64, 1), ("nonseparable", 65, 1), ("nonseparable", 13, 5),
    ("nonseparable", 32, 2), ("nonseparable", 4, 16), ("nonseparable", 3, 7), ("nonseparable", 9, 15),
    ("separable", 1, 1), ("separable", 64, 1), ("separable", 65, 3), ("separable", 2, 16), ("separable", 129, 2),
    ("stationary", 3, 1), ("stationary", 128, 2), ("stationary", 1, 2), ("stationary", 7, 16),
]


@pytest.mark.parametrize("model,N,M", SHAPES)
def test_edge_shapes_match_the_oracle(model, N, M, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import synth
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    from oracle import nmgp_oracle as O
    S = 3
    subs = [synth.sample_subject(N, M, 900 + s)[:2] + (synth.start_point(model, N, M, 900 + s, 0.05),) for s in range(S)]
    xs, Ys, ps = (np.stack(a) for a in zip(*subs))
    for engine in ("auto", "left"):
        plan = LogPosteriorPlan(model, xs, Ys, HYPER[model])
        plan.set_engine(engine)
        vals, grad, info = plan.value_and_grad_host(torch.from_numpy(ps))
        plan.close()
        assert int(info.abs().sum()) == 0
        for s in range(S):
            ov, og = O.value_and_grad(model, ps[s], Ys[s], xs[s], **HYPER[model])
            for k in range(ov.numel()):
                assert rel_err(float(vals[s, k]), float(ov[k])) < 1e-8, (model, N, M, engine, s, k, float(vals[s, k]), float(ov[k]))
            # the reference differentiates through two `symeig` calls; with (near-)degenerate eigenvalues their backward returns
            # NaN (the 1/(lambda_i - lambda_j) terms) -- compare where the reference has a number
            ok = np.isfinite(og.numpy())
            assert np.isfinite(grad[s].numpy()).all() and ok.sum() >= 2
            assert rel_err(grad[s].numpy()[ok], og.numpy()[ok]) < 1e-7, (model, N, M, engine, s)


@pytest.mark.parametrize("model", ["stationary", "separable", "nonseparable"])
def test_empty_plan(model, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    N, M = 10, 2
    plan = LogPosteriorPlan(model, np.zeros((0, N)), np.zeros((0, N, M)), HYPER[model])
    vals, grad, info = plan.value_and_grad(torch.zeros((0, plan.P), dtype=torch.float64, device="cuda"))
    vh, gh, ih = plan.value_and_grad_host(torch.zeros((0, plan.P), dtype=torch.float64))
    plan.close()
    assert vals.shape == (0, 6) and grad.shape == (0, plan.P) and info.shape == (0,) and vh.shape == (0, 6)


def test_single_new_input_and_sample(cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import synth
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    from oracle import nmgp_oracle as O
    from oracle import nmgp_predict_oracle as PO
    for N, M in [(1, 1), (5, 16), (64, 1), (70, 2)]:
        T = M * (M + 1) // 2
        x, Y, _ = synth.sample_subject(N, M, 950)
        p = synth.start_point("nonseparable", N, M, 950, 0.05)
        hy = {k: v for k, v in HYPER["nonseparable"].items() if k not in ("a", "b")}
        plan = LogPosteriorPlan("nonseparable", x, Y, HYPER["nonseparable"])
        xs = torch.tensor([0.37], dtype=torch.float64)
        mu_l, s2_l, mu_u, s2_u = plan.predict_prior_moments(torch.from_numpy(p), xs)
        mu_f, s2_y, info = plan.predict_moments(torch.from_numpy(p), xs, mu_l.unsqueeze(2), mu_u.unsqueeze(2))
        plan.close()
        pt = torch.from_numpy(p)
        pct, Lv = PO.pointwise_predict_plugin(pt[:N], pt[N:N + N * T], pt[-1], torch.from_numpy(Y), torch.from_numpy(x), xs, **hy)
        assert int(info[0]) == 0
        assert rel_err(mu_f[0, 0, 0].cpu().numpy(), pct[0, 1]) < 1e-8, (N, M)
        assert rel_err(np.sqrt(s2_y[0, 0, 0].cpu().numpy()), (pct[0, 2] - pct[0, 1]) / 1.96) < 1e-7, (N, M)


@pytest.mark.parametrize("log_s2", [-4.0, -6.0, -8.0, -11.5])
def test_guarded_takahashi_sweep_on_ill_conditioned_covariances_at_its_size_limit(log_s2, cuda_device):
    """The Takahashi inverse sweep (what the 10 000-subject sweep uses for <= 16 block columns) is not backward stable: the
    error of the trailing inverse block is carried into every new block column, and for smooth covariances with little
    noise it grows geometrically (measured, profiles/r02_takahashi_stress.txt: n = 1024, noise variance e^-8: gradient off by
    5e-3; e^-11.5: garbage).  The production path therefore routes ill-conditioned matrices -- (min pivot / max pivot)^2 of
    the factor below a threshold -- to the backward-stable W^T W inverse on the device (engine_potri_ll_guarded).  At the
    sweep's size limit (n = 1024 = 16 blocks), from the drivers' noise level down to e^-11.5, engine 'left' (guarded sweep)
    must agree with 'left_stable' (W^T W for everything) inside the 1e-9 contract, so that a subject's gradient does not
    depend on which path its batch size selects (api.cu:run_potri)."""
    from nonstationary_multivariate_gaussian_process_b200 import synth
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    N, M, S = 128, 8, 4
    subs = []
    for s in range(S):
        x, Y, _ = synth.sample_subject(N, M, 40 + s)
        p = synth.start_point("nonseparable", N, M, 40 + s, 0.02)
        p[-1] = log_s2
        subs.append((x, Y, p))
    xs, Ys, ps = (np.stack(a) for a in zip(*subs))
    out = {}
    for engine in ("left", "left_stable"):
        plan = LogPosteriorPlan("nonseparable", xs, Ys, HYPER["nonseparable"], prior=False)
        plan.set_engine(engine)
        vals, grad, info = plan.value_and_grad_host(torch.from_numpy(ps))
        plan.close()
        assert int(info.abs().sum()) == 0
        out[engine] = (vals.numpy(), grad.numpy())
    for s in range(S):
        ev = rel_err(out["left"][0][s, 0], out["left_stable"][0][s, 0])
        eg = rel_err(out["left"][1][s], out["left_stable"][1][s])
        assert ev < 1e-11 and eg < 1e-9, (log_s2, s, ev, eg)
