"""GPU: subjects with different numbers of time points / outputs (the personalized drivers' patients) through RaggedPlans:
grouped by shape, evaluated per group, returned in subject order -- equal to one single-subject plan per subject (to rounding) and to the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HYPER = {"mu_tilde_l": 0.0, "alpha_tilde_l": 5.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 5.0, "beta_L": 1.0, "a": 1.0, "b": 1.0}
SHAPES = [(30, 3), (45, 2), (30, 3), (64, 4), (45, 2), (30, 3), (17, 2)]


def _subjects(model):
    from nonstationary_multivariate_gaussian_process_b200 import synth
    xs, Ys, ps = [], [], []
    for s, (N, M) in enumerate(SHAPES):
        x, Y, _ = synth.sample_subject(N, M, s)
        xs.append(x), Ys.append(Y), ps.append(synth.start_point(model, N, M, s, 0.05))
    return xs, Ys, ps


def test_ragged_equals_single_subject_plans_and_oracle(cuda_device):
    import torch
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan, RaggedPlans
    from oracle import nmgp_oracle as O
    xs, Ys, ps = _subjects("nonseparable")
    rp = RaggedPlans("nonseparable", xs, Ys, HYPER)
    assert list(rp.groups) == [(30, 3), (45, 2), (64, 4), (17, 2)] and rp.groups[(30, 3)] == [0, 2, 5]
    vals, grads, info = rp.value_and_grad([torch.from_numpy(p) for p in ps])
    assert int(info.abs().sum()) == 0 and vals.shape == (len(SHAPES), 6)
    for s in range(len(SHAPES)):
        one = LogPosteriorPlan("nonseparable", xs[s], Ys[s], HYPER)
        v1, g1, _ = one.value_and_grad(torch.from_numpy(ps[s]).to(cuda_device))
        # batch invariance (the engine may pick another blocking for a batch of one: rounding-level differences)
        assert abs(float(v1[0, 0] - vals[s, 0])) <= 1e-12 * abs(float(v1[0, 0])), s
        assert float((g1[0] - grads[s]).norm() / g1[0].norm()) < 1e-10, s
        ov, og = O.value_and_grad("nonseparable", ps[s], Ys[s], xs[s], **HYPER)
        assert abs(float(vals[s, 1]) - float(ov[1])) <= 1e-9 * abs(float(ov[1]))        # likelihood
        assert abs(float(vals[s, 0]) - float(ov[0])) <= 1e-7 * abs(float(ov[0]))        # total: prior conditioning floor
        assert float(torch.linalg.norm(grads[s].cpu() - og) / torch.linalg.norm(og)) < 1e-6
    # value only; wrong parameter length; wrong count
    v2, g2, _ = rp.value_and_grad(ps, need_grad=False)
    assert g2 is None and torch.equal(v2, vals)
    with pytest.raises(ValueError):
        rp.value_and_grad(ps[:-1])
    with pytest.raises(ValueError):
        rp.value_and_grad([ps[1]] + ps[1:])


def test_ragged_map_fit_descends_for_every_subject(cuda_device):
    import torch
    from nonstationary_multivariate_gaussian_process_b200.batched import RaggedPlans
    xs, Ys, ps = _subjects("separable")
    hyper = {"mu_tilde_l": 0.0, "alpha_tilde_l": 1.0, "beta_tilde_l": 1.0, "mu_tilde_sigma": 0.0, "alpha_tilde_sigma": 1.0,
             "beta_tilde_sigma": 1.0, "a": 1.0, "b": 1.0, "c": 10.0}
    rp = RaggedPlans("separable", xs, Ys, hyper)
    v0, _, _ = rp.value_and_grad(ps, need_grad=False)
    fitted, info = rp.map_fit(ps, steps=25, lr=0.01)
    v1, _, _ = rp.value_and_grad(fitted, need_grad=False)
    assert int(info.abs().sum()) == 0 and len(fitted) == len(SHAPES)
    assert all(f.numel() == p.size for f, p in zip(fitted, ps))
    assert bool((v1[:, 0] < v0[:, 0]).all())
