"""GPU: nmgp_hyper_grad (csrc/hyper.cu) -- gradient of -log posterior with respect to the hyper-parameters of the priors,
per subject -- against the golden vectors recorded from the unmodified reference's autograd through its keyword
hyper-parameters (tests/golden/make_golden_hyper.py) and against the oracle's dense closed form."""
import numpy as np
import pytest

from test_hyper_grad_oracle import CASES, load

pytestmark = pytest.mark.gpu


def _scaled_err(got, want):
    return float((np.abs(got - want) / np.maximum(np.abs(want), 1e-3 * np.abs(want).max())).max())


# Per-slot bounds against the reference's autograd, relative to the largest entry of the slot over the fixture's subjects
# (measured on B200, tools/report_hyper_parity.py; bound = ~4 x the worst fixture).  alpha / beta / mu of the GP priors go
# through the ill-conditioned prior covariance (cond 1e8..2e10: the floor of tests/test_prior_conditioning_floor.py), the
# scalar-prior slots do not; the reference's own `b` slot comes from a central difference (make_golden_hyper.py).
def _slot_tol(name):
    if name.startswith("alpha_"):
        return 1e-5          # measured <= 2.8e-6
    if name.startswith("beta_"):
        return 6e-7          # measured <= 1.4e-7
    if name in ("mu_tilde_l", "mu_tilde_sigma", "mu_L"):
        return 3e-7          # measured <= 5.6e-8
    if name == "b":
        return 5e-7          # measured <= 1.1e-7 (finite-difference reference)
    return 1e-12             # a, c, sigma_tilde_l: measured <= 4e-14


@pytest.mark.parametrize("name", CASES)
def test_hyper_grad_matches_reference_autograd(name, cuda_device):
    import torch
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    from oracle import nmgp_oracle as O
    d = load(name)
    plan = LogPosteriorPlan(d["model"], d["x"], d["Y"], d["hyper"])
    pars = torch.from_numpy(d["pars"]).to(cuda_device)
    hg = plan.hyper_grad(pars).cpu().numpy()
    hg2 = plan.hyper_grad(pars).cpu().numpy()          # second call: cached traces
    assert np.array_equal(hg, hg2)
    names = plan.hyper_names()
    ref = d["hgrad"]
    for k, nm in enumerate(names):
        scale = max(float(np.abs(ref[:, k]).max()), 1e-300)
        if d["model"] == "stationary" and nm in ("mu_tilde_l", "sigma_tilde_l"):
            tol = 1e-12      # scalar Normal prior on tilde_l: no covariance involved
        else:
            tol = _slot_tol(nm)
        assert float(np.abs(hg[:, k] - ref[:, k]).max()) / scale < tol, (name, nm, hg[:, k], ref[:, k])
    for s in range(d["pars"].shape[0]):
        orc = O.hyper_grad(d["model"], d["pars"][s], d["x"][s], d["M"], **d["hyper"])
        # the oracle's dense closed form carries the same conditioning noise in the alpha / beta slots
        assert _scaled_err(hg[s], orc) < 2e-5, (name, s, hg[s], orc)
    # evaluation after the hyper-gradient call is unaffected (shared scratch)
    vals, grad, info = plan.value_and_grad(pars)
    assert int(info.abs().sum()) == 0
    assert np.allclose(vals[:, 0].cpu().numpy(), d["vals"], rtol=1e-8)
    # fused pass (nmgp_logpost_grad_hyper): same values, gradient and hyper-gradient, bit for bit
    v2, g2, h2, i2 = plan.value_grad_and_hyper_grad(pars)
    assert torch.equal(v2, vals) and torch.equal(g2, grad) and int(i2.abs().sum()) == 0
    assert np.array_equal(h2.cpu().numpy(), hg)
    v3, g3, h3, _ = plan.value_grad_and_hyper_grad(pars, need_grad=False)
    assert g3 is None and torch.equal(v3, vals) and np.array_equal(h3.cpu().numpy(), hg)


def test_hyper_grad_zero_without_prior(cuda_device):
    import torch
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    d = load([c for c in CASES if "nonseparable" in c][0])
    plan = LogPosteriorPlan(d["model"], d["x"], d["Y"], d["hyper"], prior=False)
    hg = plan.hyper_grad(torch.from_numpy(d["pars"]).to(cuda_device))
    assert float(hg.abs().max()) == 0.0


def test_hyper_grad_finite_difference_of_the_batched_objective(cuda_device):
    """Central differences of the CUDA objective itself (sum over subjects, new plan per perturbed hyper-parameter) for the
    float64 GP-prior slots: the quantity a tied-hyper-prior optimiser steps along."""
    import torch
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    d = load("hyper_nonseparable_N20_M2_s0")
    pars = torch.from_numpy(d["pars"]).to(cuda_device)
    tot = LogPosteriorPlan(d["model"], d["x"], d["Y"], d["hyper"]).hyper_grad(pars).sum(0).cpu().numpy()
    names = ("mu_tilde_l", "alpha_tilde_l", "beta_tilde_l", "mu_L", "alpha_L", "beta_L")
    for i, k in enumerate(names):
        h = 1e-3 * max(abs(d["hyper"][k]), 1.0)   # the objective carries ~1e-9 relative noise (prior cond. ~1e9): wide step
        f = []
        for sgn in (+1, -1):
            hy = dict(d["hyper"])
            hy[k] = hy[k] + sgn * h
            v, _, info = LogPosteriorPlan(d["model"], d["x"], d["Y"], hy).value_and_grad(pars, need_grad=False)
            assert int(info.abs().sum()) == 0
            f.append(float(v[:, 0].sum()))
        fd = (f[0] - f[1]) / (2 * h)
        assert abs(fd - tot[i]) <= 2e-3 * max(abs(tot[i]), 1.0), (k, fd, tot[i])


def test_sweep_vector_sums_successful_subjects(cuda_device):
    import torch
    from nonstationary_multivariate_gaussian_process_b200 import sharding
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    d = load("hyper_nonseparable_N40_M3_s3")
    plan = LogPosteriorPlan(d["model"], d["x"], d["Y"], d["hyper"])
    pars = torch.from_numpy(d["pars"]).to(cuda_device)
    vals, _, info = plan.value_and_grad(pars, need_grad=False)
    hg = plan.hyper_grad(pars)
    info = info.clone()
    info[1] = 3                                  # pretend subject 1 failed
    vec = sharding.local_sweep_vector(vals, info, hg)
    summ, hyp = sharding.all_reduce_sweep(vec, plan.hyper_names())
    keep = [0, 2, 3]
    assert summ["n_failed"] == 1 and summ["n_subjects"] == 4
    assert abs(summ["neg_logpost"] - d["vals"][keep].sum()) < 1e-8 * abs(d["vals"][keep].sum())
    want = d["hgrad"][keep].sum(0)
    got = np.array([hyp[k] for k in plan.hyper_names()])
    assert _scaled_err(got, want[:len(got)]) < 2e-5


def test_set_hyper_equals_new_plan(cuda_device):
    """nmgp_plan_set_hyper: values, gradients and hyper-gradients after an in-place change are those of a plan created with
    the new hyper-parameters, bit for bit; changing them back restores the original results."""
    import torch
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    for name in ("hyper_nonseparable_N40_M3_s3", "hyper_separable_N24_M2_s0", "hyper_stationary_N30_M3_s0"):
        d = load(name)
        pars = torch.from_numpy(d["pars"]).to(cuda_device)
        plan = LogPosteriorPlan(d["model"], d["x"], d["Y"], d["hyper"])
        v0, g0, h0, _ = plan.value_grad_and_hyper_grad(pars)
        new = {k: (v * 1.25 + 0.1) for k, v in d["hyper"].items()}
        plan.set_hyper(new)
        v1, g1, h1, i1 = plan.value_grad_and_hyper_grad(pars)
        fresh = LogPosteriorPlan(d["model"], d["x"], d["Y"], new)
        v2, g2, h2, i2 = fresh.value_grad_and_hyper_grad(pars)
        assert int(i1.abs().sum()) == 0 and int(i2.abs().sum()) == 0
        assert torch.equal(v1, v2) and torch.equal(g1, g2) and torch.equal(h1, h2)
        assert not torch.equal(v0, v1)
        plan.set_hyper(d["hyper"])
        v3, g3, h3, _ = plan.value_grad_and_hyper_grad(pars)
        assert torch.equal(v3, v0) and torch.equal(g3, g0) and torch.equal(h3, h0)


def test_tied_hyper_descent_lowers_the_summed_objective(cuda_device):
    """The loop the north star describes for subjects sharing hyper-priors: sweep -> summed hyper-gradient -> step on the
    shared prior means -> nmgp_plan_set_hyper.  A few small gradient steps on (mu_tilde_l, mu_L) must lower the sum of
    -log posterior over the subjects."""
    import torch
    from nonstationary_multivariate_gaussian_process_b200 import sharding
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    d = load("hyper_nonseparable_N40_M3_s3")
    pars = torch.from_numpy(d["pars"]).to(cuda_device)
    plan = LogPosteriorPlan(d["model"], d["x"], d["Y"], d["hyper"])
    hy = dict(d["hyper"])
    totals = []
    for _ in range(4):
        vals, _, hg, info = plan.value_grad_and_hyper_grad(pars, need_grad=False)
        summ, shared = sharding.all_reduce_sweep(sharding.local_sweep_vector(vals, info, hg), plan.hyper_names())
        totals.append(summ["neg_logpost"])
        for k in ("mu_tilde_l", "mu_L"):
            hy[k] -= 1e-3 * shared[k]      # curvature of the sum in mu_L is ~46: a stable step
        plan.set_hyper(hy)
    assert totals[-1] < totals[0] and all(b <= a for a, b in zip(totals, totals[1:])), totals


def test_sweep_reduce_kernel_equals_the_torch_sums(cuda_device):
    import torch
    from nonstationary_multivariate_gaussian_process_b200 import sharding
    rng = np.random.RandomState(3)
    for S in (1, 7, 1000, 10000):
        vals = torch.from_numpy(rng.standard_normal((S, 6)))
        hg = torch.from_numpy(rng.standard_normal((S, 9)))
        info = torch.from_numpy((rng.rand(S) < 0.1).astype(np.int32) * 4)
        vals[info != 0] = float("nan")
        hg[info != 0] = float("nan")
        want = sharding.local_sweep_vector(vals, info, hg)                       # host tensors: torch sums
        got = sharding.local_sweep_vector(vals.to(cuda_device), info.to(cuda_device), hg.to(cuda_device))
        got2 = sharding.local_sweep_vector(vals.to(cuda_device), info.to(cuda_device), hg.to(cuda_device))
        assert torch.equal(got, got2)
        assert torch.allclose(got.cpu(), want, rtol=1e-12, atol=1e-10), (S, got, want)
        none = sharding.local_sweep_vector(vals.to(cuda_device), info.to(cuda_device), None).cpu()
        assert torch.equal(none[:8], got.cpu()[:8]) and float(none[8:].abs().max()) == 0.0


def test_tied_map_fit_lowers_the_population_objective(cuda_device):
    """sharding.tied_map_fit on one rank: subjects' parameters and the tied hyper-parameters (prior means in linear, prior
    scales in log space) are fitted together; the all-reduced objective falls and the plan ends with the returned values."""
    import torch
    from nonstationary_multivariate_gaussian_process_b200 import sharding
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    d = load("hyper_nonseparable_N40_M3_s3")
    plan = LogPosteriorPlan(d["model"], d["x"], d["Y"], d["hyper"])
    tied = ("mu_tilde_l", "mu_L", "alpha_L")
    pars, hyper, trace = sharding.tied_map_fit(plan, d["pars"], steps=30, lr=0.01, tied=tied, hyper_lr=0.02)
    tot = [t["neg_logpost"] for t in trace]
    assert all(t["n_failed"] == 0 and t["n_subjects"] == 4 for t in trace)
    assert tot[-1] < tot[0] and np.isfinite(tot).all()
    assert all(hyper[k] != d["hyper"][k] for k in tied) and hyper["alpha_L"] > 0
    assert all(hyper[k] == d["hyper"][k] for k in d["hyper"] if k not in tied)
    fresh = LogPosteriorPlan(d["model"], d["x"], d["Y"], hyper)
    v1, _, _ = plan.value_and_grad(pars, need_grad=False)
    v2, _, _ = fresh.value_and_grad(pars, need_grad=False)
    assert torch.equal(v1, v2)
