"""Device-resident HMC (LogPosteriorPlan.hmc_sample: nmgp_hmc_kick / nmgp_hmc_drift / nmgp_hmc_accept around the batched
value+gradient call) against the same leapfrog / Metropolis algorithm run per subject on the CPU with the oracle as the
potential -- the role the external HMC_Sampler plays in the reference drivers (Separable_model.py:209-210)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def cpu_chain(model, q0, Y, x, hyper, momenta, log_u, eps, L):
    from oracle import nmgp_oracle as O
    q = q0.clone()
    vals, grad = O.value_and_grad(model, q, Y, x, **hyper)
    U = float(vals[0])
    out, acc = [], []
    for it in range(momenta.shape[0]):
        p0 = momenta[it].clone()
        p, qp = p0.clone(), q.clone()
        p -= 0.5 * eps * grad
        for l in range(L):
            qp += eps * p
            vp, gp = O.value_and_grad(model, qp, Y, x, **hyper)
            p -= (eps if l + 1 < L else 0.5 * eps) * gp
        Up = float(vp[0])
        dH = (U + 0.5 * float(p0 @ p0)) - (Up + 0.5 * float(p @ p))
        ok = np.isfinite(dH) and float(log_u[it]) < dH
        if ok:
            q, U, grad = qp, Up, gp
        out.append(q.clone()); acc.append(ok)
    return torch.stack(out), np.array(acc), U


@pytest.mark.parametrize("model,N,M", [("separable", 24, 2), ("nonseparable", 20, 3), ("stationary", 30, 3)])
def test_hmc_chain_matches_cpu_algorithm(model, N, M, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import synth
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    hyper = {"separable": {"mu_tilde_l": -1.0, "alpha_tilde_l": 2.0, "beta_tilde_l": 0.05, "mu_tilde_sigma": 0.2,
                           "alpha_tilde_sigma": 1.0, "beta_tilde_sigma": 0.05, "a": 1.0, "b": 1.0, "c": 10.0},
             "nonseparable": {"mu_tilde_l": -1.0, "alpha_tilde_l": 2.0, "beta_tilde_l": 0.05, "mu_L": 0.1, "alpha_L": 1.5,
                              "beta_L": 0.05, "a": 1.0, "b": 1.0},
             "stationary": {"mu_tilde_l": 0.0, "sigma_tilde_l": 1.0, "a": 1.0, "b": 1.0, "c": 10.0}}[model]
    S, n_s, L, eps = 3, 6, 4, 1e-4
    subs = [synth.sample_subject(N, M, 70 + s)[:2] + (synth.start_point(model, N, M, 70 + s, 0.05),) for s in range(S)]
    xs, Ys, ps = (np.stack(a) for a in zip(*subs))
    P = ps.shape[1]
    rng = np.random.default_rng(9)
    momenta = torch.from_numpy(rng.standard_normal((n_s, S, P)))
    log_u = torch.full((n_s, S), -50.0, dtype=torch.float64)     # accept whenever the proposal is finite ...
    log_u[2, :] = 50.0                                           # ... except sample 2 (all chains) and sample 4 of chain 1
    log_u[4, 1] = 50.0
    plan = LogPosteriorPlan(model, xs, Ys, hyper)
    samples, rate, U = plan.hmc_sample(torch.from_numpy(ps), n_s, eps, L, momenta=momenta, log_uniforms=log_u)
    plan.close()
    samples = samples.cpu()
    assert samples.shape == (n_s, S, P)
    for s in range(S):
        ref, acc, Uref = cpu_chain(model, torch.from_numpy(ps[s]), torch.from_numpy(Ys[s]), torch.from_numpy(xs[s]), hyper,
                                   momenta[:, s], log_u[:, s], eps, L)
        assert not acc[2] and (s != 1 or not acc[4]) and acc.sum() >= 3          # forced rejections; the rest mostly move
        err = float((samples[:, s] - ref).abs().max() / ref.abs().max())
        assert err < 1e-8, (model, s, err)
        assert torch.equal(samples[2, s], samples[1, s])          # a rejected proposal duplicates the current state
        assert abs(float(rate[s]) - acc.mean()) < 1e-12 and abs(float(U[s]) - Uref) / abs(Uref) < 1e-8


def test_hmc_device_generator_is_reproducible_and_moves(cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import synth
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    N, M, S = 40, 3, 16
    hyper = {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0,
             "a": 1e-2, "b": 1e-2}
    subs = [synth.sample_subject(N, M, s)[:2] + (synth.start_point("nonseparable", N, M, s, 0.02),) for s in range(S)]
    xs, Ys, ps = (np.stack(a) for a in zip(*subs))
    plan = LogPosteriorPlan("nonseparable", xs, Ys, hyper)
    a, ra, _ = plan.hmc_sample(torch.from_numpy(ps), 8, 1e-4, 5, seed=3, keep_every=2)
    b, rb, _ = plan.hmc_sample(torch.from_numpy(ps), 8, 1e-4, 5, seed=3, keep_every=2)
    plan.close()
    assert a.shape == (4, S, ps.shape[1]) and torch.equal(a, b) and torch.equal(ra, rb)
    assert float(ra.mean()) > 0.5 and float((a[-1].cpu() - torch.from_numpy(ps)).abs().max()) > 0
