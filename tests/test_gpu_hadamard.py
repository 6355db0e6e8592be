"""GPU parity of the irregularly sampled ("Hadamard") objectives (Utility/logpos.py:465-716) against the reference's golden
vectors (tests/golden/hadamard_*.npz), through the C ABI (nmgp_plan_create_hadamard + nmgp_logpost_grad)."""
import numpy as np
import pytest
import torch

from conftest import hadamard_cases, load_hadamard_golden, rel_err

pytestmark = pytest.mark.gpu

# same tolerances and rationale as tests/test_gpu_parity_golden.py: 1e-9 for the likelihood and for Prior=False; the GP-prior
# terms (cond 1e8 .. 1e10 with the drivers' hyper-parameters) at their conditioning floor
TOL_LOGLIK = 1e-9
TOL_NOPRIOR = 1e-9
TOL_PRIOR = 1e-7
TOL_TOTAL = 1e-7
TOL_GRAD = 1e-6


@pytest.mark.parametrize("engine", ["auto", "left"])
@pytest.mark.parametrize("name", hadamard_cases())
def test_cuda_matches_reference_golden(name, engine, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    g = load_hadamard_golden(name)
    plan = LogPosteriorPlan(g["model"], g["x"], g["y"], g["hyper"], prior=g["prior"], indx=g["indx"])
    plan.set_engine(engine)
    vals, grad, info = plan.value_and_grad_host(torch.from_numpy(g["pars"]))
    plan.close()
    vals, grad, ref = vals.numpy()[0], grad.numpy()[0], g["vals"]
    assert int(info[0]) == 0 and plan.M == g["M"] and grad.shape[0] == g["grad"].shape[0]
    assert rel_err(vals[1], ref[1]) < TOL_LOGLIK, (name, "loglik", vals[1], ref[1])
    if not g["prior"]:
        assert rel_err(vals[0], ref[0]) < TOL_NOPRIOR and rel_err(grad, g["grad"]) < TOL_NOPRIOR, (name, rel_err(grad, g["grad"]))
        return
    assert rel_err(vals[0], ref[0]) < TOL_TOTAL, (name, "total", vals[0], ref[0])
    for k in range(2, len(ref)):
        assert rel_err(vals[k], ref[k]) < TOL_PRIOR, (name, k, vals[k], ref[k])
    assert rel_err(grad, g["grad"]) < TOL_GRAD, (name, "grad", rel_err(grad, g["grad"]))


def test_reference_signatures_and_autograd(cuda_device):
    """nlogpos_obj_hadamard{,_SVC,_S}(pars, x, indx, y, **hyper, verbose=True) + .backward() on CPU leaves, as a driver would."""
    from nonstationary_multivariate_gaussian_process_b200 import logpos
    fns = {"hadamard": logpos.nlogpos_obj_hadamard, "hadamard_svc": logpos.nlogpos_obj_hadamard_SVC,
           "hadamard_s": logpos.nlogpos_obj_hadamard_S}
    for name in ("hadamard_sep_N30_M3_s0_h1_p", "hadamard_svc_N25_M2_s0_h1_p", "hadamard_s_N40_M3_s0_h1_p"):
        g = load_hadamard_golden(name)
        p = torch.from_numpy(g["pars"]).clone().requires_grad_(True)
        out = fns[g["model"]](p, torch.from_numpy(g["x"]), torch.from_numpy(g["indx"]), torch.from_numpy(g["y"]), verbose=True,
                              Prior=g["prior"], **g["hyper"])
        out[0].backward()
        assert len(out) == len(g["vals"])
        for k in range(len(out)):
            assert rel_err(float(out[k]), g["vals"][k]) < TOL_TOTAL, (name, k)
        assert rel_err(p.grad.numpy(), g["grad"]) < TOL_GRAD
        scalar = fns[g["model"]](torch.from_numpy(g["pars"]), torch.from_numpy(g["x"]), torch.from_numpy(g["indx"]),
                                 torch.from_numpy(g["y"]), Prior=g["prior"], **g["hyper"])
        assert scalar.dim() == 0 and rel_err(float(scalar), g["vals"][0]) < TOL_TOTAL


@pytest.mark.parametrize("model", ["hadamard", "hadamard_svc", "hadamard_s"])
def test_batched_subjects_match_the_oracle(model, cuda_device):
    """300 subjects in one plan (left-looking engine) against the CPU oracle on a few of them."""
    from nonstationary_multivariate_gaussian_process_b200 import synth
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    from oracle import nmgp_oracle as O
    N, M, S = 70, 3, 300
    hyper = {"hadamard": {"mu_tilde_l": -1.0, "alpha_tilde_l": 2.0, "beta_tilde_l": 0.05, "mu_tilde_sigma": 0.2,
                          "alpha_tilde_sigma": 1.0, "beta_tilde_sigma": 0.05, "a": 1.0, "b": 1.0, "c": 10.0},
             "hadamard_svc": {"mu_tilde_l": -1.0, "alpha_tilde_l": 2.0, "beta_tilde_l": 0.05, "mu_L": 0.1, "alpha_L": 1.5,
                              "beta_L": 0.05, "a": 1.0, "b": 1.0},
             "hadamard_s": {"mu_tilde_l": 0.0, "sigma_tilde_l": 1.0, "a": 1.0, "b": 1.0, "c": 10.0}}[model]
    cases = [synth.hadamard_case(model, N, M, 500 + s) for s in range(S)]
    xs, ixs, ys, ps = (np.stack(a) for a in zip(*cases))
    plan = LogPosteriorPlan(model, xs, ys, hyper, indx=ixs, M=M)
    vals, grad, info = plan.value_and_grad(torch.from_numpy(ps).cuda())
    plan.close()
    assert int(info.abs().sum()) == 0
    vals, grad = vals.cpu().numpy(), grad.cpu().numpy()
    for s in (0, 137, 299):
        ov, og = O.value_and_grad_hadamard(model, ps[s], xs[s], ixs[s], ys[s], **hyper)
        assert rel_err(vals[s, 0], float(ov[0])) < 1e-9 and rel_err(vals[s, 1], float(ov[1])) < 1e-10
        assert rel_err(grad[s], og.numpy()) < 1e-8, (model, s, rel_err(grad[s], og.numpy()))
