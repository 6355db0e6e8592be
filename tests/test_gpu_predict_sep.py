"""GPU parity of the separable / stationary prediction path (Utility/prediction.py:34-459, 1566-1692) against the reference's
golden outputs (tests/golden/predictsep_*.npz, predictstat_*.npz), through the C ABI (`nmgp_predict_moments_sep`)."""
import numpy as np
import pytest
import torch

from conftest import load_predict_golden, max_rel, predictsep_cases, predictstat_cases

pytestmark = pytest.mark.gpu

TOL_MOMENTS = 1e-9       # predictive moments given the sampled (tilde_l*, tilde_sigma*)
TOL_PRIOR_WELL = 1e-9    # conditional prior moments, well-conditioned prior covariances (hyper index 1)
TOL_PRIOR_ILL_MEAN = 1e-5   # drivers' hyper-parameters: see tests/test_gpu_predict.py for the conditioning argument
TOL_PRIOR_ILL_SD = 2e-3


def well_conditioned(name):
    return "_h1" in name


def split(g):
    N, M = g["N"], g["M"]
    T = M * (M + 1) // 2
    p = torch.from_numpy(g["pars"])
    return p[:N], p[N:2 * N], p[2 * N:2 * N + T], p[-1]


@pytest.mark.parametrize("name", predictsep_cases())
def test_separable_moments_and_prior_conditionals(name, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    from nonstationary_multivariate_gaussian_process_b200 import prediction
    g = load_predict_golden(name)
    M = g["M"]
    plan = LogPosteriorPlan("separable", g["x"], g["Y"], dict(g["hyper"], a=1, b=1, c=10))
    pars = torch.from_numpy(g["pars"])
    mu_l, s2_l, mu_s, s2_s = (t[0].cpu().numpy() for t in plan.predict_prior_moments(pars, g["grids"]))
    ref = dict(mu_l=g["l_loc"][:, 0], sd_l=g["l_scale"][:, 0], mu_s=g["s_loc"][:, 0], sd_s=g["s_scale"][:, 0])
    if well_conditioned(name):
        assert max_rel(mu_l, ref["mu_l"]) < TOL_PRIOR_WELL and max_rel(np.sqrt(s2_l), ref["sd_l"]) < TOL_PRIOR_WELL
        assert max_rel(mu_s[:, 0], ref["mu_s"]) < TOL_PRIOR_WELL and max_rel(np.sqrt(s2_s), ref["sd_s"]) < TOL_PRIOR_WELL
    else:
        assert max_rel(mu_l, ref["mu_l"]) < TOL_PRIOR_ILL_MEAN and max_rel(mu_s[:, 0], ref["mu_s"]) < TOL_PRIOR_ILL_MEAN
        assert np.abs(np.sqrt(s2_l) - ref["sd_l"]).max() < TOL_PRIOR_ILL_SD
        assert np.abs(np.sqrt(s2_s) - ref["sd_s"]).max() < TOL_PRIOR_ILL_SD
    # moments for exactly the reference's draws (pointwise_predmap_sampling: a2 = sigma*^2 diag(B_f), prediction.py:254)
    for engine in ("auto", "left"):
        plan.set_engine(engine)
        mu_f, quad, info = plan.predict_moments_sep(pars, g["grids"], torch.from_numpy(g["l_draw"])[None],
                                                    torch.from_numpy(g["s_draw"])[None])
        assert int(info[0]) == 0
        mu_f, quad = mu_f[0].cpu().numpy(), quad[0].cpu().numpy()
        a2 = np.exp(g["s_draw"])[..., None] ** 2 * prediction._B_diag(split(g)[2], M).numpy().reshape(1, 1, -1)
        s2y = a2 - quad + np.exp(g["pars"][-1])
        assert max_rel(mu_f, g["y_loc"]) < TOL_MOMENTS, (engine, max_rel(mu_f, g["y_loc"]))
        assert max_rel(np.sqrt(s2y), g["y_scale"]) < TOL_MOMENTS, (engine, max_rel(np.sqrt(s2y), g["y_scale"]))
    plan.close()


@pytest.mark.parametrize("name", predictsep_cases())
def test_separable_reference_signatures(name, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import prediction
    g = load_predict_golden(name)
    N, M = g["N"], g["M"]
    T = M * (M + 1) // 2
    tl, ts, uL, te = split(g)
    Y, x, grids = torch.from_numpy(g["Y"]), torch.from_numpy(g["x"]), torch.from_numpy(g["grids"])
    tol = 1e-8 if well_conditioned(name) else 5e-3
    mp = prediction.pointwise_predmap(tl, ts, uL, te, Y, x, grids, **g["hyper"])
    assert tuple(mp.shape) == g["map_percentiles"].shape and max_rel(mp.numpy(), g["map_percentiles"]) < tol
    assert max_rel(prediction.point_predmap(tl, ts, uL, te, Y, x, grids[1], **g["hyper"]).numpy(), g["map_percentiles"][1]) < tol
    torch.manual_seed(1000 + g["seed"])
    q, mean, std = prediction.pointwise_predmap_sampling(g["n_sample"], tl, ts, uL, te, Y, x, grids, **g["hyper"])
    assert q.shape == g["quantiles"].shape
    assert max_rel(q, g["quantiles"]) < tol and max_rel(mean, g["mean"]) < tol and max_rel(std, g["std"]) < 10 * tol
    hp = torch.from_numpy(g["hist_pars"])
    H = int(g["hist_n_sample"])
    torch.manual_seed(5000 + g["seed"])
    hy = prediction.pointwise_predsample(hp[:, :N], hp[:, N:2 * N], hp[:, 2 * N:2 * N + T], hp[:, -1], Y, x,
                                         torch.from_numpy(g["hist_grids"]), N_sample=H, **g["hyper"])
    assert hy.shape == g["hist_y"].shape and max_rel(hy, g["hist_y"]) < tol


@pytest.mark.parametrize("name", predictstat_cases())
def test_stationary_reference_signatures(name, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import prediction
    g = load_predict_golden(name)
    M = g["M"]
    T = M * (M + 1) // 2
    p = torch.from_numpy(g["pars"])
    Y, x, grids = torch.from_numpy(g["Y"]), torch.from_numpy(g["x"]), torch.from_numpy(g["grids"])
    mp = prediction.pointwise_predmap_S(p[0], p[1], p[2:2 + T], p[-1], Y, x, grids)
    assert tuple(mp.shape) == g["map_percentiles"].shape and max_rel(mp.numpy(), g["map_percentiles"]) < TOL_MOMENTS
    tm, tsd = prediction.test_predmap_S(p[0], p[1], p[2:2 + T], p[-1], Y, x, x[:6])
    assert max_rel(tm.numpy(), g["test_mean"]) < TOL_MOMENTS and max_rel(tsd.numpy(), g["test_std"]) < TOL_MOMENTS
    hp = torch.from_numpy(g["hist_pars"])
    np.random.seed(7000 + g["seed"])
    hy = prediction.pointwise_predsample_S(hp[:, 0], hp[:, 1], hp[:, 2:2 + T], hp[:, -1], Y, x, grids)
    assert hy.shape == g["hist_y"].shape and max_rel(hy, g["hist_y"]) < TOL_MOMENTS
