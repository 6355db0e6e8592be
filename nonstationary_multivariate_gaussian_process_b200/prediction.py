"""Posterior prediction of the nonseparable model with the reference's names, signatures and return conventions
(Utility/prediction.py:1038-1262), evaluated by the B200 CUDA library through the C ABI.

    point_predmap_inhomogeneous_sampling     (n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, x_star, **hyper)  :1038
    pointwise_predmap_inhomogeneous_sampling (n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, grids,  **hyper)  :1194
    test_predmap_inhomogeneous_sampling      (n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, x_test, **hyper)  :1237

(callers: Nonseparable_Model/Nonseparable_model.py:377, 387, 399 with n_sample = 100 and 201 grid points.)

What runs where.  The reference draws, per new input and per sample, tilde_l* and uL* from the GP priors conditioned on
the MAP values and then y from the predictive normal given them, all from torch's global generator.  Those draws stay on
the host, issued with the same shapes in the same order (`Normal(loc, scale).sample()` is `z * scale + loc` with
`z = empty(shape).normal_()`), so that a run seeded like a reference run consumes the same random stream.  Everything the
draws need is computed on the GPU in two batched calls over all new inputs and samples at once:
`nmgp_predict_prior_moments` (conditional moments of the priors) and `nmgp_predict_moments` (predictive mean / variance;
one factorisation + inverse of the covariance per call instead of the reference's n x n `symeig` + `cholesky` per new
input and per sample).  There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _lib, logpos, utils


def _plan_and_pars(tilde_l, uL_vecs, tilde_sigma2_err, Y, x, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L,
                   kwargs):
    torch = _lib.require_cuda()
    hyper = dict(mu_tilde_l=mu_tilde_l, alpha_tilde_l=alpha_tilde_l, beta_tilde_l=beta_tilde_l, mu_L=mu_L, alpha_L=alpha_L,
                 beta_L=beta_L, a=kwargs.get("a", 1), b=kwargs.get("b", 1))   # same key as the MAP loop's plan
    plan = logpos._get_plan("nonseparable", Y, x, hyper, True)
    pars = torch.cat([torch.as_tensor(tilde_l, dtype=torch.float64).reshape(-1),
                      torch.as_tensor(uL_vecs, dtype=torch.float64).reshape(-1),
                      torch.as_tensor(tilde_sigma2_err, dtype=torch.float64).reshape(1)]).detach()
    return plan, pars


def _sample_all(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, grids, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L,
                alpha_L, beta_L, mode, kwargs):
    """All draws for the new inputs `grids` [G]: dict with tl_star [G,ns], uL_star [G,ns,T], y [G,ns,M] (as applicable)."""
    torch = _lib.require_cuda()
    plan, pars = _plan_and_pars(tilde_l, uL_vecs, tilde_sigma2_err, Y, x, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L,
                                alpha_L, beta_L, kwargs)
    M = plan.M
    T = M * (M + 1) // 2
    g = torch.as_tensor(grids, dtype=torch.float64).reshape(-1)
    G = int(g.numel())
    mu_l, s2_l, mu_u, s2_u = (t[0].cpu() for t in plan.predict_prior_moments(pars, g))
    # standard-normal draws in the reference's order: per new input, per sample: tilde_l* (1), uL* (T), y (M)
    f64 = torch.float64
    z_l = torch.empty((G, n_sample), dtype=f64) if mode in ("y", "smoothness") else None
    z_u = torch.empty((G, n_sample, T), dtype=f64) if mode in ("y", "cov") else None
    z_y = torch.empty((G, n_sample, M), dtype=f64) if mode == "y" else None
    for gi in range(G):
        for s in range(n_sample):
            if z_l is not None:
                z_l[gi, s] = torch.empty((), dtype=f64).normal_()       # prediction.py:1104 / :1113
            if z_u is not None:
                z_u[gi, s] = torch.empty(T, dtype=f64).normal_()        # :1107 / :1124
            if z_y is not None:
                z_y[gi, s] = torch.empty(M, dtype=f64).normal_()        # :1169
    out = {}
    if z_l is not None:
        out["tl_star"] = z_l.mul(torch.sqrt(s2_l).unsqueeze(1)).add(mu_l.unsqueeze(1))
    if z_u is not None:
        out["uL_star"] = z_u.mul(torch.sqrt(s2_u).reshape(G, 1, 1)).add(mu_u.unsqueeze(1))
    if mode == "y":
        mu_f, s2_y, info = plan.predict_moments(pars, g, out["tl_star"].unsqueeze(0), out["uL_star"].unsqueeze(0))
        if int(info[0]) != 0:
            raise _lib.NmgpError(f"prediction: the covariance is not positive definite (pivot {int(info[0])})")
        out["mu_f"], out["s2_y"] = mu_f[0].cpu(), s2_y[0].cpu()
        out["y"] = z_y.mul(torch.sqrt(out["s2_y"])).add(out["mu_f"])
    return out, M


def _summaries(y):
    """Per new input: 2.5 / 97.5 percentiles, mean and std over the samples (prediction.py:1186-1190)."""
    ys = y.numpy()
    q = np.stack([np.percentile(s, q=[2.5, 97.5], axis=0) for s in ys])
    mean = np.stack([np.mean(s, axis=0) for s in ys])
    std = np.stack([np.std(s, axis=0) for s in ys])
    return q, mean, std


def _lower_factors(uL_star, M):
    torch = _lib.require_cuda()
    G, ns, T = uL_star.shape
    Lv = utils.uLvec2Lvec(uL_star.reshape(-1, T), M)
    idx = torch.tril_indices(M, M)
    out = torch.zeros((G * ns, M, M), dtype=uL_star.dtype)
    out[:, idx[0], idx[1]] = Lv
    return out.reshape(G, ns, M, M).numpy()


def pointwise_predmap_inhomogeneous_sampling(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, grids, mu_tilde_l,
                                             alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, pred_smoothness=False,
                                             pred_cov=False, *args, **kwargs):
    """Posterior predictive summaries on a grid from MAP estimates (prediction.py:1194-1235).
    Returns (quantiles [G,2,M], mean [G,M], std [G,M]); with pred_smoothness=True the sampled tilde_l* [G,n_sample];
    with pred_cov=True the sampled factors L* [G,n_sample,M,M]."""
    mode = "smoothness" if pred_smoothness else ("cov" if pred_cov else "y")
    out, M = _sample_all(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, grids, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                         mu_L, alpha_L, beta_L, mode, kwargs)
    if pred_smoothness:
        return out["tl_star"].numpy()
    if pred_cov:
        return _lower_factors(out["uL_star"], M)
    return _summaries(out["y"])


def point_predmap_inhomogeneous_sampling(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, x_star, mu_tilde_l,
                                         alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, pred_smoothness=False,
                                         pred_cov=False, *args, **kwargs):
    """One new input x_star (prediction.py:1038-1192): (quantiles [2,M], mean [M], std [M]), or the sampled
    tilde_l* [n_sample] / L* [n_sample,M,M]."""
    res = pointwise_predmap_inhomogeneous_sampling(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x,
                                                   np.asarray(x_star, dtype=np.float64).reshape(1), mu_tilde_l,
                                                   alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L,
                                                   pred_smoothness=pred_smoothness, pred_cov=pred_cov, **kwargs)
    if pred_smoothness or pred_cov:
        return res[0]
    return res[0][0], res[1][0], res[2][0]


def test_predmap_inhomogeneous_sampling(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, x_test, mu_tilde_l,
                                        alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, *args, **kwargs):
    """Posterior predictive summaries at test inputs (prediction.py:1237-1262): (quantiles, mean, std)."""
    return pointwise_predmap_inhomogeneous_sampling(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, x_test, mu_tilde_l,
                                                    alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, **kwargs)


test_predmap_inhomogeneous_sampling.__test__ = False   # not a pytest test
