"""Posterior prediction of the nonseparable model with the reference's names, signatures and return conventions
(Utility/prediction.py:1038-1262), evaluated by the B200 CUDA library through the C ABI.

    point_predmap_inhomogeneous_sampling     (n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, x_star, **hyper)  :1038
    pointwise_predmap_inhomogeneous_sampling (n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, grids,  **hyper)  :1194
    test_predmap_inhomogeneous_sampling      (n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, x_test, **hyper)  :1237
    point_/pointwise_/test_predmap_inhomogeneous      (plug-in of the conditional means, no sampling)               :912-1036
    point_/pointwise_/test_predsample_inhomogeneous   (one draw per posterior sample of an HMC history)             :1265-1400

and of the separable and stationary models (block form of the Kronecker covariance, `nmgp_predict_moments_sep`):

    point_/pointwise_/test_predmap, ..._predmap_sampling, ..._predsample                                            :34-459
    pointwise_predmap_S, test_predmap_S, pointwise_predsample_S, test_predsample_S                                  :1566-1692

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

from . import _lib, logpos, settings, utils


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


def _draw_interleaved(*bufs):
    """Fill the [G, n, k]-shaped (or [G, n]) float64 buffers with standard-normal draws in the reference's order -- per
    new input, per sample: one draw for each buffer in turn -- consuming torch's CPU random stream exactly as the reference's
    `Normal(loc, scale).sample()` calls do (`z = empty(shape).normal_()`, one call per quantity: a tensor of fewer than 16
    elements draws element by element, a larger one through the vectorised fill, so the grouping into calls matters and is
    kept).  In-place `normal_()` on pre-split views of the preallocated buffers: same stream, a tenth of the Python
    overhead of allocating a tensor per draw (0.49 s -> 0.20 s for the drivers' 201 x 100 grid; what is left is the dispatch of 60 300 calls)."""
    views = []
    for b in bufs:
        flat = b.reshape(-1) if b.dim() == 2 else b.reshape(-1, b.shape[-1])
        views.append(flat.split(1) if b.dim() == 2 else flat.unbind(0))
    if len(views) == 1:
        for (a,) in zip(*views):
            a.normal_()
    elif len(views) == 2:
        for a, b in zip(*views):
            a.normal_()
            b.normal_()
    else:
        for a, b, c in zip(*views):
            a.normal_()
            b.normal_()
            c.normal_()


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
    # prediction.py:1104 / :1113 (tilde_l*), :1107 / :1124 (uL*), :1169 (y)
    _draw_interleaved(*[z for z in (z_l, z_u, z_y) if z is not None])
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


# ------------------------------------------------------------------------------------- plug-in variant (no sampling)
def pointwise_predmap_inhomogeneous(tilde_l, uL_vecs, tilde_sigma2_err, Y, x, grids, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                                    mu_L, alpha_L, beta_L, *args, **kwargs):
    """Plug-in prediction (prediction.py:990-1012): tilde_l* and uL* are set to their conditional means.
    Returns (percentiles [G,3,M] = mu_f -/+ 1.96 sd, L* vectors [G,T]) as CPU tensors."""
    torch = _lib.require_cuda()
    plan, pars = _plan_and_pars(tilde_l, uL_vecs, tilde_sigma2_err, Y, x, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L,
                                alpha_L, beta_L, kwargs)
    g = torch.as_tensor(grids, dtype=torch.float64).reshape(-1)
    mu_l, _, mu_u, _ = plan.predict_prior_moments(pars, g)
    mu_f, s2_y, info = plan.predict_moments(pars, g, mu_l.unsqueeze(2), mu_u.unsqueeze(2))
    if int(info[0]) != 0:
        raise _lib.NmgpError(f"prediction: the covariance is not positive definite (pivot {int(info[0])})")
    mu_f, sd = mu_f[0, :, 0].cpu(), torch.sqrt(s2_y[0, :, 0]).cpu()
    pct = torch.stack([mu_f - 1.96 * sd, mu_f, mu_f + 1.96 * sd], dim=1)        # prediction.py:984
    return pct, utils.uLvec2Lvec(mu_u[0].cpu(), plan.M)


def point_predmap_inhomogeneous(tilde_l, uL_vecs, tilde_sigma2_err, Y, x, x_star, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                                mu_L, alpha_L, beta_L, *args, **kwargs):
    """(prediction.py:912-988): (percentiles [3,M], L* vector [T])."""
    pct, Lv = pointwise_predmap_inhomogeneous(tilde_l, uL_vecs, tilde_sigma2_err, Y, x,
                                              np.asarray(x_star, dtype=np.float64).reshape(1), mu_tilde_l, alpha_tilde_l,
                                              beta_tilde_l, mu_L, alpha_L, beta_L, **kwargs)
    return pct[0], Lv[0]


def test_predmap_inhomogeneous(tilde_l, L_vecs, tilde_sigma2_err, Y, x, x_test, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L,
                               alpha_L, beta_L, *args, **kwargs):
    """(prediction.py:1014-1036)"""
    return pointwise_predmap_inhomogeneous(tilde_l, L_vecs, tilde_sigma2_err, Y, x, x_test, mu_tilde_l, alpha_tilde_l,
                                           beta_tilde_l, mu_L, alpha_L, beta_L, **kwargs)


test_predmap_inhomogeneous.__test__ = False


# ------------------------------------------------------------------------------------- posterior-sample variant
_HIST_PLANS = {}


def _history_plan(H, Y, x, hyper):
    return _replicated_plan("nonseparable", H, Y, x, hyper)


def pointwise_predsample_inhomogeneous(tilde_l_hist, uL_vecs_hist, tilde_sigma2_err_hist, Y, x, grids, mu_tilde_l,
                                       alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, N_sample, *args, **kwargs):
    """One predictive draw per posterior sample of a parameter history, at every new input (prediction.py:1359-1378):
    returns a numpy array [G, N_sample, M].  Every history entry has its own covariance: the entries are evaluated as the
    'subjects' of ONE batched plan (one factorisation + inverse each), all new inputs at once.
    As in the reference (prediction.py:1300-1311) the L* conditional is taken on the CONSTRAINED triangles L_vecs and the
    drawn vector is used as the factor itself."""
    torch = _lib.require_cuda()
    f64 = torch.float64
    tl_h = torch.as_tensor(tilde_l_hist, dtype=f64)[-N_sample:]
    ul_h = torch.as_tensor(uL_vecs_hist, dtype=f64)[-N_sample:]
    ts_h = torch.as_tensor(tilde_sigma2_err_hist, dtype=f64).reshape(-1)[-N_sample:]
    H, N = int(tl_h.shape[0]), int(tl_h.shape[1])
    hyper = dict(mu_tilde_l=mu_tilde_l, alpha_tilde_l=alpha_tilde_l, beta_tilde_l=beta_tilde_l, mu_L=mu_L, alpha_L=alpha_L,
                 beta_L=beta_L, a=kwargs.get("a", 1), b=kwargs.get("b", 1))
    plan = _history_plan(H, Y, x, hyper)
    M = plan.M
    T = M * (M + 1) // 2
    g = torch.as_tensor(grids, dtype=f64).reshape(-1)
    G = int(g.numel())
    pars = torch.cat([tl_h, ul_h, ts_h.unsqueeze(1)], dim=1).detach()
    L_h = utils.uLvec2Lvec(ul_h.reshape(H * N, T), M).reshape(H, N * T)
    pars_L = torch.cat([tl_h, L_h, ts_h.unsqueeze(1)], dim=1).detach()          # residuals L_vecs - mu_L, prediction.py:1302
    mu_l, s2_l, _, _ = (t.cpu() for t in plan.predict_prior_moments(pars, g))    # [H,G]
    _, _, mu_L_c, s2_L_c = (t.cpu() for t in plan.predict_prior_moments(pars_L, g))
    z_l = torch.empty((G, H), dtype=f64)
    z_u = torch.empty((G, H, T), dtype=f64)
    z_y = torch.empty((G, H, M), dtype=f64)
    _draw_interleaved(z_l, z_u, z_y)   # the reference's order: per new input, per history entry: tilde_l*, L*, y
    tl_star = z_l.t().mul(torch.sqrt(s2_l)).add(mu_l)                            # [H,G]
    L_star = z_u.transpose(0, 1).mul(torch.sqrt(s2_L_c).unsqueeze(2)).add(mu_L_c)   # [H,G,T]
    mu_f, s2_y, info = plan.predict_moments(pars, g, tl_star.unsqueeze(2), L_star.unsqueeze(2), raw_factor=True)
    if int(info.abs().sum()) != 0:
        raise _lib.NmgpError("prediction: a covariance of the history is not positive definite")
    mu_f, sd = mu_f[:, :, 0].cpu().transpose(0, 1), torch.sqrt(s2_y[:, :, 0]).cpu().transpose(0, 1)   # [G,H,M]
    return z_y.mul(sd).add(mu_f).numpy()


def point_predsample_inhomogeneous(tilde_l_hist, uL_vecs_hist, tilde_sigma2_err_hist, Y, x, x_star, mu_tilde_l, alpha_tilde_l,
                                   beta_tilde_l, mu_L, alpha_L, beta_L, N_sample, *args, **kwargs):
    """(prediction.py:1265-1357): tensor [N_sample, M]."""
    torch = _lib.require_cuda()
    res = pointwise_predsample_inhomogeneous(tilde_l_hist, uL_vecs_hist, tilde_sigma2_err_hist, Y, x,
                                             np.asarray(x_star, dtype=np.float64).reshape(1), mu_tilde_l, alpha_tilde_l,
                                             beta_tilde_l, mu_L, alpha_L, beta_L, N_sample, **kwargs)
    return torch.from_numpy(res[0])


def test_predsample_inhomogeneous(tilde_l_hist, uL_vecs_hist, tilde_sigma2_err_hist, Y, x, x_test, mu_tilde_l, alpha_tilde_l,
                                  beta_tilde_l, mu_L, alpha_L, beta_L, N_sample, *args, **kwargs):
    """(prediction.py:1380-1399)"""
    return pointwise_predsample_inhomogeneous(tilde_l_hist, uL_vecs_hist, tilde_sigma2_err_hist, Y, x, x_test, mu_tilde_l,
                                              alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, N_sample, **kwargs)


test_predsample_inhomogeneous.__test__ = False


# ===================================================================================== separable model (prediction.py:34-459)
def _B_diag(uL_vec, M):
    """diag(L L^T) for the unconstrained triangle(s) uL_vec [..., T]  (prediction.py:87-88)."""
    torch = _lib.require_cuda()
    uL = torch.as_tensor(uL_vec, dtype=torch.float64)
    Lv = utils.uLvec2Lvec(uL.reshape(-1, uL.shape[-1]), M)
    idx = torch.tril_indices(M, M)
    L = torch.zeros((Lv.shape[0], M, M), dtype=torch.float64)
    L[:, idx[0], idx[1]] = Lv
    return torch.diagonal(L @ L.transpose(1, 2), dim1=1, dim2=2).reshape(*uL.shape[:-1], M)


def _sep_hyper(mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, kwargs):
    return dict(mu_tilde_l=mu_tilde_l, alpha_tilde_l=alpha_tilde_l, beta_tilde_l=beta_tilde_l, mu_tilde_sigma=mu_tilde_sigma,
                alpha_tilde_sigma=alpha_tilde_sigma, beta_tilde_sigma=beta_tilde_sigma, a=kwargs.get("a", 1),
                b=kwargs.get("b", 1), c=kwargs.get("c", 10))


def _sep_plan_and_pars(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, hyper):
    torch = _lib.require_cuda()
    plan = logpos._get_plan("separable", Y, x, hyper, True)
    f64 = torch.float64
    pars = torch.cat([torch.as_tensor(tilde_l, dtype=f64).reshape(-1), torch.as_tensor(tilde_sigma, dtype=f64).reshape(-1),
                      torch.as_tensor(uL_vec, dtype=f64).reshape(-1),
                      torch.as_tensor(tilde_sigma2_err, dtype=f64).reshape(1)]).detach()
    return plan, pars


def _check_info(info):
    if int(info.abs().sum()) != 0:
        raise _lib.NmgpError("prediction: a covariance is not positive definite")


def pointwise_predmap(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, grids, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                      mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, *args, **kwargs):
    """Plug-in prediction of the separable model (prediction.py:337-430): tensor [G,3,M] = mu_f -/+ 1.96 sd."""
    torch = _lib.require_cuda()
    hyper = _sep_hyper(mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, kwargs)
    plan, pars = _sep_plan_and_pars(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, hyper)
    g = torch.as_tensor(grids, dtype=torch.float64).reshape(-1)
    mu_l, _, mu_s, _ = plan.predict_prior_moments(pars, g)
    mu_f, quad, info = plan.predict_moments_sep(pars, g, mu_l.unsqueeze(2), mu_s)
    _check_info(info)
    mu_f, quad = mu_f[0, :, 0].cpu(), quad[0, :, 0].cpu()
    s2e = torch.exp(torch.as_tensor(tilde_sigma2_err, dtype=torch.float64))
    sig_star = torch.exp(mu_s[0, :, 0].cpu())
    a2 = _B_diag(uL_vec, plan.M).reshape(1, -1) * (settings.jitter + sig_star * sig_star).unsqueeze(1)   # prediction.py:389-393
    s2y = a2 - quad + s2e
    s2y[s2y <= 0] = settings.precision
    sd = torch.sqrt(s2y)
    return torch.stack([mu_f - 1.96 * sd, mu_f, mu_f + 1.96 * sd], dim=1)


def point_predmap(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_star, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                  mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, *args, **kwargs):
    """(prediction.py:337-408): tensor [3,M]."""
    return pointwise_predmap(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x,
                             np.asarray(x_star, dtype=np.float64).reshape(1), mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                             mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, **kwargs)[0]


def test_predmap(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_test, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                 mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, *args, **kwargs):
    """(prediction.py:432-459)"""
    return pointwise_predmap(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_test, mu_tilde_l, alpha_tilde_l,
                             beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, **kwargs)


test_predmap.__test__ = False


def pointwise_predmap_sampling(n_sample, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, grids, mu_tilde_l,
                               alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, *args,
                               **kwargs):
    """Sampling prediction of the separable model from MAP estimates (prediction.py:189-306): per new input and sample
    tilde_l* (1 draw), tilde_sigma* (1), y (M); returns (quantiles [G,2,M], mean [G,M], std [G,M])."""
    torch = _lib.require_cuda()
    f64 = torch.float64
    hyper = _sep_hyper(mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, kwargs)
    plan, pars = _sep_plan_and_pars(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, hyper)
    M = plan.M
    g = torch.as_tensor(grids, dtype=f64).reshape(-1)
    G = int(g.numel())
    mu_l, s2_l, mu_s, s2_s = (t[0].cpu() for t in plan.predict_prior_moments(pars, g))
    z_l, z_s, z_y = torch.empty((G, n_sample), dtype=f64), torch.empty((G, n_sample), dtype=f64), torch.empty((G, n_sample, M), dtype=f64)
    _draw_interleaved(z_l, z_s, z_y)   # prediction.py:213 (tilde_l*), :223 (tilde_sigma*), :266 (y)
    tl_star = z_l.mul(torch.sqrt(s2_l).unsqueeze(1)).add(mu_l.unsqueeze(1))
    ts_star = z_s.mul(torch.sqrt(s2_s).unsqueeze(1)).add(mu_s[:, 0].unsqueeze(1))
    mu_f, quad, info = plan.predict_moments_sep(pars, g, tl_star.unsqueeze(0), ts_star.unsqueeze(0))
    _check_info(info)
    mu_f, quad = mu_f[0].cpu(), quad[0].cpu()
    s2e = torch.exp(torch.as_tensor(tilde_sigma2_err, dtype=f64))
    sig_star = torch.exp(ts_star)
    a2 = (sig_star ** 2).unsqueeze(2) * _B_diag(uL_vec, M).reshape(1, 1, -1)                            # prediction.py:254
    s2y = a2 - quad + s2e
    s2y[s2y <= 0] = settings.precision
    return _summaries(z_y.mul(torch.sqrt(s2y)).add(mu_f))


def point_predmap_sampling(n_sample, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_star, mu_tilde_l, alpha_tilde_l,
                           beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, *args, **kwargs):
    """(prediction.py:189-277): (quantiles [2,M], mean [M], std [M])."""
    q, m, sd = pointwise_predmap_sampling(n_sample, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x,
                                          np.asarray(x_star, dtype=np.float64).reshape(1), mu_tilde_l, alpha_tilde_l,
                                          beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, **kwargs)
    return q[0], m[0], sd[0]


def test_predmap_sampling(n_sample, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_test, mu_tilde_l, alpha_tilde_l,
                          beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, *args, **kwargs):
    """(prediction.py:308-335)"""
    return pointwise_predmap_sampling(n_sample, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, x_test, mu_tilde_l,
                                      alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, **kwargs)


test_predmap_sampling.__test__ = False


def _replicated_plan(model, H, Y, x, hyper):
    """One plan whose H 'subjects' are the same data under H different parameter vectors (a posterior history)."""
    torch = _lib.require_cuda()
    from .batched import LogPosteriorPlan
    Y = torch.as_tensor(Y, dtype=torch.float64)
    x = torch.as_tensor(x, dtype=torch.float64).reshape(-1)
    key = (model, H, logpos._fingerprint(Y), logpos._fingerprint(x), tuple(sorted((k, float(v)) for k, v in hyper.items())),
           torch.cuda.current_device())
    plan = _HIST_PLANS.get(key)
    if plan is None:
        for old in _HIST_PLANS.values():
            old.close()
        _HIST_PLANS.clear()
        plan = LogPosteriorPlan(model, x.unsqueeze(0).expand(H, -1), Y.unsqueeze(0).expand(H, -1, -1), hyper)
        _HIST_PLANS[key] = plan
    return plan


def pointwise_predsample(tilde_l_hist, tilde_sigma_hist, uL_vec_hist, tilde_sigma2_err_hist, Y, x, grids, mu_tilde_l,
                         alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, N_sample, *args,
                         **kwargs):
    """One predictive draw per entry of a posterior history of the separable model, at every new input
    (prediction.py:34-157): numpy array [G, N_sample, M].  The entries are the 'subjects' of one batched plan."""
    torch = _lib.require_cuda()
    f64 = torch.float64
    tl_h = torch.as_tensor(tilde_l_hist, dtype=f64)[-N_sample:]
    ts_h = torch.as_tensor(tilde_sigma_hist, dtype=f64)[-N_sample:]
    ul_h = torch.as_tensor(uL_vec_hist, dtype=f64)[-N_sample:]
    te_h = torch.as_tensor(tilde_sigma2_err_hist, dtype=f64).reshape(-1)[-N_sample:]
    H = int(tl_h.shape[0])
    hyper = _sep_hyper(mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, kwargs)
    plan = _replicated_plan("separable", H, Y, x, hyper)
    M = plan.M
    g = torch.as_tensor(grids, dtype=f64).reshape(-1)
    G = int(g.numel())
    pars = torch.cat([tl_h, ts_h, ul_h, te_h.unsqueeze(1)], dim=1).detach()
    mu_l, s2_l, mu_s, s2_s = (t.cpu() for t in plan.predict_prior_moments(pars, g))       # [H,G], [H,G], [H,G,1], [H,G]
    z_l, z_s, z_y = torch.empty((G, H), dtype=f64), torch.empty((G, H), dtype=f64), torch.empty((G, H, M), dtype=f64)
    _draw_interleaved(z_l, z_s, z_y)   # per new input, per history entry: tilde_l*, tilde_sigma*, y  (prediction.py:60, 70, 118)
    tl_star = z_l.t().mul(torch.sqrt(s2_l)).add(mu_l)                                     # [H,G]
    ts_star = z_s.t().mul(torch.sqrt(s2_s)).add(mu_s[:, :, 0])
    mu_f, quad, info = plan.predict_moments_sep(pars, g, tl_star.unsqueeze(2), ts_star.unsqueeze(2))
    _check_info(info)
    mu_f, quad = mu_f[:, :, 0].cpu(), quad[:, :, 0].cpu()                                 # [H,G,M]
    sig_star = torch.exp(ts_star)
    a2 = _B_diag(ul_h, M).unsqueeze(1) * (settings.jitter + sig_star * sig_star).unsqueeze(2)   # prediction.py:101
    s2y = a2 - quad + torch.exp(te_h).reshape(H, 1, 1)
    s2y[s2y <= 0] = settings.precision
    return z_y.mul(torch.sqrt(s2y).transpose(0, 1)).add(mu_f.transpose(0, 1)).numpy()


def point_predsample(tilde_l_hist, tilde_sigma_hist, uL_vec_hist, tilde_sigma2_err_hist, Y, x, x_star, mu_tilde_l,
                     alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, N_sample, *args,
                     **kwargs):
    """(prediction.py:34-131): tensor [N_sample, M]."""
    torch = _lib.require_cuda()
    res = pointwise_predsample(tilde_l_hist, tilde_sigma_hist, uL_vec_hist, tilde_sigma2_err_hist, Y, x,
                               np.asarray(x_star, dtype=np.float64).reshape(1), mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                               mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, N_sample, **kwargs)
    return torch.from_numpy(res[0])


def test_predsample(tilde_l_hist, tilde_sigma_hist, uL_vec_hist, tilde_sigma2_err_hist, Y, x, x_test, mu_tilde_l,
                    alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, N_sample, *args, **kwargs):
    """(prediction.py:159-187)"""
    return pointwise_predsample(tilde_l_hist, tilde_sigma_hist, uL_vec_hist, tilde_sigma2_err_hist, Y, x, x_test, mu_tilde_l,
                                alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, N_sample,
                                **kwargs)


test_predsample.__test__ = False


# ===================================================================================== stationary model (prediction.py:1566-1692)
_S_HYPER = dict(mu_tilde_l=0.0, sigma_tilde_l=1.0, a=1, b=1, c=10)   # priors do not enter the predictive moments


def _stationary_moments(tilde_ls, tilde_sigmas, uL_vecs, tilde_sigma2_errs, Y, x, grids):
    """mu_f, sigma2_y [H,G,M] for H stationary parameter sets (prediction.py:1587-1596), CPU tensors."""
    torch = _lib.require_cuda()
    f64 = torch.float64
    tl = torch.as_tensor(tilde_ls, dtype=f64).reshape(-1)
    ts = torch.as_tensor(tilde_sigmas, dtype=f64).reshape(-1)
    H = int(tl.numel())
    ul = torch.as_tensor(uL_vecs, dtype=f64).reshape(H, -1)
    te = torch.as_tensor(tilde_sigma2_errs, dtype=f64).reshape(-1)
    plan = logpos._get_plan("stationary", Y, x, _S_HYPER, True) if H == 1 else _replicated_plan("stationary", H, Y, x, _S_HYPER)
    g = torch.as_tensor(grids, dtype=f64).reshape(-1)
    G = int(g.numel())
    pars = torch.cat([tl.unsqueeze(1), ts.unsqueeze(1), ul, te.unsqueeze(1)], dim=1).detach()
    mu_f, quad, info = plan.predict_moments_sep(pars, g, tl.reshape(H, 1, 1).expand(H, G, 1), ts.reshape(H, 1, 1).expand(H, G, 1))
    _check_info(info)
    mu_f, quad = mu_f[:, :, 0].cpu(), quad[:, :, 0].cpu()
    sigma = torch.exp(ts)
    s2y = ((sigma ** 2).reshape(H, 1) * _B_diag(ul, plan.M)).unsqueeze(1) - quad + torch.exp(te).reshape(H, 1, 1)
    s2y[s2y < 0] = settings.precision                                                     # prediction.py:1595
    return mu_f, s2y


def pointwise_predmap_S(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, grids, *args, **kwargs):
    """(prediction.py:1566-1598): tensor [G,3,M] = mu_f -/+ 1.96 sd."""
    torch = _lib.require_cuda()
    mu_f, s2y = _stationary_moments(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, grids)
    sd = torch.sqrt(s2y[0])
    return torch.stack([mu_f[0] - 1.96 * sd, mu_f[0], mu_f[0] + 1.96 * sd], dim=1)


def test_predmap_S(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, test_x, *args, **kwargs):
    """(prediction.py:1601-1638): (mean [G,M], std [G,M]) tensors."""
    torch = _lib.require_cuda()
    mu_f, s2y = _stationary_moments(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, test_x)
    return mu_f[0], torch.sqrt(s2y[0])


test_predmap_S.__test__ = False


def pointwise_predsample_S(tilde_ls, tilde_sigmas, uL_vecs, tilde_sigma2_errs, Y, x, grids, *args, **kwargs):
    """One draw per posterior sample and new input, numpy's global generator, one scalar per (sample, input) shared by the M
    outputs (prediction.py:1640-1665): numpy array [H, G, M]."""
    torch = _lib.require_cuda()
    mu_f, s2y = _stationary_moments(tilde_ls, tilde_sigmas, uL_vecs, tilde_sigma2_errs, Y, x, grids)
    H, G, _ = mu_f.shape
    z = torch.from_numpy(np.array([[np.random.randn() for _ in range(G)] for _ in range(H)], dtype=np.float64).reshape(H, G, 1))
    return (mu_f + z * torch.sqrt(s2y)).numpy()


def test_predsample_S(tilde_ls, tilde_sigmas, uL_vecs, tilde_sigma2_errs, Y, x, test_x, *args, **kwargs):
    """(prediction.py:1667-1692)"""
    return pointwise_predsample_S(tilde_ls, tilde_sigmas, uL_vecs, tilde_sigma2_errs, Y, x, test_x)


test_predsample_S.__test__ = False
