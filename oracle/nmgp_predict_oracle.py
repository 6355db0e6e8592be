"""CPU oracle for posterior prediction of the nonseparable model.  TEST INFRASTRUCTURE ONLY (see nmgp_oracle.py: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import anything under `oracle/`).

Restates the reference's algorithm, `Utility/prediction.py:1038-1262`
(`point_/pointwise_/test_predmap_inhomogeneous_sampling`), in torch CPU float64 with the same dense operation
sequence per grid point and sample:

  conditional of the GP priors at x*:   proj = Sigma_p^-1 k_p (LU solve),  mu = mu_p + proj . (v - mu_p),
                                        sigma2 = k_p(x*,x*) - proj . k_p,  clipped to 1e-6 when negative  (:1060-1092)
  draws, in this order per sample:      tilde_l* (1), uL* (T), y (M)  from torch's global generator        (:1113-1169)
  predictive moments:                   Sigma^-1 by `symeig`, mu_f = k_f^T Sigma^-1 y,
                                        Sigma_f = A - (k_f^T chol(Sigma^-1)) (..)^T, sigma2_y = diag + sigma2_err,
                                        values <= 0 clipped to 1e-6                                         (:1130-1167)

The only liberty taken: Sigma^-1 and its Cholesky factor do not depend on the grid point or the sample, so they are formed
once instead of G * n_sample times (the same LAPACK calls on the same matrix return the same bits).

Parity pin: `tests/golden/predict_*.npz`, generated from the unmodified reference by `tests/golden/make_golden_predict.py`
(every draw's loc / scale / value recorded); `tests/test_oracle_golden.py` checks this file against all of them.
"""
from __future__ import annotations

import numpy as np
import torch

from . import nmgp_oracle as O

PRECISION = 1e-6  # Utility/settings.py:6


def rbf_cross(x1: torch.Tensor, x2: torch.Tensor, alpha: float, beta: float) -> torch.Tensor:
    """alpha^2 exp(-0.5 |(x1_i - x2_j)/beta|^2), no jitter (Utility/kernels.py:24-43 with X2 given)."""
    a, b = x1 / beta, x2 / beta
    d = ((a * a).view(-1, 1) + (b * b).view(1, -1)) - 2.0 * (a.view(-1, 1) * b.view(1, -1))
    return torch.exp(-0.5 * d) * alpha ** 2


def gibbs_cross(x1, ell1, x2, ell2) -> torch.Tensor:
    """sqrt(2 l1_i l2_j / (l1_i^2 + l2_j^2)) exp(-d_ij / (l1_i^2 + l2_j^2)), no jitter (Utility/kernels.py:46-73, X2 given)."""
    d = ((x1 * x1).view(-1, 1) + (x2 * x2).view(1, -1)) - 2.0 * (x1.view(-1, 1) * x2.view(1, -1))
    A = (ell1 ** 2).view(-1, 1) + (ell2 ** 2).view(1, -1)
    B = ell1.view(-1, 1) * ell2.view(1, -1)
    return torch.sqrt(2.0 * B / A) * torch.exp(-d / A)


def prior_conditional(x, x_star, v_cols, mu, alpha, beta):
    """Conditional mean(s) and variance of a GP prior at x_star given its values v_cols [N, C] at x
    (prediction.py:1060-1068 for tilde_l, :1070-1081 for the T columns of uL).  Returns (mu [C], sigma2 scalar)."""
    Sigma = O.rbf_cov(x, alpha, beta)
    k = rbf_cross(x, x_star.view(1), alpha, beta)                       # N x 1
    proj = torch.linalg.solve(Sigma, k).view(-1)                         # torch.solve(input=k, A=Sigma)
    mu_c = mu + (proj.view(-1, 1) * (v_cols - mu)).sum(0)
    s2 = (alpha ** 2 + O.JITTER) - torch.dot(proj, k.view(-1))           # RBF_cov(x*)[0,0] = jitter + alpha^2
    return mu_c, s2


def predictive_moments(x, ell, Lmats, y_om, invS, invL, s2e, x_star, tilde_l_star, L_star):
    """mu_f [M], sigma2_y [M] at one x_star for one sampled (tilde_l*, L*)  (prediction.py:1144-1165)."""
    N, M = Lmats.shape[0], Lmats.shape[1]
    order = torch.arange(N * M).view(N, M).t().reshape(-1)
    l_star = torch.exp(tilde_l_star).view(1)
    k_x = gibbs_cross(x, ell, x_star.view(1), l_star).view(-1)            # N
    A_f = (k_x.view(N, 1, 1) * Lmats).reshape(N * M, M)                   # rows (i,m): k_i * L_i
    k_f = (L_star @ A_f.t()).t()[order]                                   # NM x M, output-major rows
    mu_f = k_f.t() @ (invS @ y_om)
    Tm = k_f.t() @ invL
    A = (1.0 + O.JITTER) * (L_star @ L_star.t())                          # Nonstationary_RBF_cov(x*)[0,0] = 1 + jitter
    Sigma_y = A - Tm @ Tm.t() + s2e * torch.eye(M, dtype=O.DTYPE)
    s2y = torch.diagonal(Sigma_y).clone()
    s2y[s2y <= 0] = PRECISION
    return mu_f, s2y


def pointwise_predict(n_sample, tilde_l, uL_vecs, tilde_sigma2_err, Y, x, grids, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                      mu_L, alpha_L, beta_L, mode="y"):
    """All draws of `pointwise_predmap_inhomogeneous_sampling` (prediction.py:1194-1235) with torch's global generator
    consumed in the reference's order.  mode: "y" (default), "smoothness" (pred_smoothness=True), "cov" (pred_cov=True).
    Returns a dict of arrays [G, ns, ...]: loc / scale / draw of every Normal the reference samples in that mode."""
    N, M = Y.shape
    T = O.tril_size(M)
    G = grids.numel()
    y_om = Y.t().reshape(-1)
    ell = torch.exp(tilde_l)
    s2e = torch.exp(tilde_sigma2_err)
    Lmats = O.tril_vec_to_matrix(O.unconstrained_to_tril_vec(uL_vecs.view(N, T), M), M)
    out = {k: [] for k in ("l_loc", "l_scale", "l_draw", "u_loc", "u_scale", "u_draw", "y_loc", "y_scale", "y_draw")}
    if mode == "y":
        Sigma = O.nonseparable_cov(x, tilde_l, uL_vecs, M)
        w, V = torch.linalg.eigh(Sigma, UPLO="U")
        invS = (V @ torch.diag(1.0 / (s2e + w))) @ V.t()                  # prediction.py:1144-1146
        invL = torch.linalg.cholesky(invS)
    for g in range(G):
        xs = grids[g]
        if mode in ("y", "smoothness"):
            mu_l, s2_l = prior_conditional(x, xs, tilde_l.view(N, 1), mu_tilde_l, alpha_tilde_l, beta_tilde_l)
            mu_l = mu_l[0]
            if s2_l < 0:
                s2_l = torch.tensor(PRECISION, dtype=O.DTYPE)
        if mode in ("y", "cov"):
            mu_u, s2_u = prior_conditional(x, xs, uL_vecs.view(N, T), mu_L, alpha_L, beta_L)
            s2_u = s2_u.expand(T).clone()
            s2_u[s2_u < 0] = PRECISION
        for _ in range(n_sample):
            if mode in ("y", "smoothness"):
                sc = torch.sqrt(s2_l)
                tl_star = torch.empty((), dtype=O.DTYPE).normal_().mul(sc).add(mu_l)
                out["l_loc"].append(float(mu_l)); out["l_scale"].append(float(sc)); out["l_draw"].append(float(tl_star))
            if mode in ("y", "cov"):
                sc = torch.sqrt(s2_u)
                uL_star = torch.empty(T, dtype=O.DTYPE).normal_().mul(sc).add(mu_u)
                out["u_loc"].append(mu_u.numpy().copy()); out["u_scale"].append(sc.numpy().copy())
                out["u_draw"].append(uL_star.numpy().copy())
            if mode == "y":
                L_star = O.tril_vec_to_matrix(O.unconstrained_to_tril_vec(uL_star, M), M)
                mu_f, s2y = predictive_moments(x, ell, Lmats, y_om, invS, invL, s2e, xs, tl_star, L_star)
                sc = torch.sqrt(s2y)
                yd = torch.empty(M, dtype=O.DTYPE).normal_().mul(sc).add(mu_f)
                out["y_loc"].append(mu_f.numpy().copy()); out["y_scale"].append(sc.numpy().copy())
                out["y_draw"].append(yd.numpy().copy())
    res = {}
    for k, v in out.items():
        if v:
            a = np.asarray(v)
            res[k] = a.reshape(G, n_sample, *a.shape[1:])
    return res


def summarise(y_draw: np.ndarray):
    """quantiles [G,2,M], mean [G,M], std [G,M] as the reference forms them per grid point (prediction.py:1186-1190)."""
    q = np.stack([np.percentile(s, q=[2.5, 97.5], axis=0) for s in y_draw])
    return q, y_draw.mean(axis=1), y_draw.std(axis=1)


def pointwise_predict_plugin(tilde_l, uL_vecs, tilde_sigma2_err, Y, x, grids, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L,
                             alpha_L, beta_L):
    """`pointwise_predmap_inhomogeneous` (prediction.py:912-1012): conditional means plugged in, no draws.
    Returns (percentiles [G,3,M], L* vectors [G,T]) as numpy arrays."""
    N, M = Y.shape
    T = O.tril_size(M)
    y_om = Y.t().reshape(-1)
    ell, s2e = torch.exp(tilde_l), torch.exp(tilde_sigma2_err)
    Lmats = O.tril_vec_to_matrix(O.unconstrained_to_tril_vec(uL_vecs.view(N, T), M), M)
    w, V = torch.linalg.eigh(O.nonseparable_cov(x, tilde_l, uL_vecs, M), UPLO="U")
    invS = (V @ torch.diag(1.0 / (s2e + w))) @ V.t()
    invL = torch.linalg.cholesky(invS)
    pct, Lv = [], []
    for xs in grids:
        mu_l, _ = prior_conditional(x, xs, tilde_l.view(N, 1), mu_tilde_l, alpha_tilde_l, beta_tilde_l)
        mu_u, _ = prior_conditional(x, xs, uL_vecs.view(N, T), mu_L, alpha_L, beta_L)
        Lvec = O.unconstrained_to_tril_vec(mu_u, M)
        mu_f, s2y = predictive_moments(x, ell, Lmats, y_om, invS, invL, s2e, xs, mu_l[0], O.tril_vec_to_matrix(Lvec, M))
        sd = torch.sqrt(s2y)
        pct.append(torch.stack([mu_f - 1.96 * sd, mu_f, mu_f + 1.96 * sd]).numpy())
        Lv.append(Lvec.numpy())
    return np.stack(pct), np.stack(Lv)


def pointwise_predict_history(tilde_l_hist, uL_vecs_hist, tilde_sigma2_err_hist, Y, x, grids, mu_tilde_l, alpha_tilde_l,
                              beta_tilde_l, mu_L, alpha_L, beta_L, N_sample):
    """`pointwise_predsample_inhomogeneous` (prediction.py:1265-1378): per new input, per history entry (its own covariance):
    tilde_l* ~ conditional of the tilde_l prior, L* ~ conditional taken on the CONSTRAINED triangles (prediction.py:1302) and
    used as the factor itself (:1311), then y.  Returns a dict of [G, H, ...] arrays like `pointwise_predict`."""
    N, M = Y.shape
    T = O.tril_size(M)
    y_om = Y.t().reshape(-1)
    tl_h, ul_h, ts_h = tilde_l_hist[-N_sample:], uL_vecs_hist[-N_sample:], tilde_sigma2_err_hist[-N_sample:]
    H, G = tl_h.shape[0], grids.numel()
    pre = []
    for h in range(H):
        s2e = torch.exp(ts_h[h])
        w, V = torch.linalg.eigh(O.nonseparable_cov(x, tl_h[h], ul_h[h], M), UPLO="U")
        invS = (V @ torch.diag(1.0 / (s2e + w))) @ V.t()
        Lvecs = O.unconstrained_to_tril_vec(ul_h[h].view(N, T), M)
        pre.append((invS, torch.linalg.cholesky(invS), s2e, Lvecs, O.tril_vec_to_matrix(Lvecs, M)))
    out = {k: [] for k in ("l_loc", "l_scale", "l_draw", "u_loc", "u_scale", "u_draw", "y_loc", "y_scale", "y_draw")}
    for g in range(G):
        xs = grids[g]
        for h in range(H):
            invS, invL, s2e, Lvecs, Lmats = pre[h]
            mu_l, s2_l = prior_conditional(x, xs, tl_h[h].view(N, 1), mu_tilde_l, alpha_tilde_l, beta_tilde_l)
            if s2_l < 0:
                s2_l = torch.tensor(PRECISION, dtype=O.DTYPE)
            sc_l = torch.sqrt(s2_l)
            tl_star = torch.empty((), dtype=O.DTYPE).normal_().mul(sc_l).add(mu_l[0])
            mu_u, s2_u = prior_conditional(x, xs, Lvecs, mu_L, alpha_L, beta_L)
            s2_u = s2_u.expand(T).clone()
            s2_u[s2_u < 0] = PRECISION
            sc_u = torch.sqrt(s2_u)
            L_star_vec = torch.empty(T, dtype=O.DTYPE).normal_().mul(sc_u).add(mu_u)
            mu_f, s2y = predictive_moments(x, torch.exp(tl_h[h]), Lmats, y_om, invS, invL, s2e, xs, tl_star,
                                           O.tril_vec_to_matrix(L_star_vec, M))
            sc_y = torch.sqrt(s2y)
            yd = torch.empty(M, dtype=O.DTYPE).normal_().mul(sc_y).add(mu_f)
            for k, v in (("l_loc", mu_l[0]), ("l_scale", sc_l), ("l_draw", tl_star)):
                out[k].append(float(v))
            for k, v in (("u_loc", mu_u), ("u_scale", sc_u), ("u_draw", L_star_vec), ("y_loc", mu_f), ("y_scale", sc_y),
                         ("y_draw", yd)):
                out[k].append(v.numpy().copy())
    res = {}
    for k, v in out.items():
        a = np.asarray(v)
        res[k] = a.reshape(G, H, *a.shape[1:])
    return res


# ===================================================================================== separable / stationary models
def gibbs_cross_sigma(x1, sig1, ell1, x2, sig2, ell2) -> torch.Tensor:
    """sig1_i sig2_j sqrt(2 l1_i l2_j/(l1_i^2+l2_j^2)) exp(-d_ij/(l1_i^2+l2_j^2))  (Utility/kernels.py:46-73, X2 given)."""
    return (sig1.view(-1, 1) * sig2.view(1, -1)) * gibbs_cross(x1, ell1, x2, ell2)


def _sep_setup(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x):
    N, M = Y.shape
    L = O.tril_vec_to_matrix(O.unconstrained_to_tril_vec(uL_vec, M), M)
    B_f = L @ L.t()
    ell, sig, s2e = torch.exp(tilde_l), torch.exp(tilde_sigma), torch.exp(tilde_sigma2_err)
    K_x = O.gibbs_cov(x, ell, sig)
    wB, vB = torch.linalg.eigh(B_f, UPLO="U")                                    # prediction.py:95-96
    wK, vK = torch.linalg.eigh(K_x, UPLO="U")
    Kt = torch.kron(vB.t(), vK.t())                                              # what kron_mv(v_B.t(), v_K.t(), .) applies
    w = 1.0 / (s2e + torch.kron(wB, wK))                                         # kronecker_product_diag, :99-100
    b = Kt @ Y.t().reshape(-1)
    return dict(B_f=B_f, ell=ell, sig=sig, s2e=s2e, Kt=Kt, w=w, b=b, M=M, N=N)


def sep_predictive_moments(su, x, x_star, tilde_l_star, tilde_sigma_star, jitter_in_a2):
    """mu_f [M], sigma2_y [M] of the separable model at one x_star (prediction.py:97-116 / 241-261 / 381-398).
    jitter_in_a2: a2 = diag(B_f (x) Nonstationary_RBF_cov(x*)) = diag(B_f) (1e-6 + sigma*^2)  (:101, :389) versus
    a2 = sigma*^2 diag(B_f) (:254)."""
    l_star, sig_star = torch.exp(tilde_l_star).view(1), torch.exp(tilde_sigma_star).view(1)
    k_x = gibbs_cross_sigma(x, su["sig"], su["ell"], x_star.view(1), sig_star, l_star)      # N x 1
    k_f = torch.kron(su["B_f"], k_x)                                                          # NM x M
    A = (su["Kt"] @ k_f).t()                                                                  # M x NM
    mu_f = A @ (su["b"] * su["w"])
    var_star = (O.JITTER + sig_star * sig_star) if jitter_in_a2 else sig_star ** 2
    a2 = torch.diagonal(su["B_f"]) * var_star
    s2y = a2 - (A * su["w"] * A).sum(dim=1) + su["s2e"]
    s2y = s2y.clone()
    s2y[s2y <= 0] = PRECISION
    return mu_f, s2y


def _clip_neg(s2):
    return s2 if s2 >= 0 else torch.tensor(PRECISION, dtype=O.DTYPE)


def sep_pointwise_plugin(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, grids, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                         mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma):
    """`pointwise_predmap` (prediction.py:337-430): numpy [G,3,M]."""
    su = _sep_setup(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x)
    N = su["N"]
    out = []
    for xs in grids:
        mu_l, _ = prior_conditional(x, xs, tilde_l.view(N, 1), mu_tilde_l, alpha_tilde_l, beta_tilde_l)
        mu_s, _ = prior_conditional(x, xs, tilde_sigma.view(N, 1), mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma)
        mu_f, s2y = sep_predictive_moments(su, x, xs, mu_l[0], mu_s[0], True)
        sd = torch.sqrt(s2y)
        out.append(torch.stack([mu_f - 1.96 * sd, mu_f, mu_f + 1.96 * sd]).numpy())
    return np.stack(out)


def sep_pointwise_predict(n_sample, tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, grids, mu_tilde_l, alpha_tilde_l,
                          beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma):
    """`pointwise_predmap_sampling` (prediction.py:189-306): all draws [G,ns,...], generator consumed in the reference's order
    (per new input, per sample: tilde_l*, tilde_sigma*, y)."""
    su = _sep_setup(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x)
    N, M = su["N"], su["M"]
    keys = ("l_loc", "l_scale", "l_draw", "s_loc", "s_scale", "s_draw", "y_loc", "y_scale", "y_draw")
    out = {k: [] for k in keys}
    for xs in grids:
        mu_l, s2_l = prior_conditional(x, xs, tilde_l.view(N, 1), mu_tilde_l, alpha_tilde_l, beta_tilde_l)
        mu_s, s2_s = prior_conditional(x, xs, tilde_sigma.view(N, 1), mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma)
        sc_l, sc_s = torch.sqrt(_clip_neg(s2_l)), torch.sqrt(_clip_neg(s2_s))
        for _ in range(n_sample):
            tl_star = torch.empty((), dtype=O.DTYPE).normal_().mul(sc_l).add(mu_l[0])
            ts_star = torch.empty((), dtype=O.DTYPE).normal_().mul(sc_s).add(mu_s[0])
            mu_f, s2y = sep_predictive_moments(su, x, xs, tl_star, ts_star, False)
            sc_y = torch.sqrt(s2y)
            yd = torch.empty(M, dtype=O.DTYPE).normal_().mul(sc_y).add(mu_f)
            for k, v in zip(keys, (mu_l[0], sc_l, tl_star, mu_s[0], sc_s, ts_star)):
                out[k].append(float(v))
            for k, v in zip(keys[6:], (mu_f, sc_y, yd)):
                out[k].append(v.numpy().copy())
    G = grids.numel()
    return {k: np.asarray(v).reshape(G, n_sample, *np.asarray(v).shape[1:]) for k, v in out.items()}


def sep_pointwise_history(tilde_l_hist, tilde_sigma_hist, uL_vec_hist, tilde_sigma2_err_hist, Y, x, grids, mu_tilde_l,
                          alpha_tilde_l, beta_tilde_l, mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, N_sample):
    """`pointwise_predsample` (prediction.py:34-157): per new input, per history entry: tilde_l*, tilde_sigma*, y."""
    tl_h, ts_h, ul_h, te_h = (t[-N_sample:] for t in (tilde_l_hist, tilde_sigma_hist, uL_vec_hist, tilde_sigma2_err_hist))
    H, G = tl_h.shape[0], grids.numel()
    sus = [_sep_setup(tl_h[h], ts_h[h], ul_h[h], te_h[h], Y, x) for h in range(H)]
    N, M = sus[0]["N"], sus[0]["M"]
    keys = ("l_loc", "l_scale", "l_draw", "s_loc", "s_scale", "s_draw", "y_loc", "y_scale", "y_draw")
    out = {k: [] for k in keys}
    for xs in grids:
        for h in range(H):
            mu_l, s2_l = prior_conditional(x, xs, tl_h[h].view(N, 1), mu_tilde_l, alpha_tilde_l, beta_tilde_l)
            sc_l = torch.sqrt(_clip_neg(s2_l))
            tl_star = torch.empty((), dtype=O.DTYPE).normal_().mul(sc_l).add(mu_l[0])
            mu_s, s2_s = prior_conditional(x, xs, ts_h[h].view(N, 1), mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma)
            sc_s = torch.sqrt(_clip_neg(s2_s))
            ts_star = torch.empty((), dtype=O.DTYPE).normal_().mul(sc_s).add(mu_s[0])
            mu_f, s2y = sep_predictive_moments(sus[h], x, xs, tl_star, ts_star, True)
            sc_y = torch.sqrt(s2y)
            yd = torch.empty(M, dtype=O.DTYPE).normal_().mul(sc_y).add(mu_f)
            for k, v in zip(keys, (mu_l[0], sc_l, tl_star, mu_s[0], sc_s, ts_star)):
                out[k].append(float(v))
            for k, v in zip(keys[6:], (mu_f, sc_y, yd)):
                out[k].append(v.numpy().copy())
    return {k: np.asarray(v).reshape(G, H, *np.asarray(v).shape[1:]) for k, v in out.items()}


def stationary_moments(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, grids):
    """mu_f, sigma2_y [G,M] of the stationary model with the dense inverse of B_f (x) K_x + sigma2 I (prediction.py:1578-1596)."""
    N, M = Y.shape
    L = O.tril_vec_to_matrix(O.unconstrained_to_tril_vec(uL_vec, M), M)
    B_f = L @ L.t()
    l, sigma, s2e = torch.exp(tilde_l), torch.exp(tilde_sigma), torch.exp(tilde_sigma2_err)
    K_x = O.rbf_cov(x, float(sigma), float(l))
    invS = torch.inverse(torch.kron(B_f, K_x) + s2e * torch.eye(N * M, dtype=O.DTYPE))
    y = Y.t().reshape(-1)
    mus, s2s = [], []
    for xs in grids:
        k_x = rbf_cross(x, xs.view(1), float(sigma), float(l))
        k_f = torch.kron(B_f, k_x).t()
        mu_f = (k_f @ invS) @ y
        s2y = sigma ** 2 * torch.diagonal(B_f) - torch.diagonal((k_f @ invS) @ k_f.t()) + s2e
        s2y = s2y.clone()
        s2y[s2y < 0] = PRECISION
        mus.append(mu_f.numpy()); s2s.append(s2y.numpy())
    return np.stack(mus), np.stack(s2s)
