"""CPU oracle for the NMGP log-posterior + gradient hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  The product
package (`nonstationary_multivariate_gaussian_process_b200`) never does and fails
loudly when its CUDA library is missing.

It is a restatement (vectorised, no per-time-point Python loops) of the reference's
*algorithm* for the three objectives, in torch CPU float64, differentiated by
`torch.autograd` exactly like the reference drivers do (`NegLog.backward()`,
e.g. Separable_Model/Separable_model.py:165):

  * nonseparable  -- Utility/logpos.py:299-380 (`nlogpos_obj_SVC` / `logpos_SVC`):
                     dense NM x NM covariance, explicit `inverse` + `logdet`.
  * separable     -- Utility/logpos.py:216-296 (`nlogpos_obj` / `logpos`):
                     Kronecker eigen path of Utility/distributions.py:26-52.
  * stationary    -- Utility/logpos.py:383-462 (`nlogpos_obj_S` / `logpos_S`).

Parity pin: the reference ships no tests or golden vectors, so the
pin is `tests/golden/*.npz`, produced by importing the unmodified reference from
/root/reference in the build container (`tests/golden/make_golden.py`, committed);
`tests/test_oracle_golden.py` checks this oracle against every one of them.
"""
from __future__ import annotations

import math

import torch

JITTER = 1e-6  # Utility/settings.py:3
DTYPE = torch.float64  # Utility/settings.py:4  (DoubleTensor)

_LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------- transforms
def tril_size(M: int) -> int:
    return M * (M + 1) // 2


def diag_slots(M: int) -> torch.Tensor:
    """Positions of the diagonal inside a row-major lower-triangle vector
    (Utility/utils.py:12: cumsum(1..M)-1)."""
    m = torch.arange(1, M + 1)
    return torch.cumsum(m, 0) - 1


def unconstrained_to_tril_vec(uL: torch.Tensor, M: int) -> torch.Tensor:
    """exp() on the diagonal slots, identity elsewhere; works on [..., T]
    (Utility/utils.py:10-22 `uLvec2Lvec`, :38-46 `uLvecs2Lvecs`)."""
    mask = torch.zeros(tril_size(M), dtype=torch.bool)
    mask[diag_slots(M)] = True
    return torch.where(mask, torch.exp(uL), uL)


def tril_vec_to_matrix(Lvec: torch.Tensor, M: int) -> torch.Tensor:
    """Scatter [..., T] row-major lower-triangle vectors into [..., M, M]
    (Utility/utils.py:56-74 `vec2lowtriangle`)."""
    r, c = torch.tril_indices(M, M)
    out = torch.zeros(*Lvec.shape[:-1], M, M, dtype=Lvec.dtype)
    out[..., r, c] = Lvec
    return out


# ----------------------------------------------------------------------------- kernels
def sq_dist(x: torch.Tensor) -> torch.Tensor:
    """|x_i|^2 + |x_j|^2 - 2 x_i x_j, in that order (Utility/kernels.py:13-20)."""
    x2 = x * x
    return (x2.view(-1, 1) + x2.view(1, -1)) - 2.0 * (x.view(-1, 1) * x.view(1, -1))


def rbf_cov(x: torch.Tensor, alpha: float, beta: float) -> torch.Tensor:
    """alpha^2 exp(-0.5 |(x_i-x_j)/beta|^2) + jitter I  (Utility/kernels.py:24-43):
    the inputs are divided by beta first, the jitter identity is the accumulator."""
    xs = x / beta
    cov = torch.eye(x.numel(), dtype=DTYPE) * JITTER
    cov = cov + torch.exp(-0.5 * sq_dist(xs)) * alpha ** 2
    return cov


def gibbs_cov(x: torch.Tensor, ell: torch.Tensor, sigma: torch.Tensor | None = None) -> torch.Tensor:
    """sigma_i sigma_j sqrt(2 l_i l_j / (l_i^2+l_j^2)) exp(-d_ij/(l_i^2+l_j^2)) + jitter I
    (Utility/kernels.py:46-73, self-covariance branch)."""
    N = x.numel()
    if sigma is None:
        sigma = torch.ones(N, dtype=DTYPE)
    A = (ell ** 2).view(-1, 1) + (ell ** 2).view(1, -1)
    B = ell.view(-1, 1) * ell.view(1, -1)
    C = sigma.view(-1, 1) * sigma.view(1, -1)
    cov = torch.eye(N, dtype=DTYPE) * JITTER
    return cov + C * torch.sqrt(2.0 * B / A) * torch.exp(-sq_dist(x) / A)


# ----------------------------------------------------------------------------- densities
def mvn_logpdf_full(v: torch.Tensor, mean: float, cov: torch.Tensor) -> torch.Tensor:
    """Normalised MVN log-density through a Cholesky factor and a triangular solve, which is
    what torch.distributions.MultivariateNormal.log_prob does (call sites
    Utility/logpos.py:274,279,358,365)."""
    Lc = torch.linalg.cholesky(cov)
    r = (v - mean).unsqueeze(-1)
    z = torch.linalg.solve_triangular(Lc, r, upper=False).squeeze(-1)
    half_logdet = torch.log(torch.diagonal(Lc)).sum()
    return -0.5 * (v.shape[-1] * _LOG_2PI + (z * z).sum(-1)) - half_logdet


def normal_logpdf(v: torch.Tensor, mean: float, std: float) -> torch.Tensor:
    """torch.distributions.Normal(mean, std).log_prob (Utility/logpos.py:283,446,450).
    The reference passes Python numbers, which torch turns into *default-dtype (float32)* tensors:
    std is rounded to float32, var = std32*std32 and log(std32) are evaluated in float32, then
    promoted.  That rounding (about 1e-8 relative in the term) is part of the reference's result, so the
    oracle calls the same third-party routine instead of restating it in float64."""
    return torch.distributions.Normal(mean, std).log_prob(v)


def inverse_gamma_logpdf(s2: torch.Tensor, a: float, b: float) -> torch.Tensor:
    """(-a-1) log s2 - b/s2 + a log b - lgamma(a)  (Utility/distributions.py:126-134)."""
    return (-a - 1.0) * torch.log(s2) - b / s2 + a * math.log(b) - math.lgamma(a)


def kron_eig_loglik(y_om: torch.Tensor, B: torch.Tensor, K: torch.Tensor, s2: torch.Tensor) -> torch.Tensor:
    """Un-normalised log N(y; 0, B (x) K + s2 I) by diagonalising both factors
    (Utility/distributions.py:26-52; the mat-vec is Utility/kronecker_operation.py:72-85).
    y_om is output-major: y[m*N+n] = Y[n,m] (Utility/logpos.py:250)."""
    wB, vB = torch.linalg.eigh(B, UPLO="U")
    wK, vK = torch.linalg.eigh(K, UPLO="U")
    M, N = B.shape[0], K.shape[0]
    Ymat = y_om.view(M, N).t()                      # N x M
    a = (vK.t() @ Ymat @ vB).t().reshape(-1)        # (vB^T (x) vK^T) y, output-major
    t = (wB.view(-1, 1) * wK.view(1, -1)).reshape(-1)
    return -0.5 * torch.log(t + s2).sum() - 0.5 * torch.dot(a / (s2 + t), a)


def dense_loglik(y: torch.Tensor, Sigma: torch.Tensor) -> torch.Tensor:
    """-0.5 logdet - 0.5 y^T inv(Sigma) y with an explicit inverse, as the nonseparable
    objective does (Utility/logpos.py:352-354 -> Utility/distributions.py:10-23)."""
    inv = torch.inverse(Sigma)
    return -0.5 * torch.logdet(Sigma) - 0.5 * torch.dot(y, inv @ y)


# ----------------------------------------------------------------------------- covariance assembly
def nonseparable_cov(x: torch.Tensor, tilde_l: torch.Tensor, uL: torch.Tensor, M: int) -> torch.Tensor:
    """K of Utility/logpos.py:339-349 in the reference's output-major ordering
    (row (m,i) -> m*N+i), *without* the noise term:
        K[(m,i),(m',j)] = K_x[i,j] * (L_i L_j^T)[m,m'].
    """
    N = x.numel()
    Lmats = tril_vec_to_matrix(unconstrained_to_tril_vec(uL.view(N, -1), M), M)   # N,M,M
    Kx = gibbs_cov(x, torch.exp(tilde_l))
    Lstack = Lmats.reshape(N * M, M)                # time-major rows (i,m)
    Ki = Lstack @ Lstack.t()                        # logpos.py:111-118
    order = torch.arange(N * M).view(N, M).t().reshape(-1)
    Ki = Ki[:, order][order]                        # logpos.py:347-348
    return Kx.repeat(M, M) * Ki                     # ones(M,M) (x) K_x, Hadamard with K_i


# ----------------------------------------------------------------------------- objectives
def logpost_nonseparable(pars, Y, x, mu_tilde_l=0.0, alpha_tilde_l=5.0, beta_tilde_l=1.0, mu_L=0.0,
                         alpha_L=5.0, beta_L=1.0, a=1, b=1, Prior=True):
    """Returns (logpost, loglik, lp_tilde_l, lp_uL, lp_sigma2)  -- Utility/logpos.py:326-380."""
    N, M = Y.shape
    T = tril_size(M)
    tilde_l, uL, tilde_s2 = pars[:N], pars[N:N + N * T], pars[-1]      # logpos.py:32-43
    y = Y.t().reshape(-1)
    s2 = torch.exp(tilde_s2)
    Sigma = nonseparable_cov(x, tilde_l, uL, M) + s2 * torch.eye(N * M, dtype=DTYPE)
    loglik = dense_loglik(y, Sigma)
    lp_l = mvn_logpdf_full(tilde_l, mu_tilde_l, rbf_cov(x, alpha_tilde_l, beta_tilde_l))
    cov_L = rbf_cov(x, alpha_L, beta_L)
    lp_uL = mvn_logpdf_full(uL.view(N, T).t(), mu_L, cov_L).sum()     # T columns, logpos.py:362-365
    lp_s2 = inverse_gamma_logpdf(s2, a, b)
    res = loglik
    if Prior:
        res = res + lp_l + lp_uL + lp_s2 + tilde_s2                    # Jacobian logpos.py:376
    return res, loglik, lp_l, lp_uL, lp_s2


def _separable_core(tilde_l_vec, tilde_sigma_vec, uL, tilde_s2, Y, x):
    N, M = Y.shape
    L = tril_vec_to_matrix(unconstrained_to_tril_vec(uL, M), M)
    B = L @ L.t()
    Kx = gibbs_cov(x, torch.exp(tilde_l_vec), torch.exp(tilde_sigma_vec))
    y = Y.t().reshape(-1)
    return kron_eig_loglik(y, B, Kx, torch.exp(tilde_s2))


def logpost_separable(pars, Y, x, mu_tilde_l=0.0, alpha_tilde_l=1.0, beta_tilde_l=1.0, mu_tilde_sigma=0.0,
                      alpha_tilde_sigma=1.0, beta_tilde_sigma=1.0, a=1, b=1, c=10, Prior=True):
    """Returns (logpost, loglik, lp_tilde_l, lp_tilde_sigma, lp_uL, lp_sigma2) -- logpos.py:237-296.
    (The reference's NaN retry with unseeded random jitter, logpos.py:267-268, is not restated:
    it is non-reproducible and is never a parity target; SURVEY.md appendix.)"""
    N, M = Y.shape
    T = tril_size(M)
    tilde_l, tilde_sigma, uL, tilde_s2 = pars[:N], pars[N:2 * N], pars[2 * N:2 * N + T], pars[-1]
    loglik = _separable_core(tilde_l, tilde_sigma, uL, tilde_s2, Y, x)
    lp_l = mvn_logpdf_full(tilde_l, mu_tilde_l, rbf_cov(x, alpha_tilde_l, beta_tilde_l))
    lp_sig = mvn_logpdf_full(tilde_sigma, mu_tilde_sigma, rbf_cov(x, alpha_tilde_sigma, beta_tilde_sigma))
    lp_uL = normal_logpdf(uL, 0.0, c).sum()
    s2 = torch.exp(tilde_s2)
    lp_s2 = inverse_gamma_logpdf(s2, a, b)
    res = loglik
    if Prior:
        res = res + lp_l + lp_sig + lp_uL + lp_s2 + tilde_s2
    return res, loglik, lp_l, lp_sig, lp_uL, lp_s2


def logpost_stationary(pars, Y, x, mu_tilde_l, sigma_tilde_l, a=1, b=1, c=10, Prior=True):
    """Returns (logpost, loglik, lp_tilde_l, lp_uL, lp_sigma2) -- logpos.py:405-462.
    pars = [tilde_l, tilde_sigma, uL(T), tilde_sigma2_err] (logpos.py:46-57)."""
    N, M = Y.shape
    T = tril_size(M)
    tilde_l, tilde_sigma, uL, tilde_s2 = pars[0], pars[1], pars[2:2 + T], pars[-1]
    ones = torch.ones(N, dtype=DTYPE)
    loglik = _separable_core(tilde_l * ones, tilde_sigma * ones, uL, tilde_s2, Y, x)
    lp_l = normal_logpdf(tilde_l, mu_tilde_l, sigma_tilde_l)
    lp_uL = normal_logpdf(uL, 0.0, c).sum()
    s2 = torch.exp(tilde_s2)
    lp_s2 = inverse_gamma_logpdf(s2, a, b)
    res = loglik
    if Prior:
        res = res + lp_l + lp_uL + lp_s2 + tilde_s2
    return res, loglik, lp_l, lp_uL, lp_s2


# ----------------------------------------------------------------------------- irregular sampling ("Hadamard")
def inverse_gamma_logpdf_u(s2: torch.Tensor, a: float, b: float) -> torch.Tensor:
    """(-a-1) log s2 - b/s2, WITHOUT the normaliser: what the Hadamard objectives use (Utility/distributions.py:116-124,
    call sites logpos.py:555, 650, 708)."""
    return (-a - 1.0) * torch.log(s2) - b / s2


def _hadamard_loglik(y, Kx, Ki, s2):
    """-0.5 logdet - 0.5 y^T inv(Kx o Ki + s2 I) y with explicit inverse + logdet (logpos.py:524-527, 590-593, 690-693)."""
    Sigma = Kx * Ki + s2 * torch.eye(y.numel(), dtype=DTYPE)
    return dense_loglik(y, Sigma)


def logpost_hadamard(pars, y, x, indx, mu_tilde_l=0.0, alpha_tilde_l=1.0, beta_tilde_l=1.0, mu_tilde_sigma=0.0,
                     alpha_tilde_sigma=1.0, beta_tilde_sigma=1.0, a=1, b=1, c=10, Prior=True):
    """(logpost, loglik, lp_tilde_l, lp_tilde_sigma, lp_L, lp_sigma2) -- logpos.py:503-558.  pars = [tilde_l(N),
    tilde_sigma(N), L_vec(T) RAW, tilde_sigma2_err]; one observation per row, indx its output."""
    N = y.numel()
    M = int(torch.unique(indx).numel())
    T = tril_size(M)
    tilde_l, tilde_sigma, L_vec, tilde_s2 = pars[:N], pars[N:2 * N], pars[2 * N:2 * N + T], pars[-1]
    L = tril_vec_to_matrix(L_vec, M)                                   # vec2lowtriangle(L_vec, M): no exp, logpos.py:518
    B = L @ L.t()
    Kx = gibbs_cov(x, torch.exp(tilde_l), torch.exp(tilde_sigma))
    s2 = torch.exp(tilde_s2)
    loglik = _hadamard_loglik(y, Kx, B[indx][:, indx], s2)             # generate_K_index, logpos.py:89-99
    lp_l = mvn_logpdf_full(tilde_l, mu_tilde_l, rbf_cov(x, alpha_tilde_l, beta_tilde_l))
    lp_sig = mvn_logpdf_full(tilde_sigma, mu_tilde_sigma, rbf_cov(x, alpha_tilde_sigma, beta_tilde_sigma))
    lp_L = normal_logpdf(L_vec, 0.0, c).sum()
    lp_s2 = inverse_gamma_logpdf_u(s2, a, b)
    res = loglik
    if Prior:
        res = res + lp_l + lp_sig + lp_L + lp_s2 + tilde_s2
    return res, loglik, lp_l, lp_sig, lp_L, lp_s2


def logpost_hadamard_svc(pars, y, x, indx, mu_tilde_l=0.0, alpha_tilde_l=1.0, beta_tilde_l=1.0, mu_L=0.0, alpha_L=1.0,
                         beta_L=1.0, a=1, b=1, Prior=True):
    """(logpost, loglik, lp_tilde_l, lp_L, lp_sigma2) -- logpos.py:582-637.  pars = [tilde_l(N), L_vecs(N*T) RAW,
    tilde_sigma2_err] (logpos.py:60-72)."""
    N = y.numel()
    M = int(torch.unique(indx).numel())
    T = tril_size(M)
    tilde_l, L_vecs, tilde_s2 = pars[:N], pars[N:N + N * T], pars[-1]
    Lm = tril_vec_to_matrix(L_vecs.view(N, T), M)                      # N,M,M, raw entries (logpos.py:582-583)
    Rows = Lm[torch.arange(N), indx]                                   # row indx_n of L_n  (logpos.py:114-116)
    Kx = gibbs_cov(x, torch.exp(tilde_l))
    s2 = torch.exp(tilde_s2)
    loglik = _hadamard_loglik(y, Kx, Rows @ Rows.t(), s2)
    lp_l = mvn_logpdf_full(tilde_l, mu_tilde_l, rbf_cov(x, alpha_tilde_l, beta_tilde_l))
    lp_L = mvn_logpdf_full(L_vecs.view(N, T).t(), mu_L, rbf_cov(x, alpha_L, beta_L)).sum()     # T columns, logpos.py:615-617
    lp_s2 = inverse_gamma_logpdf_u(s2, a, b)
    res = loglik
    if Prior:
        res = res + lp_l + lp_L + lp_s2 + tilde_s2
    return res, loglik, lp_l, lp_L, lp_s2


def logpost_hadamard_s(pars, y, x, indx, mu_tilde_l, sigma_tilde_l, a=1, b=1, c=10, Prior=True):
    """(logpost, loglik, lp_tilde_l, lp_L, lp_sigma2) -- logpos.py:655-716.  pars = [tilde_l, tilde_sigma, L_vec(T) RAW,
    tilde_sigma2_err]; K_x = RBF_cov(x, alpha=sigma, beta=l) (logpos.py:686)."""
    M = int(torch.unique(indx).numel())
    T = tril_size(M)
    tilde_l, tilde_sigma, L_vec, tilde_s2 = pars[0], pars[1], pars[2:2 + T], pars[-1]
    L = tril_vec_to_matrix(L_vec, M)
    B = L @ L.t()
    l, sigma = torch.exp(tilde_l), torch.exp(tilde_sigma)
    xs = x / l
    Kx = torch.eye(x.numel(), dtype=DTYPE) * JITTER + torch.exp(-0.5 * sq_dist(xs)) * sigma ** 2      # kernels.py:24-43
    s2 = torch.exp(tilde_s2)
    loglik = _hadamard_loglik(y, Kx, B[indx][:, indx], s2)
    lp_l = normal_logpdf(tilde_l, mu_tilde_l, sigma_tilde_l)
    lp_L = normal_logpdf(L_vec, 0.0, c).sum()
    lp_s2 = inverse_gamma_logpdf_u(s2, a, b)
    res = loglik
    if Prior:
        res = res + lp_l + lp_L + lp_s2 + tilde_s2
    return res, loglik, lp_l, lp_L, lp_s2


_HADAMARD = {"hadamard": logpost_hadamard, "hadamard_svc": logpost_hadamard_svc, "hadamard_s": logpost_hadamard_s}


def value_and_grad_hadamard(model: str, pars, x, indx, y, **hyper):
    """The Hadamard objectives the way a driver would call them (verbose=True, then `.backward()`):
    (vals = [-logpost, loglik, priors...], grad)."""
    p = torch.as_tensor(pars, dtype=DTYPE).detach().clone().requires_grad_(True)
    out = _HADAMARD[model](p, torch.as_tensor(y, dtype=DTYPE), torch.as_tensor(x, dtype=DTYPE),
                           torch.as_tensor(indx).to(torch.int64), **hyper)
    neg = -out[0]
    neg.backward()
    vals = torch.stack([neg.detach()] + [torch.as_tensor(o).detach().reshape(()) for o in out[1:]])
    return vals, p.grad.detach().clone()


_MODELS = {
    "stationary": logpost_stationary,
    "separable": logpost_separable,
    "nonseparable": logpost_nonseparable,
}


def n_params(model: str, N: int, M: int) -> int:
    T = tril_size(M)
    return {"stationary": T + 3, "separable": 2 * N + T + 1, "nonseparable": N + N * T + 1}[model]


def value_and_grad(model: str, pars, Y, x, **hyper):
    """One evaluation the way a reference driver does it: the negated log-posterior with
    verbose=True, then `.backward()`.  Returns (vals, grad): vals = [-logpost, loglik, priors...]
    as a float64 tensor, grad = d(-logpost)/d pars."""
    p = torch.as_tensor(pars, dtype=DTYPE).detach().clone().requires_grad_(True)
    Y = torch.as_tensor(Y, dtype=DTYPE)
    x = torch.as_tensor(x, dtype=DTYPE)
    out = _MODELS[model](p, Y, x, **hyper)
    neg = -out[0]
    neg.backward()
    vals = torch.stack([neg.detach()] + [torch.as_tensor(o).detach().reshape(()) for o in out[1:]])
    return vals, p.grad.detach().clone()


# ----------------------------------------------------------------------------- hyper-parameter gradient
def hyper_grad(model: str, pars, x, M: int, **hyper):
    """d(-log posterior)/d(hyper-parameters), float64 numpy [9] in the keyword order of the reference signatures
    (logpos.py:383 / :216 / :299), from the closed forms of the prior terms (logpos.py:271-292, 357-376, 444-458):
    dense inverse of the prior covariance, no factor reuse -- an independent route to what csrc/hyper.cu computes.
    Pinned to the unmodified reference's autograd through its keyword hyper-parameters
    (tests/golden/make_golden_hyper.py -> tests/golden/hyper_*.npz)."""
    import numpy as np
    from scipy.special import digamma

    p = np.asarray(pars, dtype=np.float64)
    xv = np.asarray(x, dtype=np.float64)
    N, T = xv.shape[0], tril_size(M)
    s2 = math.exp(p[-1])
    a, b = float(hyper["a"]), float(hyper["b"])
    d_a = -math.log(s2) + math.log(b) - float(digamma(a))
    d_b = -1.0 / s2 + a / b
    out = np.zeros(9)
    if model == "stationary":
        mu, sd, c = float(hyper["mu_tilde_l"]), float(hyper["sigma_tilde_l"]), float(hyper["c"])
        d, uL = p[0] - mu, p[2:2 + T]
        out[:5] = [d / sd ** 2, d * d / sd ** 3 - 1.0 / sd, d_a, d_b, (uL ** 2).sum() / c ** 3 - T / c]
        return -out
    D = (xv[:, None] - xv[None, :]) ** 2

    def gp(V, mu, alpha, beta):          # V [N, nv]: the nv vectors sharing the prior N(mu 1, alpha^2 E + jitter I)
        E = np.exp(-0.5 * D / beta ** 2)
        Si = np.linalg.inv(alpha ** 2 * E + JITTER * np.eye(N))
        G = Si @ (V - mu)
        dSa, dSb = 2.0 * alpha * E, alpha ** 2 / beta ** 3 * E * D
        nv = V.shape[1]
        quad = lambda dS: -0.5 * nv * np.sum(Si * dS) + 0.5 * np.einsum("it,ij,jt->", G, dS, G)
        return G.sum(), quad(dSa), quad(dSb)

    if model == "separable":
        k1 = ("mu_tilde_sigma", "alpha_tilde_sigma", "beta_tilde_sigma")
        V0, V1 = p[:N, None], p[N:2 * N, None]
        c = float(hyper["c"])
        uL = p[2 * N:2 * N + T]
        out[8] = (uL ** 2).sum() / c ** 3 - T / c
    else:
        k1 = ("mu_L", "alpha_L", "beta_L")
        V0, V1 = p[:N, None], p[N:N + N * T].reshape(N, T)
    out[0:3] = gp(V0, float(hyper["mu_tilde_l"]), float(hyper["alpha_tilde_l"]), float(hyper["beta_tilde_l"]))
    out[3:6] = gp(V1, *(float(hyper[k]) for k in k1))
    out[6], out[7] = d_a, d_b
    return -out
