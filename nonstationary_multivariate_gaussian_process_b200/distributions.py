"""Densities with the reference's signatures (Utility/distributions.py).

Inside the objectives of `logpos` these are fused into the CUDA path; the stand-alone versions below serve callers
that use them directly.  The Kronecker density is evaluated on the GPU (library eigensolvers, cold path).
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib, kronecker_operation


def multivariate_normal_logpdf(y, mu, logdetSigma, invSigma):
    """Un-normalised: -0.5 logdet - 0.5 (y-mu)^T invSigma (y-mu)  (distributions.py:10-23)."""
    torch = _lib.require_cuda()
    y_bar = y - mu
    return -0.5 * logdetSigma - 0.5 * torch.dot(y_bar, torch.mv(invSigma, y_bar))


def multivariate_normal_logpdf0(y, mu, B, K, sigma2):
    """Un-normalised log N(y; mu, B (x) K + sigma2 I), y output-major (distributions.py:26-52)."""
    torch = _lib.require_cuda()
    dev = torch.as_tensor(y).device
    Bg, Kg, yg, mg = kronecker_operation._gpu(B, K, y, mu)
    wB, vB = torch.linalg.eigh(Bg)
    wK, vK = torch.linalg.eigh(Kg)
    a = kronecker_operation.kron_mv(vB.t(), vK.t(), yg - mg)
    t = (wB.view(-1, 1) * wK.view(1, -1)).reshape(-1)
    s2 = torch.as_tensor(sigma2, dtype=t.dtype).to(t.device)
    return (-0.5 * torch.log(t + s2).sum() - 0.5 * torch.dot(a / (s2 + t), a)).to(dev)


# The reference's "robust" variant perturbs the diagonals with unseeded random jitter to dodge NaNs of its
# eigen-gradient (distributions.py:55-96); the value path here has no such failure mode.
multivariate_normal_logpdf1 = multivariate_normal_logpdf0


def multivariate_normal_logpdf2(y, mu, B, K, sigma2):
    """Dense check of the same density (distributions.py:99-113)."""
    torch = _lib.require_cuda()
    dev = torch.as_tensor(y).device
    Bg, Kg, yg, mg = kronecker_operation._gpu(B, K, y, mu)
    n = Bg.shape[0] * Kg.shape[0]
    s2 = torch.as_tensor(sigma2, dtype=Bg.dtype).to(Bg.device)
    Sigma = torch.kron(Bg, Kg) + s2 * torch.eye(n, dtype=Bg.dtype, device=Bg.device)
    Lc = torch.linalg.cholesky(Sigma)
    z = torch.linalg.solve_triangular(Lc, (yg - mg).unsqueeze(-1), upper=False).squeeze(-1)
    return (-torch.log(torch.diagonal(Lc)).sum() - 0.5 * torch.dot(z, z)).to(dev)


def inverse_gamma_logpdf_u(x, alpha=1., beta=1.):
    """(distributions.py:116-124)"""
    torch = _lib.require_cuda()
    return (-alpha - 1) * torch.log(x) - beta / x


def inverse_gamma_logpdf(x, alpha=1., beta=1.):
    """(distributions.py:126-134)"""
    torch = _lib.require_cuda()
    return (-alpha - 1) * torch.log(x) - beta / x + alpha * np.log(beta) - math.lgamma(alpha)


def gamma_logpdf(x, alpha=1., beta=1.):
    """(distributions.py:136-137)"""
    torch = _lib.require_cuda()
    return (alpha - 1) * torch.log(x) - beta * x + alpha * np.log(beta) - math.lgamma(alpha)
