"""Densities with the reference's signatures (Utility/distributions.py).

Inside the objectives of `logpos` these are fused into the CUDA path; the stand-alone versions below serve callers
that use them directly.  The Kronecker densities run on this library's kernels (`nmgp_kron_eig_solve`, csrc/kron.cu).
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
    """Un-normalised log N(y; mu, B (x) K + sigma2 I), y output-major (distributions.py:26-52): eig(B) + M Cholesky
    factorisations on the batched engine (`nmgp_kron_eig_solve`) instead of the reference's two symeig calls."""
    torch = _lib.require_cuda()
    dev = torch.as_tensor(y).device
    r = torch.as_tensor(y, dtype=torch.float64) - torch.as_tensor(mu, dtype=torch.float64)
    logdet, quad, _ = kronecker_operation._eig_solve(sigma2, B, K, r=r)
    return (-0.5 * logdet - 0.5 * quad).to(dev)


# The reference's "robust" variant perturbs the diagonals of B and K with UNSEEDED uniform jitter of size 1e-6 to dodge
# NaNs of its eigen-gradient (distributions.py:55-96); the Cholesky path has no such failure mode, so the variant is the
# plain density (it differs from the reference's draw-dependent value by O(1e-6) relative, like two reference calls do).
multivariate_normal_logpdf1 = multivariate_normal_logpdf0


def multivariate_normal_logpdf2(y, mu, B, K, sigma2):
    """The reference's dense cross-check of the same density (distributions.py:99-113: kronecker_product + logdet +
    inverse).  Here: the explicit inverse and log-determinant from `kron_inv` / `kron_logdet`, then the dense formula."""
    torch = _lib.require_cuda()
    dev = torch.as_tensor(y).device
    logdet, _, inv = kronecker_operation._eig_solve(sigma2, B, K, want_inverse=True)
    yg, mg = kronecker_operation._gpu(y, mu)
    return multivariate_normal_logpdf(yg, mg, logdet, inv).to(dev)


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
