"""Covariance kernels with the reference's signatures (Utility/kernels.py), computed by the CUDA library.

Inputs may be CPU tensors (what `Utility/prediction.py` and `SIM_code/sim.py` pass): they are copied to the GPU,
evaluated by `nmgp_rbf_cov` / `nmgp_gibbs_cov` and copied back, so callers see CPU-tensor-in / CPU-tensor-out.
CUDA tensors stay on the device.  Not differentiable (the objectives in `logpos` carry their own analytic
gradient); inputs must be N x 1 as everywhere in the reference.
"""
from __future__ import annotations

import ctypes

from . import _lib


def _prep(t, name):
    torch = _lib.require_cuda()
    if t is None:
        return None
    t = torch.as_tensor(t, dtype=torch.float64).detach()
    if t.dim() == 2:
        if t.shape[1] != 1:
            raise NotImplementedError(f"{name}: only N x 1 inputs are supported (as used throughout the reference)")
        t = t[:, 0]
    return t.contiguous().cuda()


def _ptr(t):
    return None if t is None else t.data_ptr()


def pairwise_distances(x, y=None):
    """|x_i|^2 + |y_j|^2 - 2 x_i.y_j for N x 1 inputs (kernels.py:5-21), by `nmgp_pairwise_sqdist`."""
    torch = _lib.require_cuda()
    lib = _lib.load_library()
    dev_in = torch.as_tensor(x).device
    xs = _prep(x, "pairwise_distances")
    ys = None if y is None else _prep(y, "pairwise_distances")
    n1, n2 = xs.numel(), (xs.numel() if ys is None else ys.numel())
    out = torch.empty((n1, n2), dtype=torch.float64, device=xs.device)
    if n1 and n2:
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.nmgp_pairwise_sqdist(xs.data_ptr(), n1, _ptr(ys), n2, out.data_ptr(), ctypes.c_void_p(stream)),
                   "nmgp_pairwise_sqdist")
    return out.to(dev_in)


def RBF_cov(X1, X2=None, alpha=1., beta=1.):
    """alpha^2 exp(-0.5 |(x1_i - x2_j)/beta|^2), + jitter*I when X2 is None (kernels.py:24-43)."""
    torch = _lib.require_cuda()
    lib = _lib.load_library()
    dev_in = torch.as_tensor(X1).device
    x1 = _prep(X1, "RBF_cov")
    x2 = _prep(X2, "RBF_cov")
    n1, n2 = x1.numel(), (x1.numel() if x2 is None else x2.numel())
    out = torch.empty((n1, n2), dtype=torch.float64, device=x1.device)
    if n1 and n2:
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.nmgp_rbf_cov(x1.data_ptr(), n1, _ptr(x2), n2, float(alpha), float(beta), out.data_ptr(),
                                    ctypes.c_void_p(stream)), "nmgp_rbf_cov")
    return out.to(dev_in)


def Nonstationary_RBF_cov(X1, sigma1=None, ell1=None, X2=None, sigma2=None, ell2=None):
    """Gibbs / Paciorek kernel (kernels.py:46-73); sigma / ell default to ones; + jitter*I when X2 is None."""
    torch = _lib.require_cuda()
    lib = _lib.load_library()
    dev_in = torch.as_tensor(X1).device
    x1 = _prep(X1, "Nonstationary_RBF_cov")
    x2 = _prep(X2, "Nonstationary_RBF_cov")
    n1 = x1.numel()
    s1 = _prep(sigma1, "sigma1")
    l1 = _prep(ell1, "ell1") if ell1 is not None else torch.ones(n1, dtype=torch.float64, device=x1.device)
    if x2 is None:
        n2, s2, l2 = n1, None, None
    else:
        n2 = x2.numel()
        s2 = _prep(sigma2, "sigma2")
        l2 = _prep(ell2, "ell2") if ell2 is not None else torch.ones(n2, dtype=torch.float64, device=x1.device)
    out = torch.empty((n1, n2), dtype=torch.float64, device=x1.device)
    if n1 and n2:
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.nmgp_gibbs_cov(x1.data_ptr(), _ptr(s1), l1.data_ptr(), n1, _ptr(x2), _ptr(s2), _ptr(l2), n2,
                                      out.data_ptr(), ctypes.c_void_p(stream)), "nmgp_gibbs_cov")
    return out.to(dev_in)
