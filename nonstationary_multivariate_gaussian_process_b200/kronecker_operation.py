"""Kronecker helpers with the reference's signatures (Utility/kronecker_operation.py), on this library's CUDA kernels.

The hot path never forms a Kronecker product (the separable objective uses eig(B) + M Cholesky factors of
lam_m K + sigma2 I; the nonseparable build kernel fuses `ones(M,M) (x) K_x`).  These helpers exist so that
`Utility/prediction.py` / `SIM_code/sim.py`-style callers keep working: CPU (or CUDA) tensors in, same device out,
computed by `nmgp_kron`, `nmgp_kron_mv` and `nmgp_kron_eig_solve` (csrc/kron.cu; cold paths: SURVEY.md 8a rows a8-a10).
"""
from __future__ import annotations

import ctypes

from . import _lib


def _gpu(*ts):
    torch = _lib.require_cuda()
    return [torch.as_tensor(t, dtype=torch.float64).detach().cuda().contiguous() for t in ts]


def _stream():
    torch = _lib.require_cuda()
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def kronecker_product(t1, t2):
    """t1 (x) t2 for 2-D tensors (kronecker_operation.py:5-22)."""
    torch = _lib.require_cuda()
    dev = torch.as_tensor(t1).device
    a, b = _gpu(t1, t2)
    out = torch.empty((a.shape[0] * b.shape[0], a.shape[1] * b.shape[1]), dtype=torch.float64, device=a.device)
    if out.numel():
        _lib.check(_lib.load_library().nmgp_kron(a.data_ptr(), a.shape[0], a.shape[1], b.data_ptr(), b.shape[0], b.shape[1],
                                                 out.data_ptr(), _stream()), "nmgp_kron")
    return out.to(dev)


def kronecker_product_diag(d1, d2):
    """Diagonal of D1 (x) D2 (kronecker_operation.py:25-33)."""
    torch = _lib.require_cuda()
    return kronecker_product(torch.as_tensor(d1).reshape(-1, 1), torch.as_tensor(d2).reshape(-1, 1)).reshape(-1)


def kron_mv(B, K, y):
    """(B (x) K) y as vec(K Y B^T), y output-major (kronecker_operation.py:72-85)."""
    torch = _lib.require_cuda()
    dev = torch.as_tensor(y).device
    Bg, Kg, yg = _gpu(B, K, y)
    (M1, M2), (N1, N2) = Bg.shape, Kg.shape
    if yg.numel() != M2 * N2:
        raise RuntimeError(f"kron_mv: y has {yg.numel()} entries, expected {M2 * N2}")
    out = torch.empty(M1 * N1, dtype=torch.float64, device=Bg.device)
    scratch = torch.empty(max(M2 * N1, 1), dtype=torch.float64, device=Bg.device)
    if out.numel():
        _lib.check(_lib.load_library().nmgp_kron_mv(Bg.data_ptr(), M1, M2, Kg.data_ptr(), N1, N2, yg.data_ptr(), out.data_ptr(),
                                                    scratch.data_ptr(), _stream()), "nmgp_kron_mv")
    return out.to(dev)


def _eig_solve(sigma2, B, K, r=None, want_inverse=False):
    """(log det, r^T inv r, inverse or None) of sigma2 I + B (x) K through nmgp_kron_eig_solve."""
    torch = _lib.require_cuda()
    Bg, Kg = _gpu(B, K)
    M, N = Bg.shape[0], Kg.shape[0]
    if M > 16:
        raise NotImplementedError("kron_inv / kron_logdet: B larger than 16 x 16 (the models have M <= 16 outputs)")
    rg = _gpu(r)[0] if r is not None else None
    inv = torch.empty((M * N, M * N), dtype=torch.float64, device=Bg.device) if want_inverse else None
    out2 = torch.empty(2, dtype=torch.float64, device=Bg.device)
    info = torch.zeros(1, dtype=torch.int32, device=Bg.device)
    _lib.check(_lib.load_library().nmgp_kron_eig_solve(
        Bg.data_ptr(), M, Kg.data_ptr(), N, float(sigma2), inv.data_ptr() if want_inverse else None,
        rg.data_ptr() if rg is not None else None, out2.data_ptr(), info.data_ptr(), _stream()), "nmgp_kron_eig_solve")
    if int(info.item()) != 0:     # not positive definite: the reference's eigen path would return garbage / NaN silently
        out2 = torch.full_like(out2, float("nan"))
        if inv is not None:
            inv.fill_(float("nan"))
    return out2[0], out2[1], inv


def kron_inv(sigma2, B, K):
    """inv(sigma2 I + B (x) K) (kronecker_operation.py:36-54), via eig(B) and M Cholesky factors."""
    torch = _lib.require_cuda()
    dev = torch.as_tensor(B).device
    return _eig_solve(sigma2, B, K, want_inverse=True)[2].to(dev)


def kron_logdet(sigma2, B, K):
    """log det(sigma2 I + B (x) K) (kronecker_operation.py:57-69)."""
    torch = _lib.require_cuda()
    dev = torch.as_tensor(B).device
    return _eig_solve(sigma2, B, K)[0].to(dev)
