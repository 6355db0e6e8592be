"""Kronecker helpers with the reference's signatures (Utility/kronecker_operation.py).

The hot path never forms a Kronecker product (the separable objective uses eig(B) + M Cholesky factors of
lam_m K + sigma2 I; the nonseparable build kernel fuses `ones(M,M) (x) K_x`).  These helpers exist so that
`Utility/prediction.py`-style callers keep working: CPU (or CUDA) tensors in, same device out, evaluated on
the GPU with plain library ops (they are cold: SURVEY.md section 8a rows a8-a10).
"""
from __future__ import annotations

from . import _lib


def _gpu(*ts):
    torch = _lib.require_cuda()
    return [torch.as_tensor(t).cuda() for t in ts]


def kronecker_product(t1, t2):
    """(kronecker_operation.py:5-22)"""
    torch = _lib.require_cuda()
    dev = torch.as_tensor(t1).device
    a, b = _gpu(t1, t2)
    return torch.kron(a, b).to(dev)


def kronecker_product_diag(d1, d2):
    """Diagonal of D1 (x) D2 (kronecker_operation.py:25-33)."""
    torch = _lib.require_cuda()
    dev = torch.as_tensor(d1).device
    a, b = _gpu(d1, d2)
    return (a.view(-1, 1) * b.view(1, -1)).reshape(-1).to(dev)


def kron_mv(B, K, y):
    """(B (x) K) y as vec(K Y B^T), y output-major (kronecker_operation.py:72-85)."""
    torch = _lib.require_cuda()
    dev = torch.as_tensor(y).device
    Bg, Kg, yg = _gpu(B, K, y)
    Ym = yg.view(Bg.shape[1], Kg.shape[1]).t()
    return (Kg @ Ym @ Bg.t()).t().contiguous().view(-1).to(dev)


def _eig_pair(B, K):
    torch = _lib.require_cuda()
    Bg, Kg = _gpu(B, K)
    wB, vB = torch.linalg.eigh(Bg)
    wK, vK = torch.linalg.eigh(Kg)
    return wB, vB, wK, vK


def kron_inv(sigma2, B, K):
    """inv(sigma2 I + B (x) K) through the two eigendecompositions (kronecker_operation.py:36-54)."""
    torch = _lib.require_cuda()
    dev = torch.as_tensor(B).device
    wB, vB, wK, vK = _eig_pair(B, K)
    U = torch.kron(vB, vK)
    t = (wB.view(-1, 1) * wK.view(1, -1)).reshape(-1)
    s2 = torch.as_tensor(sigma2, dtype=t.dtype).to(t.device)
    return ((U / (t + s2)) @ U.t()).to(dev)


def kron_logdet(sigma2, B, K):
    """log det(sigma2 I + B (x) K) (kronecker_operation.py:57-69)."""
    torch = _lib.require_cuda()
    dev = torch.as_tensor(B).device
    wB, _, wK, _ = _eig_pair(B, K)
    t = (wB.view(-1, 1) * wK.view(1, -1)).reshape(-1)
    s2 = torch.as_tensor(sigma2, dtype=t.dtype).to(t.device)
    return torch.log(t + s2).sum().to(dev)
