"""Parameter transforms with the reference's signatures (Utility/utils.py:10-88), vectorised.

These are host-side bookkeeping helpers the drivers use to initialise and report parameters (torch tensors or
numpy arrays in, same type out, differentiable for torch).  Inside the hot path the same transforms are fused into
the CUDA kernels (`svc_prep_kernel`, `sep_prep_kernel`); nothing here is called per evaluation.
"""
from __future__ import annotations

import numpy as np
import torch

from . import settings


def _diag_mask(M: int, like):
    T = M * (M + 1) // 2
    mask = np.zeros(T, dtype=bool)
    mask[np.cumsum(np.arange(1, M + 1)) - 1] = True          # utils.py:12
    return torch.from_numpy(mask).to(like.device) if isinstance(like, torch.Tensor) else mask


def uLvec2Lvec(uL_vec, M):
    """exp on the diagonal slots of a row-major lower-triangle vector (utils.py:10-22)."""
    mask = _diag_mask(M, uL_vec)
    if isinstance(uL_vec, torch.Tensor):
        return torch.where(mask, torch.exp(uL_vec), uL_vec)
    return np.where(mask, np.exp(uL_vec), uL_vec)


def Lvec2uLvec(L_vec, M):
    """log on the diagonal slots (utils.py:24-36)."""
    mask = _diag_mask(M, L_vec)
    if isinstance(L_vec, torch.Tensor):
        return torch.where(mask, torch.log(torch.where(mask, L_vec, torch.ones_like(L_vec))), L_vec)
    return np.where(mask, np.log(np.where(mask, L_vec, 1.0)), L_vec)


def uLvecs2Lvecs(uL_vecs, N, M):
    """N stacked time-major triangles (utils.py:38-46)."""
    T = M * (M + 1) // 2
    return uLvec2Lvec(uL_vecs.reshape(N, T), M).reshape(-1)


def Lvecs2uLvecs(L_vecs, N, M):
    """(utils.py:48-54)"""
    T = M * (M + 1) // 2
    return Lvec2uLvec(L_vecs.reshape(N, T), M).reshape(-1)


def vec2lowtriangle(x, N=None):
    """Scatter a length N(N+1)/2 vector into an N x N lower-triangular matrix (utils.py:56-74)."""
    if N * (N + 1) // 2 != x.shape[0]:
        raise ValueError("check the dimension size!")
    if isinstance(x, torch.Tensor):
        mat = torch.zeros([N, N], dtype=x.dtype, device=x.device)
        idx = torch.tril_indices(N, N, device=x.device)
        mat[idx[0], idx[1]] = x
        return mat
    mat = np.zeros([N, N])
    r, c = np.tril_indices(N)
    mat[r, c] = x
    return mat


def lowtriangle2vec(L, N=None):
    """(utils.py:77-88)"""
    if N is None:
        N = L.shape[0]
    if isinstance(L, torch.Tensor):
        idx = torch.tril_indices(N, N, device=L.device)
        return L[idx[0], idx[1]]
    r, c = np.tril_indices(N)
    return L[r, c]
