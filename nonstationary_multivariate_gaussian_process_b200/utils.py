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


# ------------------------------------------------------------------------------------------------ driver-side helpers
# Data splitting and scores the drivers call around the hot path (Utility/utils.py:91-196).  Host-side numpy; the random
# splits go through the same third-party routine as the reference (sklearn's train_test_split), so a given random_state
# yields the same partition.
def data_split(x, Y, test_size=0.25, random_state=22, shuffle=True):
    """Random train / test split of (x [N], Y [N,M]), each part re-sorted by x (utils.py:138-156)."""
    from sklearn.model_selection import train_test_split
    x_tr, x_te, Y_tr, Y_te = train_test_split(x, Y, test_size=test_size, random_state=random_state, shuffle=shuffle)
    o_tr, o_te = np.argsort(x_tr), np.argsort(x_te)
    return x_tr[o_tr], x_te[o_te], Y_tr[o_tr], Y_te[o_te]


def data_split_extrapolation(x, Y, size=5):
    """Hold out the last `size` time points (utils.py:159-164)."""
    return x[:-size], x[-size:], Y[:-size], Y[-size:]


def data_split_non(x, indx, y, test_size=0.25, random_state=22, shuffle=True):
    """Random split of irregularly sampled data (x, output index, y) (utils.py:91-103)."""
    from sklearn.model_selection import train_test_split
    return tuple(train_test_split(x, indx, y, test_size=test_size, random_state=random_state, shuffle=shuffle))


def data_split_non_chunk(x, indx, y, chunk_size=0.2, random_state=22, fix=False):
    """Per output m, hold out one contiguous chunk of int(chunk_size * n_m) samples; its start is drawn with
    np.random.choice after np.random.seed(random_state), or spread evenly over the outputs when fix=True (utils.py:106-135)."""
    M = len(np.unique(indx))
    parts = [[] for _ in range(6)]
    np.random.seed(random_state)
    for m in range(M):
        sel = indx == m
        xm, ym = x[sel], y[sel]
        n_m = xm.shape[0]
        n_te = int(chunk_size * n_m)
        n_tr = n_m - n_te
        start = int(np.floor(m * n_tr / (M - 1))) if fix else np.random.choice(n_tr)
        te = np.arange(start, start + n_te)
        tr = np.concatenate([np.arange(0, start), np.arange(start + n_te, n_m)])
        for lst, v in zip(parts, (xm[tr], xm[te], m * np.ones(n_tr), m * np.ones(n_te), ym[tr], ym[te])):
            lst.append(v)
    return tuple(np.concatenate(p) for p in parts)


def MSE(x, y, axis=None):
    """Mean squared error (utils.py:167-174)."""
    return np.mean((x - y) ** 2, axis=axis)


def RMSE(x, y, axis=None):
    """Root mean squared error (utils.py:177-184)."""
    return np.sqrt(np.mean((x - y) ** 2, axis=axis))


def LPD(mean_array, std_array, y_array):
    """Mean log predictive density of y under N(mean, std^2), element-wise (utils.py:187-199)."""
    mu, sd, yv = (np.asarray(a, dtype=np.float64).reshape(-1) for a in (mean_array, std_array, y_array))
    z = (yv - mu) / sd
    return float(np.mean(-0.5 * z * z - np.log(sd) - 0.5 * np.log(2.0 * np.pi)))
