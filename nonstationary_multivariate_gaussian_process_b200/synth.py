"""Synthetic multivariate non-stationary time series in the shape of the reference's simulator.

Restates the recipe of SIM_code/sim.py:173-275 (`SIM_MNTS`) for a general number of outputs M
(the reference hard-wires M=2 at sim.py:220,243-249):

  x            = sort(U(0,1)^N)                                   sim.py:177
  tilde_l(x)   = 3 (x-1)^3 - 3                                    sim.py:180
  std_m(x)     = 1 + x^2 (m even) | 2 - x^2 (m odd)               sim.py:220
  corr(x)      = cos(pi x)^|m-m'|   (== cos(pi x) for M=2)        sim.py:235,243-249
  L_i          = chol(D_i R_i D_i)                                sim.py:250
  sigma2_err   = 1e-2                                             sim.py:256
  y            ~ N(0, K + sigma2_err I), K as in logpos.py:339-349 sim.py:258-264

Host-side numpy only (input generation, not part of the timed path).
"""
from __future__ import annotations

import numpy as np


def tril_size(M: int) -> int:
    return M * (M + 1) // 2


def diag_slots(M: int) -> np.ndarray:
    return np.cumsum(np.arange(1, M + 1)) - 1


def truth(N: int, M: int, seed: int):
    """x [N], tilde_l [N], uL [N,T] (unconstrained: log on the diagonal slots), tilde_sigma2_err."""
    rng = np.random.RandomState(seed)
    x = np.sort(rng.rand(N))
    tilde_l = 3.0 * (x - 1.0) ** 3 - 3.0
    stds = np.stack([(1.0 + x ** 2) if m % 2 == 0 else (2.0 - x ** 2) for m in range(M)], axis=1)  # N,M
    rho = np.cos(np.pi * x)
    idx = np.arange(M)
    lag = np.abs(idx[:, None] - idx[None, :])
    R = rho[:, None, None] ** lag[None]                      # N,M,M  AR(1)-type, PD for |rho|<1
    B = stds[:, :, None] * R * stds[:, None, :]
    L = np.linalg.cholesky(B)                                # N,M,M
    r, c = np.tril_indices(M)
    Lvec = L[:, r, c]                                        # N,T row-major lower triangle
    uL = Lvec.copy()
    d = diag_slots(M)
    uL[:, d] = np.log(Lvec[:, d])
    return x, tilde_l, uL, np.log(1e-2)


def gibbs_cov(x, ell):
    A = ell[:, None] ** 2 + ell[None, :] ** 2
    d = (x[:, None] - x[None, :]) ** 2
    K = np.sqrt(2.0 * ell[:, None] * ell[None, :] / A) * np.exp(-d / A)
    K[np.diag_indices_from(K)] += 1e-6
    return K


def dense_cov_time_major(x, tilde_l, uL, M):
    """Nonseparable covariance (no noise) in time-major ordering (row (i,m) -> i*M+m)."""
    N = x.shape[0]
    r, c = np.tril_indices(M)
    Lvec = uL.copy()
    d = diag_slots(M)
    Lvec[:, d] = np.exp(uL[:, d])
    L = np.zeros((N, M, M))
    L[:, r, c] = Lvec
    Ls = L.reshape(N * M, M)
    Kx = gibbs_cov(x, np.exp(tilde_l))
    return np.kron(Kx, np.ones((M, M))) * (Ls @ Ls.T)


def sample_subject(N: int, M: int, seed: int):
    """One synthetic subject drawn on the host: returns x [N], Y [N,M] and the generating
    parameter vector in the nonseparable layout [tilde_l(N), uL(N*T, time-major), tilde_sigma2_err]."""
    x, tilde_l, uL, ts2 = truth(N, M, seed)
    K = dense_cov_time_major(x, tilde_l, uL, M)
    K[np.diag_indices_from(K)] += np.exp(ts2)
    z = np.random.RandomState(10_000_019 + seed).standard_normal(N * M)
    y = np.linalg.cholesky(K) @ z
    Y = y.reshape(N, M)
    pars = np.concatenate([tilde_l, uL.reshape(-1), [ts2]])
    return x, Y, pars


def start_point(model: str, N: int, M: int, seed: int, noise: float = 0.1):
    """A driver-like evaluation point for `model`: the generating parameters perturbed by
    noise*N(0,1).  noise=0 gives the smooth point."""
    x, tilde_l, uL, ts2 = truth(N, M, seed)
    rng = np.random.RandomState(777 + seed)
    T = tril_size(M)
    if model == "nonseparable":
        p = np.concatenate([tilde_l, uL.reshape(-1), [-4.0]])
    elif model == "separable":
        p = np.concatenate([tilde_l, np.zeros(N), uL.mean(0), [-4.0]])
    elif model == "stationary":
        p = np.concatenate([[tilde_l.mean()], [0.0], uL.mean(0), [-4.0]])
    else:
        raise ValueError(model)
    return p + noise * rng.standard_normal(p.shape[0])


def hadamard_case(model: str, N: int, M: int, seed: int, noise: float = 0.1):
    """Irregularly sampled data for the Hadamard objectives (Utility/logpos.py:465-716): N observation times, each belonging
    to ONE of the M outputs (every output present).  Returns x [N], indx [N] (int64), y [N] and a parameter vector in the
    layout of `model` in {'hadamard', 'hadamard_svc', 'hadamard_s'} whose cross-output factor entries are RAW (these
    objectives apply no exp to the diagonal, logpos.py:518, 582-583, 682)."""
    x, Y, _ = sample_subject(N, M, seed)
    rng = np.random.RandomState(4242 + seed)
    indx = np.concatenate([np.arange(M), rng.randint(0, M, size=N - M)])
    rng.shuffle(indx)
    y = Y[np.arange(N), indx]
    base = {"hadamard": "separable", "hadamard_svc": "nonseparable", "hadamard_s": "stationary"}[model]
    p = start_point(base, N, M, seed, noise)
    T = tril_size(M)
    d = diag_slots(M)
    if model == "hadamard_svc":
        blk = p[N:N + N * T].reshape(N, T)
        blk[:, d] = np.exp(blk[:, d])
    else:
        o = 2 * N if model == "hadamard" else 2
        blk = p[o:o + T]
        blk[d] = np.exp(blk[d])
    return x, indx.astype(np.int64), y, p
