"""Empirical initialisation of the model parameters by windowed estimation, with the reference's names, signatures and
return conventions (Utility/empirical_estimation.py:33-133).

`local_estimation` runs once per subject before the MAP loop: per time point a window of observations, per output an
experimental semivariogram and a two-parameter Gaussian-variogram fit, plus the windowed covariance and its factor.  It is
host code here as well: the fit itself is `scipy.optimize.curve_fit` in the reference, and calling the same routine on the
same numbers is what makes the initial parameter vector -- and with it the whole MAP trajectory -- identical.  What changes is
the semivariogram: the reference builds it with a Python double loop over all pairs of the window (N * M * w^2 / 2
interpreter iterations per subject, the bulk of its run time); here it is one vectorised numpy expression producing the same
floating-point values in the same (i < j, row-major) order.  Plotting (`visualization`, the `check=` branch) is not part of
this package.
"""
from __future__ import annotations

import pickle

import numpy as np
from scipy.optimize import curve_fit

from . import settings, utils


def SV(x, Y, indx):
    """Experimental semivariogram of output `indx` over all pairs i < j of the window, in the reference's order
    (empirical_estimation.py:33-55): lag = x_j - x_i, sv = 0.5 (Y[j] - Y[i])^2."""
    N = x.shape[0]
    i, j = np.triu_indices(N, k=1)                      # row-major pairs (0,1), (0,2), ..., (N-2,N-1)
    d = Y[j, indx] - Y[i, indx]
    # the reference squares numpy SCALARS (`(a - b)**2` -> libm pow), which differs from the array square d * d by one ulp in
    # ~0.1 % of the pairs -- enough to move an ill-posed variogram fit; np.float_power goes through the same pow
    return x[j] - x[i], 0.5 * np.float_power(d, 2)


def variogram_Gaussian(s, sigma, l):
    """sigma^2 (1 - exp(-s^2 / (2 l^2)))  (empirical_estimation.py:58-59)."""
    return sigma ** 2 * (1 - np.exp(-0.5 * s ** 2 / l ** 2))


def global_estimation(x, Y):
    """Sample covariance of the outputs and its lower factor as a row-major triangle (empirical_estimation.py:62-67)."""
    M = Y.shape[1]
    S = np.cov(Y.T)
    return S, utils.lowtriangle2vec(np.linalg.cholesky(S), M)


def local_estimation(x, Y, window_size=30, save_dir=None, folder_name=None, subfolder_name=None, check=False):
    """Windowed estimates at every time point (empirical_estimation.py:70-133).  Returns
    (est_sigmas [N], est_ls [N], smooth_ls [N], est_stds [N,M], est_R [N,M,M], est_B [N,M,M], est_L_vecs [N*T],
    est_tilde_sigma2_err = -4)."""
    if check:
        raise NotImplementedError("the plotting / pdb branch of the reference (check=True) is not part of this package")
    N, M = Y.shape
    est_sigmas, est_ls, est_B, est_L_vecs, est_stds, est_R = [], [], [], [], [], []
    for n in range(N):
        start, end = max(0, n - window_size), min(n + window_size, N - 1)
        xn, Yn = x[start:end], Y[start:end]
        cofs = []
        for m in range(M):
            lag, sv = SV(xn, Yn, m)
            cof_u, _ = curve_fit(variogram_Gaussian, lag, sv, maxfev=2000)
            cofs.append(cof_u)
        cof = np.mean(np.stack(cofs), axis=0)
        est_sigmas.append(np.abs(cof[0]))
        est_ls.append(np.abs(cof[1]))
        S = np.matmul(Yn.T, Yn) / (Yn.shape[0] - 1)
        try:
            L_f = np.linalg.cholesky(S)
        except np.linalg.LinAlgError:
            S = S + np.diag(np.ones(M) * settings.precision)
            L_f = np.linalg.cholesky(S)
        est_B.append(S)
        est_L_vecs.append(utils.lowtriangle2vec(L_f, M).reshape(-1))
        D = np.sqrt(np.diag(S))
        est_stds.append(D)
        est_R.append(np.diag(1. / D).dot(S).dot(np.diag(1. / D)))
    est_ls = np.array(est_ls)
    smooth_ls = np.array([np.mean(est_ls[max(0, n - 10):min(n + 10, N - 1)]) for n in range(N)])
    return (np.array(est_sigmas), est_ls, smooth_ls, np.stack(est_stds), np.stack(est_R), np.stack(est_B),
            np.concatenate(est_L_vecs), -4)


def save_res(est_ls, smooth_ls, est_L_vecs, est_tilde_sigma2_err, save_dir=None, folder_name=None, subfolder_name=None):
    """Pickle [log est_ls, log smooth_ls, est_L_vecs, est_tilde_sigma2_err] where the drivers look for them
    (empirical_estimation.py:184-190)."""
    path = save_dir + folder_name + (subfolder_name if subfolder_name is not None else "")
    with open(path + "empirical_est.pickle", "wb") as res:
        pickle.dump([np.log(est_ls), np.log(smooth_ls), est_L_vecs, est_tilde_sigma2_err], res)
