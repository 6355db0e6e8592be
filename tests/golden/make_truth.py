"""High-precision (mpmath, 50 digits) values and gradients of the GP-prior terms for a few golden cases.

The GP-prior covariances have cond 1e8..2e10, so the reference's own FP64 result is only accurate to ~1e-9..1e-8
there.  These files let the GPU tests check that the CUDA path is no further from the exact answer
than the reference is.  Run in the build container:  python tests/golden/make_truth.py
Writes tests/golden/truth/<case>.npz with, per GP prior: exact log-density `lp` and exact gradient `dlp` w.r.t. the
parameter vector it applies to (float64-rounded), evaluated at the float64 inputs of the golden case.
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
mp.mp.dps = 50
CASES = ["separable_N200_M5_s3_h0_p", "separable_N200_M5_s4_h0_p", "nonseparable_N100_M6_s3_h0_p",
         "nonseparable_N100_M6_s4_h0_p", "separable_N30_M3_s1_h0_p", "nonseparable_N21_M3_s1_h1_p"]


def exact_prior(x, V, mu, alpha, beta):
    """V: [N, nv] columns; returns (sum of lp over columns, dlp [N, nv]) for MVN(mu 1, alpha^2 exp(-.5 d/beta^2)+1e-6 I)."""
    N = len(x)
    xs = [mp.mpf(float(v)) / mp.mpf(float(beta)) for v in x]
    a2 = mp.mpf(float(alpha)) ** 2
    C = mp.matrix(N, N)
    for i in range(N):
        for j in range(N):
            C[i, j] = a2 * mp.exp(-(xs[i] - xs[j]) ** 2 / 2)
        C[i, i] += mp.mpf("1e-6")
    L = mp.cholesky(C)
    hld = sum(mp.log(L[i, i]) for i in range(N))
    Cinv = None
    lp = mp.mpf(0)
    grad = np.zeros(V.shape)
    for v in range(V.shape[1]):
        r = mp.matrix([mp.mpf(float(t)) - mp.mpf(float(mu)) for t in V[:, v]])
        z = mp.lu_solve(L, r) if False else None
        # forward / backward substitution with the exact factor
        zz = mp.matrix(N, 1)
        for i in range(N):
            s = r[i]
            for k in range(i):
                s -= L[i, k] * zz[k]
            zz[i] = s / L[i, i]
        g = mp.matrix(N, 1)
        for i in reversed(range(N)):
            s = zz[i]
            for k in range(i + 1, N):
                s -= L[k, i] * g[k]
            g[i] = s / L[i, i]
        quad = sum(zz[i] ** 2 for i in range(N))
        lp += -(N * mp.log(2 * mp.pi) + quad) / 2 - hld
        grad[:, v] = [-float(g[i]) for i in range(N)]
    return float(lp), grad


def main():
    os.makedirs(os.path.join(HERE, "truth"), exist_ok=True)
    for name in CASES:
        z = np.load(os.path.join(HERE, name + ".npz"))
        model, N, M = str(z["model"]), int(z["N"]), int(z["M"])
        T = M * (M + 1) // 2
        h = json.loads(str(z["hyper"]))
        x, pars = z["x"], z["pars"]
        out = {}
        if model == "nonseparable":
            out["lp0"], out["dlp0"] = exact_prior(x, pars[:N, None], h["mu_tilde_l"], h["alpha_tilde_l"], h["beta_tilde_l"])
            out["lp1"], out["dlp1"] = exact_prior(x, pars[N:N + N * T].reshape(N, T), h["mu_L"], h["alpha_L"], h["beta_L"])
        else:
            out["lp0"], out["dlp0"] = exact_prior(x, pars[:N, None], h["mu_tilde_l"], h["alpha_tilde_l"], h["beta_tilde_l"])
            out["lp1"], out["dlp1"] = exact_prior(x, pars[N:2 * N, None], h["mu_tilde_sigma"], h["alpha_tilde_sigma"],
                                                  h["beta_tilde_sigma"])
        np.savez_compressed(os.path.join(HERE, "truth", name + ".npz"), **out)
        print(name, out["lp0"], out["lp1"], "ref:", z["vals"][2:4])


if __name__ == "__main__":
    main()
