"""Golden vectors for posterior prediction of the nonseparable model, produced by
running the UNMODIFIED reference `Utility/prediction.py:1038-1262`.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_predict.py

Harness (no reference file is touched): `torch.symeig` / `torch.solve` / `torch.cholesky` compatibility shims
 and a recording stand-in for `torch.distributions.Normal` inside the reference module's namespace, so
that every draw's (loc, scale, sample) is captured in call order -- the reference returns only percentiles / mean / std
of the sampled outputs.  One .npz per case: inputs, the seed, the function's return values in its three modes
(default, pred_smoothness, pred_cov), and the recorded per-draw triplets of the default mode:

    l_loc, l_scale, l_draw      [G, ns]       tilde_l_star ~ Normal(mu_l, sqrt(sigma2_l))        prediction.py:1113
    u_loc, u_scale, u_draw      [G, ns, T]    uL_vec_star  ~ Normal(mu_uL_vec, sqrt(sigma2_uL))  prediction.py:1124
    y_loc, y_scale, y_draw      [G, ns, M]    sampled_y    ~ Normal(mu_f, sqrt(sigma2_y))        prediction.py:1169

plus `map_percentiles`, `map_Lvecs` (pointwise_predmap_inhomogeneous, prediction.py:990-1012: plug-in of the conditional
means, no sampling) and `hist_*` (pointwise_predsample_inhomogeneous, prediction.py:1359-1378: one draw per posterior
sample of a short synthetic parameter history; the same triplets with the history index in place of the sample index).
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

torch.symeig = lambda A, eigenvectors=False, upper=True: torch.linalg.eigh(A, UPLO="U" if upper else "L")
torch.solve = lambda input, A: (torch.linalg.solve(A, input), None)
torch.cholesky = lambda A, upper=False: torch.linalg.cholesky(A, upper=upper)
from Utility import prediction  # noqa: E402  (the reference)

from nonstationary_multivariate_gaussian_process_b200 import synth  # noqa: E402

_RealNormal = prediction.Normal
RECORD = []


class RecordingNormal:
    """Same draws as torch.distributions.Normal (it IS one inside); remembers (loc, scale, sample)."""

    def __init__(self, loc, scale):
        self.d = _RealNormal(loc=loc, scale=scale)
        self.loc, self.scale = loc, scale

    def sample(self):
        s = self.d.sample()
        RECORD.append((torch.as_tensor(self.loc).detach().clone().reshape(-1), torch.as_tensor(self.scale).detach().clone().reshape(-1),
                       s.detach().clone().reshape(-1)))
        return s


prediction.Normal = RecordingNormal

HYPER = [
    # the driver's dictionary (Nonseparable_model.py:274-275): prior covariances with cond ~ 1e9 and more
    {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0},
    # short prior length scales: well-conditioned prior covariances, so the conditional moments are reproducible to 1e-9
    {"mu_tilde_l": -1.0, "alpha_tilde_l": 2.0, "beta_tilde_l": 0.03, "mu_L": 0.2, "alpha_L": 1.5, "beta_L": 0.02},
]

# (N, M, seed, noise, hyper index, number of grid points, n_sample, grid kind)
CASES = [
    (12, 2, 0, 0.1, 1, 4, 5, "grid"),
    (25, 3, 1, 0.1, 1, 6, 7, "grid"),
    (40, 6, 2, 0.05, 1, 5, 6, "grid"),    # T = 21 >= 16: torch's vectorised normal_ path for the uL draws
    (30, 4, 3, 0.1, 1, 30, 3, "train"),   # x_test = x (Nonseparable_model.py:387)
    (25, 3, 4, 0.1, 0, 6, 7, "grid"),     # driver hyper-parameters (ill-conditioned priors)
    (100, 6, 5, 0.1, 0, 3, 4, "grid"),    # BASELINE config 4 per-subject shape
]


H_HIST = 3   # posterior samples used by the history variant (the history holds one more: `[-N_sample:]` is exercised)
G_HIST = 3   # new inputs for the history variant


def stack(records, width):
    loc = np.stack([np.broadcast_to(r[0].numpy(), (width,)) for r in records])
    sc = np.stack([np.broadcast_to(r[1].numpy(), (width,)) for r in records])
    dr = np.stack([r[2].numpy() for r in records])
    return loc, sc, dr


def main():
    torch.set_num_threads(8)
    names = []
    for N, M, seed, noise, hidx, G, ns, kind in CASES:
        T = M * (M + 1) // 2
        x, Y, _ = synth.sample_subject(N, M, seed)
        pars = synth.start_point("nonseparable", N, M, seed, noise)
        hyper = HYPER[hidx]
        tl = torch.from_numpy(pars[:N].copy())
        uL = torch.from_numpy(pars[N:N + N * T].copy())
        ts = torch.tensor(float(pars[-1]), dtype=torch.float64)
        xt, Yt = torch.from_numpy(x), torch.from_numpy(Y)
        grids = torch.from_numpy(x.copy()) if kind == "train" else torch.linspace(0.03, 0.97, G, dtype=torch.float64)
        G = grids.numel()
        out = {}
        _stdout = sys.stdout
        sys.stdout = open(os.devnull, "w")   # the reference prints every grid point
        try:
            RECORD.clear()
            torch.manual_seed(1000 + seed)
            q, mean, std = prediction.pointwise_predmap_inhomogeneous_sampling(ns, tl, uL, ts, Yt, xt, grids, **hyper)
            rec = list(RECORD)
            RECORD.clear()
            torch.manual_seed(2000 + seed)
            smooth = prediction.pointwise_predmap_inhomogeneous_sampling(ns, tl, uL, ts, Yt, xt, grids, pred_smoothness=True, **hyper)
            RECORD.clear()
            torch.manual_seed(3000 + seed)
            cov = prediction.pointwise_predmap_inhomogeneous_sampling(ns, tl, uL, ts, Yt, xt, grids, pred_cov=True, **hyper)
            RECORD.clear()
            torch.manual_seed(4000 + seed)
            tq, tmean, tstd = prediction.test_predmap_inhomogeneous_sampling(ns, tl, uL, ts, Yt, xt, grids[:2], **hyper)
            # plug-in (no sampling) variant, prediction.py:912-1036
            RECORD.clear()
            map_pct, map_Lvecs = prediction.pointwise_predmap_inhomogeneous(tl, uL, ts, Yt, xt, grids, **hyper)
            # posterior-sample variant, prediction.py:1265-1400: a short synthetic history of parameter vectors
            hist = np.stack([synth.start_point("nonseparable", N, M, 100 + seed + h, noise) for h in range(H_HIST + 1)])
            ht = torch.from_numpy(hist)
            RECORD.clear()
            torch.manual_seed(5000 + seed)
            hist_y = prediction.pointwise_predsample_inhomogeneous(ht[:, :N], ht[:, N:N + N * T], ht[:, -1], Yt, xt,
                                                                   grids[:G_HIST], N_sample=H_HIST, **hyper)
            hrec = list(RECORD)
        finally:
            sys.stdout.close()
            sys.stdout = _stdout
        assert len(rec) == 3 * G * ns
        Gh = min(G_HIST, G)
        assert len(hrec) == 3 * Gh * H_HIST and hist_y.shape == (Gh, H_HIST, M)
        hl = stack(hrec[0::3], 1)
        hu = stack(hrec[1::3], T)
        hy = stack(hrec[2::3], M)
        out.update(
            map_percentiles=map_pct.numpy(), map_Lvecs=map_Lvecs.numpy(), hist_pars=hist, hist_n_sample=H_HIST,
            hist_grids=grids[:G_HIST].numpy(), hist_y=hist_y,
            hist_l_loc=hl[0].reshape(Gh, H_HIST), hist_l_scale=hl[1].reshape(Gh, H_HIST), hist_l_draw=hl[2].reshape(Gh, H_HIST),
            hist_u_loc=hu[0].reshape(Gh, H_HIST, T), hist_u_scale=hu[1].reshape(Gh, H_HIST, T),
            hist_u_draw=hu[2].reshape(Gh, H_HIST, T),
            hist_y_loc=hy[0].reshape(Gh, H_HIST, M), hist_y_scale=hy[1].reshape(Gh, H_HIST, M))
        l_loc, l_scale, l_draw = stack(rec[0::3], 1)
        u_loc, u_scale, u_draw = stack(rec[1::3], T)
        y_loc, y_scale, y_draw = stack(rec[2::3], M)
        out.update(
            l_loc=l_loc.reshape(G, ns), l_scale=l_scale.reshape(G, ns), l_draw=l_draw.reshape(G, ns),
            u_loc=u_loc.reshape(G, ns, T), u_scale=u_scale.reshape(G, ns, T), u_draw=u_draw.reshape(G, ns, T),
            y_loc=y_loc.reshape(G, ns, M), y_scale=y_scale.reshape(G, ns, M), y_draw=y_draw.reshape(G, ns, M))
        name = f"predict_N{N}_M{M}_s{seed}_h{hidx}_{kind}"
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), N=N, M=M, x=x, Y=Y, pars=pars, hyper=json.dumps(hyper), grids=grids.numpy(),
            n_sample=ns, seed=seed, quantiles=q, mean=mean, std=std, smooth=smooth, cov=cov, test_quantiles=tq,
            test_mean=tmean, test_std=tstd, torch_version=torch.__version__, threads=torch.get_num_threads(), **out)
        names.append(name)
        print(name, q.shape, mean.shape, smooth.shape, cov.shape, float(np.abs(mean).max()), float(l_scale.min()), float(u_scale.min()))
    with open(os.path.join(HERE, "MANIFEST_predict.json"), "w") as f:
        json.dump({"cases": names, "torch": torch.__version__, "generator": "tests/golden/make_golden_predict.py",
                   "reference": "Corleno/Nonstationary_Multivariate_Gaussian_Process Utility/prediction.py:1038-1262"}, f, indent=1)


if __name__ == "__main__":
    main()
