"""Golden vectors for posterior prediction of the SEPARABLE and STATIONARY models (second half),
produced by running the UNMODIFIED reference `Utility/prediction.py:34-459` and `:1566-1692`.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_predict_sep.py

Same harness as make_golden_predict.py (torch.symeig / torch.solve shims, a recording stand-in for
`torch.distributions.Normal` in the reference module's namespace).  One .npz per case.

separable (`predictsep_*.npz`):
    map_percentiles [G,3,M]                              pointwise_predmap            (plug-in, no sampling)
    quantiles / mean / std, l_*, s_*, y_* [G,ns,...]     pointwise_predmap_sampling   (seed 1000+s): loc, scale, draw of
                                                         tilde_l*, tilde_sigma*, y per (new input, sample)
    hist_pars [H+1,P], hist_y [Gh,H,M], hist_*           pointwise_predsample         (seed 5000+s), last H entries
stationary (`predictstat_*.npz`):
    map_percentiles [G,3,M], test_mean / test_std [G,M]  pointwise_predmap_S, test_predmap_S
    hist_pars [H,P], hist_y [H,G,M]                      pointwise_predsample_S with np.random.seed(7000+s)
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
    # Separable_model_mpisim.py:296-297 (ill-conditioned prior covariances)
    {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_tilde_sigma": 0.0, "alpha_tilde_sigma": 1.0,
     "beta_tilde_sigma": 1.0},
    # short prior length scales: well-conditioned prior covariances
    {"mu_tilde_l": -1.0, "alpha_tilde_l": 2.0, "beta_tilde_l": 0.03, "mu_tilde_sigma": 0.3, "alpha_tilde_sigma": 0.8,
     "beta_tilde_sigma": 0.04},
]
# (N, M, seed, noise, hyper index, G, n_sample)
SEP_CASES = [(20, 2, 0, 0.1, 1, 4, 5), (45, 3, 1, 0.1, 1, 5, 4), (64, 5, 2, 0.05, 1, 3, 6), (30, 3, 3, 0.1, 0, 4, 5)]
# (N, M, seed, G)
STAT_CASES = [(30, 3, 0, 5), (50, 2, 1, 7), (70, 4, 2, 4)]
H_HIST, G_HIST = 3, 3


def stack(records, width):
    loc = np.stack([np.broadcast_to(r[0].numpy(), (width,)) for r in records])
    sc = np.stack([np.broadcast_to(r[1].numpy(), (width,)) for r in records])
    dr = np.stack([r[2].numpy() for r in records])
    return loc, sc, dr


class quiet:
    def __enter__(self):
        self.old = sys.stdout
        sys.stdout = open(os.devnull, "w")

    def __exit__(self, *a):
        sys.stdout.close()
        sys.stdout = self.old


def main():
    torch.set_num_threads(8)
    names = []
    for N, M, seed, noise, hidx, G, ns in SEP_CASES:
        T = M * (M + 1) // 2
        x, Y, _ = synth.sample_subject(N, M, seed)
        pars = synth.start_point("separable", N, M, seed, noise)
        hyper = HYPER[hidx]
        pt = torch.from_numpy(pars)
        tl, tsg, uL, te = pt[:N].clone(), pt[N:2 * N].clone(), pt[2 * N:2 * N + T].clone(), pt[-1].clone()
        xt, Yt = torch.from_numpy(x), torch.from_numpy(Y)
        grids = torch.linspace(0.03, 0.97, G, dtype=torch.float64)
        with quiet():
            RECORD.clear()
            mp = prediction.pointwise_predmap(tl, tsg, uL, te, Yt, xt, grids, **hyper)
            RECORD.clear()
            torch.manual_seed(1000 + seed)
            q, mean, std = prediction.pointwise_predmap_sampling(ns, tl, tsg, uL, te, Yt, xt, grids, **hyper)
            rec = list(RECORD)
            hist = np.stack([synth.start_point("separable", N, M, 100 + seed + h, noise) for h in range(H_HIST + 1)])
            ht = torch.from_numpy(hist)
            RECORD.clear()
            torch.manual_seed(5000 + seed)
            hy = prediction.pointwise_predsample(ht[:, :N], ht[:, N:2 * N], ht[:, 2 * N:2 * N + T], ht[:, -1], Yt, xt,
                                                 grids[:G_HIST], N_sample=H_HIST, **hyper)
            hrec = list(RECORD)
        assert len(rec) == 3 * G * ns and len(hrec) == 3 * G_HIST * H_HIST
        l, s, yv = stack(rec[0::3], 1), stack(rec[1::3], 1), stack(rec[2::3], M)
        hl, hs, hyv = stack(hrec[0::3], 1), stack(hrec[1::3], 1), stack(hrec[2::3], M)
        name = f"predictsep_N{N}_M{M}_s{seed}_h{hidx}"
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), N=N, M=M, x=x, Y=Y, pars=pars, hyper=json.dumps(hyper), grids=grids.numpy(),
            n_sample=ns, seed=seed, map_percentiles=mp.numpy(), quantiles=q, mean=mean, std=std,
            l_loc=l[0].reshape(G, ns), l_scale=l[1].reshape(G, ns), l_draw=l[2].reshape(G, ns),
            s_loc=s[0].reshape(G, ns), s_scale=s[1].reshape(G, ns), s_draw=s[2].reshape(G, ns),
            y_loc=yv[0].reshape(G, ns, M), y_scale=yv[1].reshape(G, ns, M), y_draw=yv[2].reshape(G, ns, M),
            hist_pars=hist, hist_n_sample=H_HIST, hist_grids=grids[:G_HIST].numpy(), hist_y=hy,
            hist_l_loc=hl[0].reshape(G_HIST, H_HIST), hist_l_scale=hl[1].reshape(G_HIST, H_HIST),
            hist_l_draw=hl[2].reshape(G_HIST, H_HIST), hist_s_loc=hs[0].reshape(G_HIST, H_HIST),
            hist_s_scale=hs[1].reshape(G_HIST, H_HIST), hist_s_draw=hs[2].reshape(G_HIST, H_HIST),
            hist_y_loc=hyv[0].reshape(G_HIST, H_HIST, M), hist_y_scale=hyv[1].reshape(G_HIST, H_HIST, M),
            torch_version=torch.__version__, threads=torch.get_num_threads())
        names.append(name)
        print(name, mp.shape, q.shape, hy.shape, float(np.abs(mean).max()), float(l[1].min()), float(s[1].min()))
    for N, M, seed, G in STAT_CASES:
        T = M * (M + 1) // 2
        x, Y, _ = synth.sample_subject(N, M, seed)
        hist = np.stack([synth.start_point("stationary", N, M, 200 + seed + h, 0.1) for h in range(H_HIST)])
        pars = hist[0]
        pt = torch.from_numpy(pars)
        xt, Yt = torch.from_numpy(x), torch.from_numpy(Y)
        grids = torch.linspace(0.02, 0.98, G, dtype=torch.float64)
        with quiet():
            mp = prediction.pointwise_predmap_S(pt[0], pt[1], pt[2:2 + T], pt[-1], Yt, xt, grids)
            tm, tsd = prediction.test_predmap_S(pt[0], pt[1], pt[2:2 + T], pt[-1], Yt, xt, xt[:6])
            ht = torch.from_numpy(hist)
            np.random.seed(7000 + seed)
            hy = prediction.pointwise_predsample_S(ht[:, 0], ht[:, 1], ht[:, 2:2 + T], ht[:, -1], Yt, xt, grids)
        name = f"predictstat_N{N}_M{M}_s{seed}"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), N=N, M=M, x=x, Y=Y, pars=pars, grids=grids.numpy(), seed=seed,
                            map_percentiles=mp.numpy(), test_mean=tm.numpy(), test_std=tsd.numpy(), hist_pars=hist,
                            hist_y=hy, torch_version=torch.__version__, threads=torch.get_num_threads())
        names.append(name)
        print(name, mp.shape, tm.shape, hy.shape, float(np.abs(mp.numpy()).max()))
    with open(os.path.join(HERE, "MANIFEST_predict_sep.json"), "w") as f:
        json.dump({"cases": names, "torch": torch.__version__, "generator": "tests/golden/make_golden_predict_sep.py",
                   "reference": "Corleno/Nonstationary_Multivariate_Gaussian_Process Utility/prediction.py:34-459, 1566-1692"},
                  f, indent=1)


if __name__ == "__main__":
    main()
