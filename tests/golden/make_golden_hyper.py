"""Golden vectors for the hyper-parameter gradient (nmgp_hyper_grad), produced by the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_hyper.py

The reference keeps the hyper-parameters fixed (Nonseparable_model_mpisim.py:311-312) and never differentiates with
respect to them, but its objectives are plain torch code: passing the keyword hyper-parameters as float64 leaf tensors
with requires_grad=True and calling `.backward()` on the returned -log posterior gives d(-logpost)/d(hyper) through the
reference's own arithmetic (RBF_cov, MultivariateNormal, Normal).  Two keywords cannot be tensors because
`inverse_gamma_logpdf` (distributions.py:126-134) applies numpy / scipy functions to them:
  b: the objective is affine in  -b/sigma2 + a log b ... -> central difference of the reference itself (h = 1e-3 b);
  a: -log(sigma2) + log b - digamma(a), written out with scipy.special.digamma (recorded as analytic in the fixture).
Each fixture holds several subjects of one shape, so the test also covers the sum over subjects.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch
from scipy.special import digamma

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

torch.symeig = lambda A, eigenvectors=False, upper=True: torch.linalg.eigh(A, UPLO="U" if upper else "L")
from Utility import logpos  # noqa: E402  (the reference)

from nonstationary_multivariate_gaussian_process_b200 import synth  # noqa: E402
from nonstationary_multivariate_gaussian_process_b200.batched import HYPER_SPEC  # noqa: E402

FN = {"stationary": logpos.nlogpos_obj_S, "separable": logpos.nlogpos_obj, "nonseparable": logpos.nlogpos_obj_SVC}

# (model, N, M, seeds, noise, hyper)
CASES = [
    ("stationary", 30, 3, (0, 1, 2), 0.1, {"mu_tilde_l": 0.5, "sigma_tilde_l": 2.0, "a": 1.5, "b": 0.5, "c": 10.0}),
    ("stationary", 50, 2, (3, 4), 0.1, {"mu_tilde_l": 0.0, "sigma_tilde_l": 10.0, "a": 1e-2, "b": 1e-2, "c": 1.0}),
    ("separable", 24, 2, (0, 1, 2), 0.1,
     {"mu_tilde_l": 0.3, "alpha_tilde_l": 2.0, "beta_tilde_l": 0.5, "mu_tilde_sigma": -0.2, "alpha_tilde_sigma": 1.5,
      "beta_tilde_sigma": 0.7, "a": 2.0, "b": 1.0, "c": 3.0}),
    ("separable", 64, 4, (3, 4), 0.05,
     {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_tilde_sigma": 0.0, "alpha_tilde_sigma": 1.0,
      "beta_tilde_sigma": 1.0, "a": 1e-2, "b": 1e-2, "c": 0.1}),
    ("nonseparable", 20, 2, (0, 1, 2), 0.1,
     {"mu_tilde_l": -0.5, "alpha_tilde_l": 2.0, "beta_tilde_l": 0.4, "mu_L": 0.2, "alpha_L": 1.5, "beta_L": 0.6,
      "a": 2.0, "b": 1.0}),
    ("nonseparable", 40, 3, (3, 4, 5, 6), 0.05,
     {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0,
      "a": 1e-2, "b": 1e-2}),
    ("nonseparable", 100, 6, (7, 8), 0.02,
     {"mu_tilde_l": 0.0, "alpha_tilde_l": 5.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 5.0, "beta_L": 1.0,
      "a": 1.0, "b": 1.0}),
]


def reference_hyper_grad(model, pars, Y, x, hyper):
    names = [k for k, _ in HYPER_SPEC[model]]
    leaves = {k: torch.tensor(float(hyper[k]), dtype=torch.float64, requires_grad=True) for k in names if k not in ("a", "b")}
    kw = dict(leaves, a=hyper["a"], b=hyper["b"])
    p, Yt, xt = torch.from_numpy(pars), torch.from_numpy(Y), torch.from_numpy(x)
    val = FN[model](p, Yt, xt, verbose=False, Prior=True, **kw)
    val.backward()
    out = np.zeros(9)
    for i, k in enumerate(names):
        if k in leaves:
            out[i] = float(leaves[k].grad)
    hb = 1e-3 * hyper["b"]
    plain = {k: float(hyper[k]) for k in names}
    vp = float(FN[model](p, Yt, xt, verbose=False, Prior=True, **dict(plain, b=hyper["b"] + hb)))
    vm = float(FN[model](p, Yt, xt, verbose=False, Prior=True, **dict(plain, b=hyper["b"] - hb)))
    out[names.index("b")] = (vp - vm) / (2 * hb)
    s2 = float(np.exp(pars[-1]))
    out[names.index("a")] = -(-np.log(s2) + np.log(hyper["b"]) - digamma(hyper["a"]))
    return out, float(val)


def main():
    torch.set_num_threads(8)
    manifest = []
    for model, N, M, seeds, noise, hyper in CASES:
        xs, Ys, ps, hg, vals = [], [], [], [], []
        for s in seeds:
            x, Y, _ = synth.sample_subject(N, M, s)
            pars = synth.start_point(model, N, M, s, noise)
            g, v = reference_hyper_grad(model, pars, Y, x, hyper)
            xs.append(x), Ys.append(Y), ps.append(pars), hg.append(g), vals.append(v)
        name = f"hyper_{model}_N{N}_M{M}_s{seeds[0]}"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), model=model, N=N, M=M, x=np.stack(xs), Y=np.stack(Ys),
                            pars=np.stack(ps), hyper=json.dumps(hyper), hgrad=np.stack(hg), vals=np.array(vals),
                            torch_version=torch.__version__, threads=torch.get_num_threads())
        manifest.append(name)
        print(name, np.stack(hg).sum(0))
    with open(os.path.join(HERE, "MANIFEST_hyper.json"), "w") as f:
        json.dump({"cases": manifest, "torch": torch.__version__, "generator": "tests/golden/make_golden_hyper.py",
                   "reference": "Corleno/Nonstationary_Multivariate_Gaussian_Process Utility/logpos.py (autograd through the "
                                "keyword hyper-parameters)"}, f, indent=1)


if __name__ == "__main__":
    main()
