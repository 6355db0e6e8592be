"""Generate golden input/output vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports `Utility.logpos` from /root/reference (read-only), installs the one compatibility
shim the reference needs on torch >= 2 (`torch.symeig` was removed),
evaluates `nlogpos_obj_S`, `nlogpos_obj`, `nlogpos_obj_SVC` with verbose=True, calls
`.backward()` the way the drivers do, and writes one .npz per case next to this file:
inputs (x, Y, pars, hyper), outputs (vals = [-logpost, loglik, priors...], grad) and the
torch version / thread count that produced them.
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
from Utility import logpos  # noqa: E402  (the reference)

from nonstationary_multivariate_gaussian_process_b200 import synth  # noqa: E402

HYPER = {
    # driver dictionaries: Stationary_model.py:79, Stationary_model_mpisim.py:86,
    # Separable_model_mpisim.py:296-297, Nonseparable_model_mpisim.py:311-312, and the function defaults
    "stationary": [
        {"mu_tilde_l": 0.0, "sigma_tilde_l": 1.0, "a": 1.0, "b": 1.0, "c": 10.0},
        {"mu_tilde_l": 0.0, "sigma_tilde_l": 10.0, "a": 1e-6, "b": 1e-6, "c": 1.0},
    ],
    "separable": [
        {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_tilde_sigma": 0.0,
         "alpha_tilde_sigma": 1.0, "beta_tilde_sigma": 1.0, "a": 1e-2, "b": 1e-2, "c": 0.1},
        {"mu_tilde_l": 0.0, "alpha_tilde_l": 1.0, "beta_tilde_l": 1.0, "mu_tilde_sigma": 0.0,
         "alpha_tilde_sigma": 1.0, "beta_tilde_sigma": 1.0, "a": 1.0, "b": 1.0, "c": 10.0},
    ],
    "nonseparable": [
        {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0,
         "beta_L": 1.0, "a": 1e-2, "b": 1e-2},
        {"mu_tilde_l": 0.0, "alpha_tilde_l": 5.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 5.0,
         "beta_L": 1.0, "a": 1.0, "b": 1.0},
    ],
}
FN = {"stationary": logpos.nlogpos_obj_S, "separable": logpos.nlogpos_obj, "nonseparable": logpos.nlogpos_obj_SVC}

# (model, N, M, seed, noise, hyper index, Prior)
CASES = [
    ("stationary", 50, 2, 0, 0.1, 0, True),      # BASELINE config 1 shape
    ("stationary", 50, 2, 1, 0.0, 1, True),
    ("stationary", 17, 3, 2, 0.1, 0, True),
    ("stationary", 33, 4, 3, 0.1, 1, False),
    ("separable", 12, 2, 0, 0.1, 0, True),
    ("separable", 30, 3, 1, 0.0, 0, True),
    ("separable", 64, 5, 2, 0.1, 1, True),
    ("separable", 200, 5, 3, 0.1, 0, True),      # BASELINE config 2 shape
    ("separable", 200, 5, 4, 0.0, 0, True),
    ("separable", 45, 4, 5, 0.1, 1, False),
    ("nonseparable", 8, 2, 0, 0.1, 0, True),
    ("nonseparable", 21, 3, 1, 0.1, 1, True),
    ("nonseparable", 40, 4, 2, 0.0, 0, True),
    ("nonseparable", 100, 6, 3, 0.1, 0, True),   # BASELINE config 4 per-subject shape
    ("nonseparable", 100, 6, 4, 0.0, 0, True),
    ("nonseparable", 100, 6, 5, 0.02, 1, True),
    ("nonseparable", 60, 10, 6, 0.1, 0, True),   # config 3 outputs, reduced N
    ("nonseparable", 150, 8, 7, 0.05, 0, True),  # config 5 outputs, reduced N (n=1200)
    ("nonseparable", 33, 5, 8, 0.1, 1, False),
]


def main():
    torch.set_num_threads(8)
    manifest = []
    for model, N, M, seed, noise, hidx, prior in CASES:
        x, Y, _ = synth.sample_subject(N, M, seed)
        pars = synth.start_point(model, N, M, seed, noise)
        hyper = HYPER[model][hidx]
        p = torch.from_numpy(pars).clone().requires_grad_(True)
        if model == "stationary" and not prior:
            # logpos_S leaves the prior components undefined when Prior=False (logpos.py:444-462):
            # verbose=True would raise; take the scalar path.
            out = (FN[model](p, torch.from_numpy(Y), torch.from_numpy(x), verbose=False, Prior=False, **hyper),)
        else:
            out = FN[model](p, torch.from_numpy(Y), torch.from_numpy(x), verbose=True, Prior=prior, **hyper)
        out[0].backward()
        vals = np.array([float(o) for o in out])
        name = f"{model}_N{N}_M{M}_s{seed}_h{hidx}_{'p' if prior else 'np'}"
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), model=model, N=N, M=M, x=x, Y=Y, pars=pars,
            hyper=json.dumps(hyper), prior=prior, vals=vals, grad=p.grad.numpy(),
            torch_version=torch.__version__, threads=torch.get_num_threads())
        manifest.append(name)
        print(name, vals[:2], float(np.abs(p.grad.numpy()).max()))
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump({"cases": manifest, "torch": torch.__version__, "generator": "tests/golden/make_golden.py",
                   "reference": "Corleno/Nonstationary_Multivariate_Gaussian_Process Utility/logpos.py"}, f, indent=1)


if __name__ == "__main__":
    main()
