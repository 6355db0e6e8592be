"""Golden vectors for the irregularly sampled ("Hadamard") objectives, produced by running the UNMODIFIED reference
`Utility/logpos.py:465-716` (nlogpos_obj_hadamard, nlogpos_obj_hadamard_SVC, nlogpos_obj_hadamard_S) with verbose=True and
`.backward()`.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_hadamard.py

One .npz per case (`hadamard_*.npz`): x, indx, y, pars, hyper, Prior, vals = [-logpost, loglik, priors...], grad.
Data: N observation times, each assigned to one of M outputs (every output present), values from the simulation recipe.
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

from nonstationary_multivariate_gaussian_process_b200 import synth, utils  # noqa: E402

FN = {"hadamard": logpos.nlogpos_obj_hadamard, "hadamard_svc": logpos.nlogpos_obj_hadamard_SVC,
      "hadamard_s": logpos.nlogpos_obj_hadamard_S}
HYPER = {
    "hadamard": [{"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_tilde_sigma": 0.0, "alpha_tilde_sigma": 1.0,
                  "beta_tilde_sigma": 1.0, "a": 1e-2, "b": 1e-2, "c": 0.1},
                 {"mu_tilde_l": -1.0, "alpha_tilde_l": 2.0, "beta_tilde_l": 0.05, "mu_tilde_sigma": 0.2, "alpha_tilde_sigma": 1.0,
                  "beta_tilde_sigma": 0.05, "a": 1.0, "b": 1.0, "c": 10.0}],
    "hadamard_svc": [{"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0,
                      "a": 1e-2, "b": 1e-2},
                     {"mu_tilde_l": -1.0, "alpha_tilde_l": 2.0, "beta_tilde_l": 0.05, "mu_L": 0.1, "alpha_L": 1.5,
                      "beta_L": 0.05, "a": 1.0, "b": 1.0}],
    "hadamard_s": [{"mu_tilde_l": 0.0, "sigma_tilde_l": 10.0, "a": 1e-6, "b": 1e-6, "c": 1.0},
                   {"mu_tilde_l": 0.0, "sigma_tilde_l": 1.0, "a": 1.0, "b": 1.0, "c": 10.0}],
}
# (model, N, M, seed, hyper index, Prior)
CASES = [
    ("hadamard", 30, 3, 0, 1, True), ("hadamard", 70, 4, 1, 0, True), ("hadamard", 45, 2, 2, 1, False),
    ("hadamard_svc", 25, 2, 0, 1, True), ("hadamard_svc", 66, 3, 1, 0, True), ("hadamard_svc", 130, 5, 2, 1, True),
    ("hadamard_svc", 40, 4, 3, 0, False),
    ("hadamard_s", 40, 3, 0, 1, True), ("hadamard_s", 64, 2, 1, 0, True), ("hadamard_s", 90, 5, 2, 1, True),
]


def make_case(model, N, M, seed):
    """(x, indx, y, pars) -- synth.hadamard_case: one observation per time point, outputs assigned at random."""
    return synth.hadamard_case(model, N, M, seed)


def main():
    torch.set_num_threads(8)
    names = []
    for model, N, M, seed, hidx, prior in CASES:
        x, indx, y, pars = make_case(model, N, M, seed)
        hyper = HYPER[model][hidx]
        p = torch.from_numpy(pars).clone().requires_grad_(True)
        out = FN[model](p, torch.from_numpy(x), torch.from_numpy(indx), torch.from_numpy(y), verbose=True, Prior=prior, **hyper)
        out[0].backward()
        vals = np.array([float(o) for o in out])
        name = f"hadamard_{model.split('_')[-1] if '_' in model else 'sep'}_N{N}_M{M}_s{seed}_h{hidx}_{'p' if prior else 'np'}"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), model=model, N=N, M=M, x=x, indx=indx, y=y, pars=pars,
                            hyper=json.dumps(hyper), prior=prior, vals=vals, grad=p.grad.numpy(),
                            torch_version=torch.__version__, threads=torch.get_num_threads())
        names.append(name)
        print(name, vals[:2], float(np.abs(p.grad.numpy()).max()))
    with open(os.path.join(HERE, "MANIFEST_hadamard.json"), "w") as f:
        json.dump({"cases": names, "torch": torch.__version__, "generator": "tests/golden/make_golden_hadamard.py",
                   "reference": "Corleno/Nonstationary_Multivariate_Gaussian_Process Utility/logpos.py:465-716"}, f, indent=1)


if __name__ == "__main__":
    main()
