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


def reference_eval(model, pars, Y, x, hyper, prior):
    """(vals, grad) of the UNMODIFIED reference objective: verbose tuple + .backward() as the drivers call it."""
    p = torch.from_numpy(pars).clone().requires_grad_(True)
    if model == "stationary" and not prior:
        # logpos_S leaves the prior components undefined when Prior=False (logpos.py:444-462):
        # verbose=True would raise; take the scalar path.
        out = (FN[model](p, torch.from_numpy(Y), torch.from_numpy(x), verbose=False, Prior=False, **hyper),)
    else:
        out = FN[model](p, torch.from_numpy(Y), torch.from_numpy(x), verbose=True, Prior=prior, **hyper)
    out[0].backward()
    return np.array([float(o) for o in out]), p.grad.numpy().copy()


def case_name(model, N, M, seed, hidx, prior):
    return f"{model}_N{N}_M{M}_s{seed}_h{hidx}_{'p' if prior else 'np'}"


def write_case(name, model, N, M, x, Y, pars, hyper, prior, vals, grad, **extra):
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), model=model, N=N, M=M, x=x, Y=Y, pars=pars,
        hyper=json.dumps(hyper), prior=prior, vals=vals, grad=grad,
        torch_version=torch.__version__, threads=torch.get_num_threads(), **extra)


def add_noprior(name):
    """Prior=True fixtures also carry the reference's value and gradient with Prior=False (same inputs): the
    likelihood part of the gradient, which the CUDA path must match to 1e-9 on EVERY fixture whatever the
    conditioning of the GP-prior terms.  Existing arrays are kept byte for byte."""
    path = os.path.join(HERE, name + ".npz")
    z = np.load(path, allow_pickle=False)
    d = {k: z[k] for k in z.files}
    if not bool(d["prior"]) or "grad_noprior" in d:
        return False
    model = str(d["model"])
    hyper = json.loads(str(d["hyper"]))
    vals, grad = reference_eval(model, d["pars"], d["Y"], d["x"], hyper, False)
    d["val_noprior"], d["grad_noprior"] = vals[0], grad
    np.savez_compressed(path, **d)
    print(name, "+ Prior=False:", float(vals[0]), float(np.abs(grad).max()))
    return True


def load_manifest():
    with open(os.path.join(HERE, "MANIFEST.json")) as f:
        return json.load(f)


def save_manifest(m):
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(m, f, indent=1)


def main():
    """Existing fixtures are never regenerated (LAPACK results move at the 1e-15 level with the thread count and, for the
    ill-conditioned prior terms, at 1e-9): only missing files are written; `--extend` adds the Prior=False arrays."""
    torch.set_num_threads(8)
    manifest = load_manifest() if os.path.exists(os.path.join(HERE, "MANIFEST.json")) else {
        "cases": [], "torch": torch.__version__, "generator": "tests/golden/make_golden.py",
        "reference": "Corleno/Nonstationary_Multivariate_Gaussian_Process Utility/logpos.py"}
    for model, N, M, seed, noise, hidx, prior in CASES:
        name = case_name(model, N, M, seed, hidx, prior)
        if not os.path.exists(os.path.join(HERE, name + ".npz")):
            x, Y, _ = synth.sample_subject(N, M, seed)
            pars = synth.start_point(model, N, M, seed, noise)
            hyper = HYPER[model][hidx]
            vals, grad = reference_eval(model, pars, Y, x, hyper, prior)
            write_case(name, model, N, M, x, Y, pars, hyper, prior, vals, grad)
            print(name, vals[:2], float(np.abs(grad).max()))
        if name not in manifest["cases"]:
            manifest["cases"].append(name)
        if "--extend" in sys.argv:
            add_noprior(name)
    save_manifest(manifest)


if __name__ == "__main__":
    main()
