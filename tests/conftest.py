import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def manifest():
    with open(os.path.join(GOLDEN_DIR, "MANIFEST.json")) as f:
        return json.load(f)


def big_cases():
    """full-size fixtures (C3 n = 5000, C5 shape at n = 4096 / 8192 / 16 384; tests/golden/make_golden_big.py)"""
    return [c for c in manifest().get("big_cases", []) if os.path.exists(os.path.join(GOLDEN_DIR, c + ".npz"))]


def golden_cases():
    """log-posterior + gradient fixtures (tests/golden/make_golden.py); the full-size ones are listed by big_cases()"""
    big = set(manifest().get("big_cases", []))
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith(("predict", "hadamard", "empirical", "hyper", "map_", "units"))
                  and os.path.basename(p)[:-4] not in big)


def tolerances(name):
    """Per-fixture bounds on the prior-bearing quantities: 3 x the largest error measured over all engines on B200
    (tools/report_parity.py -> tools/update_tolerances.py), never below north_star's 1e-9.  A fixture without an entry is
    held to 1e-9 throughout."""
    t = manifest().get("tolerances", {}).get(name, {})
    return {k: max(float(t.get(k, 0.0)), 1e-9) for k in ("total", "prior", "grad")}


def prior_blocks(g):
    """[(slice of pars under GP prior k, name)] for k = 0, 1 (logpos.py:271-281 separable, :357-365 nonseparable)"""
    N, M = g["N"], g["M"]
    T = M * (M + 1) // 2
    if g["model"] == "nonseparable":
        return [(slice(0, N), "tilde_l"), (slice(N, N + N * T), "uL_vecs")]
    if g["model"] == "separable":
        return [(slice(0, N), "tilde_l"), (slice(N, 2 * N), "tilde_sigma")]
    return []


def hadamard_cases():
    """irregular-sampling ("Hadamard") objective fixtures (tests/golden/make_golden_hadamard.py)"""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "hadamard_*.npz")))


def load_hadamard_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["model"] = str(d["model"])
    d["hyper"] = json.loads(str(d["hyper"]))
    d["prior"] = bool(d["prior"])
    d["N"], d["M"] = int(d["N"]), int(d["M"])
    return d


def predict_cases():
    """posterior-prediction fixtures (tests/golden/make_golden_predict.py)"""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "predict_*.npz")))


def predictsep_cases():
    """separable-model prediction fixtures (tests/golden/make_golden_predict_sep.py)"""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "predictsep_*.npz")))


def predictstat_cases():
    """stationary-model prediction fixtures (tests/golden/make_golden_predict_sep.py)"""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "predictstat_*.npz")))


def load_predict_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    if "hyper" in d:
        d["hyper"] = json.loads(str(d["hyper"]))
    for k in ("N", "M", "n_sample", "seed"):
        if k in d:
            d[k] = int(d[k])
    return d


def max_rel(a, b):
    """max |a-b| / max |b|"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["model"] = str(d["model"])
    d["hyper"] = json.loads(str(d["hyper"]))
    d["prior"] = bool(d["prior"])
    d["N"] = int(d["N"])
    d["M"] = int(d["M"])
    return d


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) elementwise for scalars; 2-norm relative for vectors."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.ndim == 0:
        return abs(a - b) / max(abs(b), 1e-300)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
