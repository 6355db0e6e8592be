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


def golden_cases():
    """log-posterior + gradient fixtures (tests/golden/make_golden.py)"""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith(("predict", "hadamard", "empirical", "hyper", "map_")))


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
