"""CPU: the oracle's closed-form hyper-parameter gradient (oracle/nmgp_oracle.py:hyper_grad) against the golden vectors
recorded from the unmodified reference's autograd through its keyword hyper-parameters (tests/golden/make_golden_hyper.py)."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "hyper_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["model"], d["hyper"], d["N"], d["M"] = str(d["model"]), json.loads(str(d["hyper"])), int(d["N"]), int(d["M"])
    return d


def test_fixtures_present():
    assert len(CASES) >= 7


@pytest.mark.parametrize("name", CASES)
def test_oracle_hyper_grad_matches_reference_autograd(name):
    from oracle import nmgp_oracle as O
    d = load(name)
    for s in range(d["pars"].shape[0]):
        got = O.hyper_grad(d["model"], d["pars"][s], d["x"][s], d["M"], **d["hyper"])
        want = d["hgrad"][s]
        # the prior covariances have condition numbers up to 1e10: both routes carry ~1e-6 relative noise in the
        # alpha / beta slots at the drivers' hyper-parameters
        err = np.abs(got - want) / np.maximum(np.abs(want), 1e-3 * np.abs(want).max())
        assert err.max() < 2e-5, (name, s, got, want, err)
