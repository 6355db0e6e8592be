"""GPU parity: the CUDA path (through the C ABI) against the reference's golden vectors and the CPU oracle."""
import numpy as np
import pytest
import torch

import glob
import os

from conftest import GOLDEN_DIR, big_cases, golden_cases, load_golden, prior_blocks, rel_err, tolerances

pytestmark = pytest.mark.gpu

# Tolerances (relative; gradients in the 2-norm).  north_star asks 1e-9 for log-posterior and gradient.
#  * The likelihood term, everything evaluated with Prior=False and -- on EVERY fixture -- the likelihood part of the
#    gradient (the reference's own Prior=False gradient, `grad_noprior`) are held to 1e-9 (measured 1e-16..3e-13).
#  * The GP-prior terms involve alpha^2 RBF + 1e-6 I with cond 1e8..2e10.  Two *CPU* FP64 Cholesky
#    algorithms on bit-identical matrices already differ by ~4e-9 in the quadratic form and ~2e-7 in its gradient, and
#    the reference is that far from the exact (mpmath) answer itself (tests/test_prior_conditioning_floor.py;
#    test_prior_terms_are_as_close_to_the_exact_value_as_the_reference_is below).  Those components -- and totals /
#    gradients they dominate -- carry PER-FIXTURE bounds: 3 x the largest error measured over all engines on B200
#    (tests/golden/MANIFEST.json "tolerances", written by tools/update_tolerances.py from tools/report_parity.py), never
#    below 1e-9; a regression by a factor of 3 on any one fixture fails.
TOL_LOGLIK = 1e-9
TOL_NOPRIOR = 1e-9
ENGINES = ["auto", "left", "left_stable", "recursive"]


def run_plan(g, engine="auto", prior=None):
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    plan = LogPosteriorPlan(g["model"], g["x"], g["Y"], g["hyper"], prior=g["prior"] if prior is None else prior)
    plan.set_engine(engine)          # a FRESH plan per engine: no state left behind by the other factorisation path
    vals, grad, info = plan.value_and_grad_host(torch.from_numpy(g["pars"]))
    plan.close()
    return vals.numpy()[0], grad.numpy()[0], int(info[0])


def check_against_golden(name, engine):
    g = load_golden(name)
    vals, grad, info = run_plan(g, engine)
    ref = g["vals"]
    assert info == 0
    if not g["prior"]:
        assert rel_err(vals[0], ref[0]) < TOL_NOPRIOR, (name, "total", vals[0], ref[0])
        assert rel_err(grad, g["grad"]) < TOL_NOPRIOR, (name, "grad", rel_err(grad, g["grad"]))
        if len(ref) > 1:
            assert rel_err(vals[1], ref[1]) < TOL_LOGLIK
        return
    tol = tolerances(name)
    assert rel_err(vals[0], ref[0]) < tol["total"], (name, "total", vals[0], ref[0], tol)
    assert rel_err(vals[1], ref[1]) < TOL_LOGLIK, (name, "loglik", vals[1], ref[1])
    for k in range(2, len(ref)):
        assert rel_err(vals[k], ref[k]) < tol["prior"], (name, k, vals[k], ref[k], tol)
    assert rel_err(grad, g["grad"]) < tol["grad"], (name, "grad", rel_err(grad, g["grad"]), tol)
    # the likelihood part of the gradient against the reference's Prior=False gradient: north_star's 1e-9, every fixture
    v0, g0, i0 = run_plan(g, engine, prior=False)
    assert i0 == 0
    assert rel_err(v0[0], g["val_noprior"]) < TOL_NOPRIOR, (name, "Prior=False value", v0[0], g["val_noprior"])
    assert rel_err(g0, g["grad_noprior"]) < TOL_NOPRIOR, (name, "Prior=False grad", rel_err(g0, g["grad_noprior"]))


# "auto" picks the right-looking tile engine for a single subject; "left" forces the batched left-looking potrf with the
# Takahashi inverse sweep (<= 16 block columns) that the 10 000-subject sweep uses; "left_stable" the left-looking potrf
# with the W^T W inverse (PANEL_ALL / TRTRI_ROW / LAUUM) that larger matrices get; "recursive" the automatic potrf with the
# level-synchronous recursive triangular inverse (REC_T / REC_W / LAUUM) that a few large matrices get.
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", golden_cases())
def test_cuda_matches_reference_golden(name, engine, cuda_device):
    check_against_golden(name, engine)


# BASELINE.json configs[2] at full size (C3: M=10, N=500, n=5000) and the C5 shape (M=8) at n = 4096, 8192 and 16 384:
# values and gradients of the UNMODIFIED reference (tests/golden/make_golden_big.py; 7 s .. minutes of CPU each).
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", big_cases())
def test_cuda_matches_reference_golden_full_size(name, engine, cuda_device):
    check_against_golden(name, engine)


TRUTH = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "truth", "*.npz")))
# Measured on B200 (profiles/r02_parity_golden.txt, columns lp / dlp): gradients 0.9 .. 2.2 x the reference's own distance
# from the exact answer (1e-7 .. 3e-7 where the reference is at 5e-8 .. 3e-7); values 0.2 .. 9 x, all below 7e-9.
TRUTH_SLACK = 4.0


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", TRUTH)
def test_prior_terms_are_as_close_to_the_exact_value_as_the_reference_is(name, engine, cuda_device):
    """Where 1e-9-to-the-reference is not attainable (GP-prior terms, cond 1e8..2e10) the yardstick is the exact answer:
    50-digit mpmath values and gradients of both GP-prior log-densities at the fixture's float64 inputs
    (tests/golden/make_truth.py).  The CUDA path must be no further from it than TRUTH_SLACK x the reference is."""
    g = load_golden(name)
    t = np.load(os.path.join(GOLDEN_DIR, "truth", name + ".npz"))
    vals, grad, info = run_plan(g, engine)
    v0, g0, i0 = run_plan(g, engine, prior=False)
    assert info == 0 and i0 == 0
    for k, (sl, what) in enumerate(prior_blocks(g)):
        exact = -t["dlp%d" % k].reshape(-1)                       # gradient of -log prior w.r.t. the block under prior k
        ours = rel_err((grad - g0)[sl], exact)
        ref = rel_err((g["grad"] - g["grad_noprior"])[sl], exact)
        assert ours <= max(TRUTH_SLACK * ref, 1e-9), (name, what, "gradient", ours, ref)
        lp = float(t["lp%d" % k])
        ours_v, ref_v = rel_err(vals[2 + k], lp), rel_err(g["vals"][2 + k], lp)
        assert ours_v <= max(TRUTH_SLACK * ref_v, 1e-9), (name, what, "value", ours_v, ref_v)
