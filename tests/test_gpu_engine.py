"""GPU tests of the factorisation engines through the C ABI: unit potrf/potri against LAPACK (numpy), the two engines
(right-looking tile tasks, left-looking/Takahashi) against each other, ragged sizes, failure reporting."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _spd(n, batch, seed, cond=1e3):
    rng = np.random.RandomState(seed)
    A = rng.standard_normal((batch, n, n))
    S = A @ A.transpose(0, 2, 1) / n + np.eye(n) * (1.0 / cond) * 10
    return S


@pytest.mark.parametrize("n,batch", [(1, 3), (5, 2), (63, 2), (64, 3), (65, 2), (100, 4), (200, 5), (600, 2), (333, 3)])
def test_unit_potrf_potri_match_lapack(n, batch, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import _lib
    lib = _lib.load_library()
    S = _spd(n, batch, n)
    for invert in (False, True):
        A = torch.from_numpy(S.copy()).cuda()
        logdet = torch.empty(batch, dtype=torch.float64, device="cuda")
        info = torch.empty(batch, dtype=torch.int32, device="cuda")
        fn = lib.nmgp_potrf_potri_batched if invert else lib.nmgp_potrf_batched
        _lib.check(fn(A.data_ptr(), n, batch, logdet.data_ptr(), info.data_ptr(), None), "potrf")
        torch.cuda.synchronize()
        assert int(info.abs().sum()) == 0
        ref_ld = np.linalg.slogdet(S)[1]
        assert np.allclose(logdet.cpu().numpy(), ref_ld, rtol=1e-12, atol=1e-10)
        out = A.cpu().numpy()
        if invert:
            ref = np.linalg.inv(S)
            assert np.abs(out - ref).max() / np.abs(ref).max() < 1e-11
            assert np.abs(out - out.transpose(0, 2, 1)).max() / np.abs(ref).max() < 1e-12
        else:
            ref = np.linalg.cholesky(S)
            assert np.abs(np.tril(out) - ref).max() / np.abs(ref).max() < 1e-12
            # strict upper triangle untouched
            iu = np.triu_indices(n, 1)
            assert np.array_equal(out[:, iu[0], iu[1]], S[:, iu[0], iu[1]])


def test_unit_potrf_reports_failing_pivot(cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import _lib
    lib = _lib.load_library()
    n = 100
    S = _spd(n, 3, 1)
    S[1, 70, 70] = -5.0   # not positive definite from pivot 71 at the latest
    A = torch.from_numpy(S).cuda()
    logdet = torch.empty(3, dtype=torch.float64, device="cuda")
    info = torch.empty(3, dtype=torch.int32, device="cuda")
    _lib.check(lib.nmgp_potrf_batched(A.data_ptr(), n, 3, logdet.data_ptr(), info.data_ptr(), None), "potrf")
    torch.cuda.synchronize()
    i = info.cpu().numpy()
    assert i[0] == 0 and i[2] == 0 and 1 <= i[1] <= 71
    assert np.isnan(logdet[1].item()) and np.isfinite(logdet[0].item())


@pytest.mark.parametrize("N,M,S", [(100, 6, 5), (33, 5, 4), (21, 3, 3), (64, 2, 3), (75, 8, 2)])
def test_left_and_right_looking_engines_agree(N, M, S, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import synth
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    hyper = {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0,
             "a": 1e-2, "b": 1e-2}
    xs, Ys, ps = [], [], []
    for s in range(S):
        x, Y, _ = synth.sample_subject(N, M, 100 + s)
        xs.append(x); Ys.append(Y); ps.append(synth.start_point("nonseparable", N, M, 100 + s, 0.05))
    plan = LogPosteriorPlan("nonseparable", np.stack(xs), np.stack(Ys), hyper)
    p = torch.from_numpy(np.stack(ps)).cuda()
    res = {}
    for mode in ("right", "left", "left_stable", "recursive"):
        plan.set_engine(mode)
        v, g, i = plan.value_and_grad(p)
        torch.cuda.synchronize()
        assert int(i.abs().sum()) == 0
        res[mode] = (v.cpu().numpy(), g.cpu().numpy())
    for mode in ("left", "left_stable", "recursive"):
        dv = np.abs(res[mode][0][:, 1] - res["right"][0][:, 1]) / np.abs(res["right"][0][:, 1])
        dg = np.linalg.norm(res[mode][1] - res["right"][1], axis=1) / np.linalg.norm(res["right"][1], axis=1)
        assert dv.max() < 1e-12, (mode, dv)
        assert dg.max() < 1e-11, (mode, dg)


@pytest.mark.parametrize("S", [3, 1200, 4500])
def test_a_failed_subject_does_not_poison_later_evaluations(S, cuda_device):
    """A non-finite parameter vector makes one subject's factorisation fail (info > 0, NaN outputs).  The workspace is
    reused by every later call (and by other subjects of later chunks): the next evaluation with good parameters must be
    exactly what a fresh plan returns.  n = 600 has a ragged last block, the case where stale tiles are multiplied by zeros."""
    from nonstationary_multivariate_gaussian_process_b200 import synth
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    N, M = 100, 6
    hyper = {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0,
             "a": 1e-2, "b": 1e-2}
    base = [synth.sample_subject(N, M, 300 + s)[:2] + (synth.start_point("nonseparable", N, M, 300 + s, 0.05),)
            for s in range(3)]
    xs = np.stack([base[s % 3][0] for s in range(S)])
    Ys = np.stack([base[s % 3][1] for s in range(S)])
    ps = np.stack([base[s % 3][2] for s in range(S)])
    good = torch.from_numpy(ps).cuda()
    bad = good.clone()
    bad[1, 5] = float("nan")
    bad[S - 1, -1] = float("inf")
    plan = LogPosteriorPlan("nonseparable", xs, Ys, hyper)
    v0, g0, i0 = plan.value_and_grad(good)
    vb, gb, ib = plan.value_and_grad(bad)
    assert int(ib[1]) != 0 and int(ib[S - 1]) != 0 and int((ib != 0).sum()) == 2
    ok = ib == 0
    assert torch.isnan(vb[1, 0]) and torch.equal(vb[ok], v0[ok]) and torch.equal(gb[ok], g0[ok])   # S = 4500: two chunks share the workspace
    v1, g1, i1 = plan.value_and_grad(good)
    plan.close()
    assert int(i1.abs().sum()) == 0
    assert torch.equal(v1, v0) and torch.equal(g1, g0)


@pytest.mark.parametrize("j", [0, 7, 8, 36, 63, 64, 71, 130, 199])
def test_diagonal_block_kernel_reports_the_exact_pivot(j, cuda_device):
    """Diagonal matrices: pivot j IS entry j, so info must name it exactly wherever it sits in the 8 x 8 tile grid of the
    64 x 64 diagonal-block kernel (csrc/diag.cu:diag64_mma_kernel / factor_tile); a subnormal pivot counts as a failure."""
    from nonstationary_multivariate_gaussian_process_b200 import _lib
    lib = _lib.load_library()
    n = 200
    d = np.linspace(0.5, 3.0, n)
    S = np.stack([np.diag(d), np.diag(d), np.diag(d)])
    S[0, j, j] = -1.0
    S[2, j, j] = 1e-310
    A = torch.from_numpy(S.copy()).cuda()
    logdet = torch.empty(3, dtype=torch.float64, device="cuda")
    info = torch.empty(3, dtype=torch.int32, device="cuda")
    _lib.check(lib.nmgp_potrf_batched(A.data_ptr(), n, 3, logdet.data_ptr(), info.data_ptr(), None), "potrf")
    torch.cuda.synchronize()
    assert info.cpu().tolist() == [j + 1, 0, j + 1]
    assert abs(logdet[1].item() - np.log(d).sum()) < 1e-12 * n
    assert np.allclose(np.diagonal(A[1].cpu().numpy()), np.sqrt(d), rtol=1e-15, atol=0)


def test_diagonal_block_kernel_log_determinant_of_badly_scaled_pivots(cuda_device):
    """The kernel takes one logarithm per four reciprocal pivots: pivots between 1e-70 and 1e+70 must not overflow the
    product, and the factor / inverse of a graded matrix stay accurate."""
    from nonstationary_multivariate_gaussian_process_b200 import _lib
    lib = _lib.load_library()
    n = 128
    rng = np.random.RandomState(5)
    s = 10.0 ** rng.uniform(-70, 70, size=n)
    Q = _spd(n, 1, 11)[0]
    S = (s[:, None] * Q * s[None, :])[None]
    A = torch.from_numpy(S.copy()).cuda()
    logdet = torch.empty(1, dtype=torch.float64, device="cuda")
    info = torch.empty(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.nmgp_potrf_batched(A.data_ptr(), n, 1, logdet.data_ptr(), info.data_ptr(), None), "potrf")
    torch.cuda.synchronize()
    assert int(info[0]) == 0
    ref = np.linalg.slogdet(Q)[1] + 2.0 * np.log(s).sum()
    assert abs(logdet[0].item() - ref) < 1e-11 * abs(ref) + 1e-9
    L = np.tril(A[0].cpu().numpy())
    Lq = L / s[:, None]                       # chol(D Q D) = D chol(Q)
    assert np.abs(Lq - np.linalg.cholesky(Q)).max() < 1e-11


_SWITCH_SCRIPT = r"""
import json, sys
import numpy as np, torch
from nonstationary_multivariate_gaussian_process_b200 import synth
from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
N, M, S = 150, 8, 2          # n = 1200: 19 block columns, right-looking look-ahead potrf (batch * Kt < 2048)
hyper = {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0, "a": 1e-2, "b": 1e-2}
xs, Ys, ps = [], [], []
for s in range(S):
    x, Y, _ = synth.sample_subject(N, M, 500 + s)
    xs.append(x); Ys.append(Y); ps.append(synth.start_point("nonseparable", N, M, 500 + s, 0.05))
plan = LogPosteriorPlan("nonseparable", np.stack(xs), np.stack(Ys), hyper)
plan.set_engine("right")
v, g, i = plan.value_and_grad(torch.from_numpy(np.stack(ps)).cuda())
torch.cuda.synchronize()
print(json.dumps({"v": v.cpu().numpy().tolist(), "g": g.cpu().numpy().tolist(), "info": int(i.abs().sum())}))
"""


def test_potrf_variants_behind_environment_switches_agree(cuda_device):
    """The A/B switches are read once per process, so each variant runs in its own interpreter: the tensor-pipe diagonal block
    against the FMA kernels (NMGP_DIAG_MMA), the updates of the right-looking potrf on the TMA-ring kernel against the plain tile
    kernels (NMGP_REST_LL), the 128-wide diagonal step (NMGP_DIAG128).  Same factor up to rounding: values to 1e-12, gradients
    to 1e-10."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for name, env in {"default": {}, "fma_diag": {"NMGP_DIAG_MMA": "0"}, "plain_updates": {"NMGP_REST_LL": "0"},
                      "diag128": {"NMGP_DIAG128": "1"}}.items():
        e = dict(os.environ, **env)
        e["PYTHONPATH"] = root + os.pathsep + e.get("PYTHONPATH", "")
        r = subprocess.run([sys.executable, "-c", _SWITCH_SCRIPT], env=e, cwd=root, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (name, r.stderr[-2000:])
        out[name] = json.loads(r.stdout.strip().splitlines()[-1])
        assert out[name]["info"] == 0
    v0, g0 = np.array(out["default"]["v"]), np.array(out["default"]["g"])
    for name in ("fma_diag", "plain_updates", "diag128"):
        v, g = np.array(out[name]["v"]), np.array(out[name]["g"])
        assert np.abs(v[:, 1] - v0[:, 1]).max() / np.abs(v0[:, 1]).max() < 1e-12, name
        assert (np.linalg.norm(g - g0, axis=1) / np.linalg.norm(g0, axis=1)).max() < 1e-10, name
