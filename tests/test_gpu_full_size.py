"""GPU: BASELINE.json's full-size configurations, checked through properties that need no CPU oracle run
(the oracle takes seconds to minutes at these sizes):
  * the likelihood term against an independent library factorisation (torch.linalg FP64 on the GPU, test-only checker) of
    the covariance the C ABI itself emits in the reference's ordering (`nmgp_nonseparable_cov`),
  * the analytic gradient against a central finite difference of the value along a random direction,
  * batch invariance: a subject evaluated inside a large batch (left-looking engine) equals the same subject evaluated
    alone (right-looking engine)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HYPER = {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0,
         "a": 1e-2, "b": 1e-2}


def _subject(N, M, seed, noise=0.02):
    from nonstationary_multivariate_gaussian_process_b200 import synth
    x, _, _, _ = synth.truth(N, M, seed)
    Y = np.random.RandomState(seed).standard_normal((N, M))
    return x, Y, synth.start_point("nonseparable", N, M, seed, noise)


def _library_loglik(x, pars, Y, N, M):
    from nonstationary_multivariate_gaussian_process_b200 import _lib
    lib = _lib.load_library()
    n = N * M
    xd, pd = torch.from_numpy(x).cuda(), torch.from_numpy(pars).cuda()
    cov = torch.empty((n, n), dtype=torch.float64, device="cuda")
    _lib.check(lib.nmgp_nonseparable_cov(xd.data_ptr(), pd.data_ptr(), 1, N, M, cov.data_ptr(), None), "cov")
    L = torch.linalg.cholesky(cov)
    y = torch.from_numpy(Y).cuda().t().reshape(-1, 1)              # output-major, logpos.py:250
    z = torch.linalg.solve_triangular(L, y, upper=False)
    return float(-torch.log(torch.diagonal(L)).sum() - 0.5 * (z * z).sum())


def _check_value_and_directional_gradient(N, M, seed, tol_val, tol_fd, eps):
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    x, Y, p = _subject(N, M, seed)
    plan = LogPosteriorPlan("nonseparable", x, Y, HYPER, prior=False)
    pt = torch.from_numpy(p).cuda().unsqueeze(0)
    vals, grad, info = plan.value_and_grad(pt)
    assert int(info[0]) == 0
    ref = _library_loglik(x, p, Y, N, M)
    assert abs(float(vals[0, 1]) - ref) / abs(ref) < tol_val, (float(vals[0, 1]), ref)
    d = torch.from_numpy(np.random.RandomState(seed + 1).standard_normal(p.shape[0])).cuda()
    d = d / d.norm()
    vp, _, _ = plan.value_and_grad(pt + eps * d, need_grad=False)
    vm, _, _ = plan.value_and_grad(pt - eps * d, need_grad=False)
    fd = float(vp[0, 0] - vm[0, 0]) / (2 * eps)
    an = float((grad[0] * d).sum())
    assert abs(fd - an) / max(abs(an), 1e-12) < tol_fd, (fd, an)
    plan.close()


def test_config3_n5000_single_subject(cuda_device):
    # BASELINE.json configs[2]: M=10, N=500, dense NM=5000 blocked FP64 Cholesky
    _check_value_and_directional_gradient(500, 10, 3, tol_val=1e-10, tol_fd=1e-5, eps=1e-5)


def test_config5_n16384_single_subject(cuda_device):
    # BASELINE.json configs[4]: M=8, N=2048, NM=16384 Cholesky per evaluation
    _check_value_and_directional_gradient(2048, 8, 4, tol_val=1e-10, tol_fd=1e-4, eps=1e-5)


def test_config4_shape_batch_invariance(cuda_device):
    # BASELINE.json configs[3] shape (M=6, N=100): 300 subjects in one batch vs. three of them alone
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    N, M, S = 100, 6, 300
    subs = [_subject(N, M, 1000 + s) for s in range(S)]
    xs, Ys, ps = (np.stack(a) for a in zip(*subs))
    plan = LogPosteriorPlan("nonseparable", xs, Ys, HYPER)
    vals, grad, info = plan.value_and_grad(torch.from_numpy(ps).cuda())
    assert int(info.abs().sum()) == 0
    for s in (0, 137, 299):
        solo = LogPosteriorPlan("nonseparable", xs[s], Ys[s], HYPER)
        v1, g1, i1 = solo.value_and_grad(torch.from_numpy(ps[s]).cuda().unsqueeze(0))
        assert int(i1[0]) == 0
        assert abs(float(v1[0, 1] - vals[s, 1])) / abs(float(v1[0, 1])) < 1e-12          # likelihood
        assert abs(float(v1[0, 0] - vals[s, 0])) / abs(float(v1[0, 0])) < 1e-12          # total (same prior factors)
        assert float((g1[0] - grad[s]).norm() / g1[0].norm()) < 1e-10
        solo.close()
    plan.close()


def test_separable_config2_against_dense_library_factorisation(cuda_device):
    # BASELINE.json configs[1]: separable, M=5, N=200.  B (x) K + s2 I formed densely by the checker only.
    from nonstationary_multivariate_gaussian_process_b200 import kernels, synth, utils
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    N, M = 200, 5
    T = M * (M + 1) // 2
    x, Y, _ = synth.sample_subject(N, M, 11)
    p = synth.start_point("separable", N, M, 11, 0.02)
    plan = LogPosteriorPlan("separable", x, Y, {}, prior=False)
    vals, grad, info = plan.value_and_grad(torch.from_numpy(p).cuda().unsqueeze(0))
    assert int(info[0]) == 0
    tl, ts, uL, te = p[:N], p[N:2 * N], p[2 * N:2 * N + T], p[-1]
    K = kernels.Nonstationary_RBF_cov(torch.from_numpy(x).view(-1, 1), torch.from_numpy(np.exp(ts)),
                                      torch.from_numpy(np.exp(tl))).cuda()
    L = utils.vec2lowtriangle(utils.uLvec2Lvec(torch.from_numpy(uL), M), M).cuda()
    Sigma = torch.kron(L @ L.t(), K) + np.exp(te) * torch.eye(N * M, dtype=torch.float64, device="cuda")
    Lc = torch.linalg.cholesky(Sigma)
    y = torch.from_numpy(Y).cuda().t().reshape(-1, 1)
    z = torch.linalg.solve_triangular(Lc, y, upper=False)
    ref = float(-torch.log(torch.diagonal(Lc)).sum() - 0.5 * (z * z).sum())
    assert abs(float(vals[0, 1]) - ref) / abs(ref) < 1e-10
    plan.close()


@pytest.mark.parametrize("N,M,S", [(256, 8, 3), (512, 8, 2)])
def test_batched_engine_is_stable_at_large_block_counts(N, M, S, cuda_device):
    """n = 2048 / 4096 (32 / 64 block columns): the batched left-looking path must agree with the tile-task path.
    (A Takahashi inverse sweep over that many block columns amplifies rounding errors geometrically -- 5e-7 at 32
    columns, garbage at 64 -- so the plan switches the inverse to W^T W beyond 16 columns.)"""
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    subs = [_subject(N, M, 500 + s) for s in range(S)]
    xs, Ys, ps = (np.stack(a) for a in zip(*subs))
    res = {}
    for mode in ("right", "left", "recursive"):
        plan = LogPosteriorPlan("nonseparable", xs, Ys, HYPER, prior=False)
        plan.set_engine(mode)
        v, g, i = plan.value_and_grad(torch.from_numpy(ps).cuda())
        assert int(i.abs().sum()) == 0
        res[mode] = (v.cpu().numpy(), g.cpu().numpy())
        plan.close()
    for mode in ("left", "recursive"):
        dv = np.abs(res[mode][0][:, 1] - res["right"][0][:, 1]) / np.abs(res["right"][0][:, 1])
        dg = np.linalg.norm(res[mode][1] - res["right"][1], axis=1) / np.linalg.norm(res["right"][1], axis=1)
        assert dv.max() < 1e-10, (mode, dv)
        assert dg.max() < 1e-8, (mode, dg)
    ref = _library_loglik(xs[0], ps[0], Ys[0], N, M)
    assert abs(res["left"][0][0, 1] - ref) / abs(ref) < 1e-10
