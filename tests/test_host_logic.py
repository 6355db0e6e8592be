"""CPU-only tests of the host-side mirror of the reference interface: parameter slicing and transforms
(Utility/logpos.py:17-57, Utility/utils.py:10-88), hyper-parameter marshalling, subject partitioning, the synthetic
generator against the oracle's covariance."""
import numpy as np
import pytest
import torch

from oracle import nmgp_oracle as O
from nonstationary_multivariate_gaussian_process_b200 import batched, logpos, sharding, synth, utils


@pytest.mark.parametrize("M", [1, 2, 3, 6, 10])
def test_uL_transforms_round_trip_and_match_oracle(M):
    T = M * (M + 1) // 2
    rng = np.random.RandomState(M)
    u = rng.standard_normal(T)
    L = utils.uLvec2Lvec(u, M)
    assert np.allclose(utils.Lvec2uLvec(L, M), u, rtol=0, atol=1e-14)
    Lt = utils.uLvec2Lvec(torch.from_numpy(u), M)
    assert torch.equal(Lt, O.unconstrained_to_tril_vec(torch.from_numpy(u), M))
    mat = utils.vec2lowtriangle(Lt, M)
    assert torch.equal(mat, O.tril_vec_to_matrix(Lt, M))
    assert torch.equal(utils.lowtriangle2vec(mat, M), Lt)
    # diagonal slots are cumsum(1..M)-1 (utils.py:12)
    d = np.cumsum(np.arange(1, M + 1)) - 1
    assert np.all(np.asarray(L)[d] > 0)
    N = 4
    us = rng.standard_normal(N * T)
    assert np.allclose(utils.Lvecs2uLvecs(utils.uLvecs2Lvecs(us, N, M), N, M), us, atol=1e-14)


def test_vec2lowtriangle_checks_size():
    with pytest.raises(ValueError):
        utils.vec2lowtriangle(np.zeros(5), 3)


def test_parameter_slicing_matches_reference_layout():
    N, M = 7, 3
    T = 6
    p = torch.arange(2 * N + T + 1, dtype=torch.float64)
    tl, ts, uL, e = logpos.vec2pars(p, N, M)
    assert tl.tolist() == list(range(N)) and ts.tolist() == list(range(N, 2 * N))
    assert uL.tolist() == list(range(2 * N, 2 * N + T)) and float(e) == 2 * N + T
    p = torch.arange(N + N * T + 1, dtype=torch.float64)
    tl, uLs, e = logpos.vec2pars_SVC(p, N, M)
    assert uLs.numel() == N * T and float(uLs[0]) == N and float(e) == N + N * T
    p = torch.arange(T + 3, dtype=torch.float64)
    tl, ts, uL, e = logpos.vec2pars_S(p, M)
    assert float(tl) == 0 and float(ts) == 1 and uL.numel() == T and float(e) == T + 2
    for model in ("stationary", "separable", "nonseparable"):
        assert batched.n_params(model, N, M) == O.n_params(model, N, M)


def test_hyper_vector_order_defaults_and_errors():
    v = batched.hyper_vector("nonseparable", {"alpha_L": 1.0, "a": 1e-2})
    assert v.tolist() == [0.0, 5.0, 1.0, 0.0, 1.0, 1.0, 1e-2, 1.0, 0.0]     # logpos.py:299 defaults
    v = batched.hyper_vector("separable", {})
    assert v.tolist() == [0.0, 1.0, 1.0, 0.0, 1.0, 1.0, 1.0, 1.0, 10.0]      # logpos.py:216 defaults
    with pytest.raises(TypeError):
        batched.hyper_vector("stationary", {"a": 1})                         # mu_tilde_l / sigma_tilde_l are required
    with pytest.raises(TypeError):
        batched.hyper_vector("nonseparable", {"c": 3})                       # not a keyword of nlogpos_obj_SVC


@pytest.mark.parametrize("S,world", [(10000, 8), (10, 3), (5, 8), (0, 2), (256, 8)])
def test_shard_range_partitions_subjects(S, world):
    spans = [sharding.shard_range(S, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == S
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(S, world, world)


def test_local_summary_counts_failures_instead_of_summing_them():
    vals = torch.tensor([[1.0, 2, 3, 4, 5, 6], [float("nan")] * 6, [10.0, 20, 30, 40, 50, 60]], dtype=torch.float64)
    info = torch.tensor([0, 17, 0], dtype=torch.int32)
    s = sharding.local_summary(vals, info)
    assert s.tolist() == [11.0, 22.0, 33.0, 44.0, 55.0, 66.0, 1.0, 3.0]
    assert sharding.all_reduce_summary(s)["n_failed"] == 1.0                 # no process group: identity


def test_synthetic_covariance_matches_oracle_up_to_the_ordering_permutation():
    N, M = 9, 3
    x, tl, uL, ts2 = synth.truth(N, M, 3)
    K_tm = synth.dense_cov_time_major(x, tl, uL, M)
    K_om = O.nonseparable_cov(torch.from_numpy(x), torch.from_numpy(tl), torch.from_numpy(uL.reshape(-1)), M).numpy()
    perm = np.arange(N * M).reshape(N, M).T.reshape(-1)      # output-major position -> time-major index
    assert np.allclose(K_tm[np.ix_(perm, perm)], K_om, rtol=1e-13, atol=1e-15)
    assert np.all(np.diff(x) > 0) and x.min() > 0 and x.max() < 1
    xs, Y, pars = synth.sample_subject(N, M, 3)
    assert Y.shape == (N, M) and pars.shape == (O.n_params("nonseparable", N, M),)


def test_driver_side_helpers():
    """Splits and scores the drivers call around the hot path (Utility/utils.py:91-199), numpy only."""
    import numpy as np
    from scipy.stats import norm
    from nonstationary_multivariate_gaussian_process_b200 import utils
    rng = np.random.RandomState(3)
    x, Y = np.sort(rng.rand(40)), rng.randn(40, 3)
    x_tr, x_te, Y_tr, Y_te = utils.data_split(x, Y)
    assert x_tr.shape == (30,) and x_te.shape == (10,) and np.all(np.diff(x_tr) > 0) and np.all(np.diff(x_te) > 0)
    assert all(np.array_equal(Y[np.searchsorted(x, v)], row) for v, row in zip(x_te, Y_te))
    a, b, c, d = utils.data_split_extrapolation(x, Y, size=7)
    assert a.shape == (33,) and np.array_equal(b, x[-7:]) and np.array_equal(d, Y[-7:])
    indx, y = np.repeat(np.arange(4), 10).astype(float), rng.randn(40)
    xt, xs, it, is_, yt, ys = utils.data_split_non_chunk(x, indx, y, chunk_size=0.2, fix=True)
    assert xt.shape == (32,) and xs.shape == (8,) and np.array_equal(np.unique(is_), np.arange(4.0))
    for m in range(4):      # the held-out samples of an output are contiguous in its own ordering
        pos = np.searchsorted(x[indx == m], xs[is_ == m])
        assert np.array_equal(pos, np.arange(pos[0], pos[0] + 2))
    m_, s_, y_ = rng.randn(5, 3), rng.rand(5, 3) + 0.1, rng.randn(5, 3)
    assert abs(utils.LPD(m_, s_, y_) - norm.logpdf(y_, loc=m_, scale=s_).mean()) < 1e-14
    assert np.allclose(utils.RMSE(m_, y_, axis=0) ** 2, utils.MSE(m_, y_, axis=0))


def test_empirical_initialisation_matches_the_reference():
    """local_estimation / global_estimation / SV against outputs of the unmodified reference
    (tests/golden/empirical_*.npz from tests/golden/make_golden_empirical.py).  The semivariograms are bit-identical; the
    fits go through the same SciPy routine, whose Levenberg-Marquardt iteration stops at ftol = xtol = 1.49e-8 -- the fitted
    parameters (and what is averaged from them) are held to 1e-7, the covariance estimates to rounding."""
    import glob, os
    import numpy as np
    from conftest import GOLDEN_DIR
    from nonstationary_multivariate_gaussian_process_b200 import empirical_estimation as ee
    files = sorted(glob.glob(os.path.join(GOLDEN_DIR, "empirical_*.npz")))
    assert len(files) >= 3
    for f in files:
        g = np.load(f)
        x, Y, M = g["x"], g["Y"], int(g["M"])
        lag, sv = ee.SV(x[:9], Y[:9], M - 1)
        assert np.array_equal(lag, g["sv_lag"]) and np.array_equal(sv, g["sv_val"])      # bit-identical semivariogram
        out = ee.local_estimation(x, Y, window_size=int(g["window_size"]))
        for k, name in enumerate(("est_sigmas", "est_ls", "smooth_ls", "est_stds", "est_R", "est_B", "est_L_vecs")):
            assert out[k].shape == g[name].shape, (f, name)
            tol = 1e-7 if k < 3 else 1e-13
            assert np.abs(out[k] - g[name]).max() <= tol * np.abs(g[name]).max(), (f, name)
        assert out[7] == int(g["est_tilde_sigma2_err"]) == -4
        S, L_vec = ee.global_estimation(x, Y)
        assert np.allclose(S, g["global_S"], rtol=1e-14) and np.allclose(np.asarray(L_vec), g["global_L_vec"], rtol=1e-13)


def test_group_by_shape_keeps_subject_order():
    from nonstationary_multivariate_gaussian_process_b200.batched import group_by_shape
    g = group_by_shape([(30, 3), (45, 2), (30, 3), (64, 4), (45, 2), (30, 3)])
    assert list(g) == [(30, 3), (45, 2), (64, 4)]
    assert g[(30, 3)] == [0, 2, 5] and g[(45, 2)] == [1, 4] and g[(64, 4)] == [3]
    assert group_by_shape([]) == {}


def test_hyper_adam_matches_torch_adam_and_keeps_positive_values_positive():
    """HyperAdam on (mu, log alpha): the same trajectory as torch.optim.Adam on those coordinates; untied keys untouched."""
    import math
    from nonstationary_multivariate_gaussian_process_b200.sharding import HyperAdam
    hyper = {"mu_L": 0.3, "alpha_L": 2.0, "beta_L": 1.0}
    opt = HyperAdam(hyper, ("mu_L", "alpha_L"), lr=0.05)
    mu = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    la = torch.tensor(math.log(2.0), dtype=torch.float64, requires_grad=True)
    ref = torch.optim.Adam([mu, la], lr=0.05)
    for _ in range(25):
        # objective f = (mu - 1)^2 + (alpha - 0.5)^2 + alpha * mu
        g = {"mu_L": 2 * (hyper["mu_L"] - 1) + hyper["alpha_L"], "alpha_L": 2 * (hyper["alpha_L"] - 0.5) + hyper["mu_L"],
             "beta_L": 123.0}
        hyper = opt.step(hyper, g)
        ref.zero_grad()
        f = (mu - 1) ** 2 + (torch.exp(la) - 0.5) ** 2 + torch.exp(la) * mu
        f.backward()
        ref.step()
    assert abs(hyper["mu_L"] - float(mu)) < 1e-12 and abs(hyper["alpha_L"] - float(torch.exp(la))) < 1e-12
    assert hyper["beta_L"] == 1.0 and hyper["alpha_L"] > 0
    with pytest.raises(KeyError):
        HyperAdam(hyper, ("nope",), lr=0.1)
    nan = opt.step(hyper, {"mu_L": float("nan"), "alpha_L": float("nan")})
    assert nan["mu_L"] == hyper["mu_L"] and nan["alpha_L"] == hyper["alpha_L"]


def test_prediction_draws_consume_the_random_stream_like_one_tensor_per_draw():
    """prediction._draw_interleaved (in-place normal_() on views) against the reference's pattern -- a fresh tensor per
    `Normal(loc, scale).sample()` call, per new input and sample (Utility/prediction.py:1104-1169) -- under the same seed."""
    import time
    import torch
    from nonstationary_multivariate_gaussian_process_b200 import prediction as P
    f64 = torch.float64
    for G, ns, T, M in [(7, 5, 21, 6), (3, 4, 3, 2), (4, 3, 55, 10), (5, 2, 15, 5), (5, 2, 16, 16)]:
        torch.manual_seed(1234)
        ref_l, ref_u, ref_y = torch.empty((G, ns), dtype=f64), torch.empty((G, ns, T), dtype=f64), torch.empty((G, ns, M), dtype=f64)
        for gi in range(G):
            for s in range(ns):
                ref_l[gi, s] = torch.empty((), dtype=f64).normal_()
                ref_u[gi, s] = torch.empty(T, dtype=f64).normal_()
                ref_y[gi, s] = torch.empty(M, dtype=f64).normal_()
        tail_ref = torch.rand(1)
        torch.manual_seed(1234)
        z_l, z_u, z_y = torch.empty((G, ns), dtype=f64), torch.empty((G, ns, T), dtype=f64), torch.empty((G, ns, M), dtype=f64)
        P._draw_interleaved(z_l, z_u, z_y)
        assert torch.equal(z_l, ref_l) and torch.equal(z_u, ref_u) and torch.equal(z_y, ref_y)
        assert torch.equal(torch.rand(1), tail_ref)          # the generator is left in the same state
    # two-buffer and one-buffer modes (pred_cov / pred_smoothness)
    torch.manual_seed(5)
    a = torch.stack([torch.empty((), dtype=f64).normal_() for _ in range(12)]).reshape(4, 3)
    torch.manual_seed(5)
    b = torch.empty((4, 3), dtype=f64)
    P._draw_interleaved(b)
    assert torch.equal(a, b)
    # the drivers' shape: 201 new inputs x 100 samples (Nonseparable_model.py:377-399)
    z = [torch.empty((201, 100), dtype=f64), torch.empty((201, 100, 21), dtype=f64), torch.empty((201, 100, 6), dtype=f64)]
    t0 = time.perf_counter()
    P._draw_interleaved(*z)
    assert time.perf_counter() - t0 < 2.0
