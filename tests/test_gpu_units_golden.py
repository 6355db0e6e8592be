"""GPU: the stand-alone mirrors of `Utility/kernels.py`, `kronecker_operation.py`, `distributions.py`, `utils.py` and the
`logpos.py` helpers against function-level golden vectors generated from the UNMODIFIED reference
(tests/golden/make_golden_units.py), plus the C-ABI covariance entry points against the oracle and the three
self-consistency identities the reference prints in its `__main__` blocks."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, max_rel

pytestmark = pytest.mark.gpu

UNITS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "units_*.npz")))
TOL_ELEMENTWISE = 1e-14      # same operation order as the reference: covariance entries, Kronecker products
TOL_SUMS = 1e-12             # short inner products (kron_mv, gram)
TOL_SOLVE = 1e-10            # inverse / log-determinant / densities of sigma2 I + B (x) K (cond ~ 1e3)


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))


@pytest.mark.parametrize("name", UNITS)
def test_kernels_mirror(name, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import kernels
    u = _load(name)
    X1, X2 = T(u["x1"]).view(-1, 1), T(u["x2"]).view(-1, 1)
    assert max_rel(kernels.pairwise_distances(X1), u["pd_self"]) < TOL_ELEMENTWISE
    assert max_rel(kernels.pairwise_distances(X1, X2), u["pd_cross"]) < TOL_ELEMENTWISE
    a, b = float(u["alpha"]), float(u["beta"])
    assert max_rel(kernels.RBF_cov(X1, alpha=a, beta=b), u["rbf_self"]) < TOL_ELEMENTWISE
    assert max_rel(kernels.RBF_cov(X1, X2, alpha=a, beta=b), u["rbf_cross"]) < TOL_ELEMENTWISE
    assert max_rel(kernels.Nonstationary_RBF_cov(X1, sigma1=T(u["sig1"]), ell1=T(u["ell1"])), u["gibbs_self"]) < TOL_ELEMENTWISE
    assert max_rel(kernels.Nonstationary_RBF_cov(X1, ell1=T(u["ell1"])), u["gibbs_self_nosigma"]) < TOL_ELEMENTWISE
    got = kernels.Nonstationary_RBF_cov(X1, T(u["sig1"]), T(u["ell1"]), X2, T(u["sig2"]), T(u["ell2"]))
    assert max_rel(got, u["gibbs_cross"]) < TOL_ELEMENTWISE
    # CPU tensors in -> CPU tensors out (prediction.py / sim.py call these on CPU tensors); CUDA stays on the device
    assert got.device.type == "cpu"
    assert kernels.RBF_cov(X1.cuda(), alpha=a, beta=b).is_cuda


@pytest.mark.parametrize("name", UNITS)
def test_cabi_cov_entry_points_match_the_oracle(name, cuda_device):
    """nmgp_rbf_cov / nmgp_gibbs_cov / nmgp_nonseparable_cov against oracle.rbf_cov / gibbs_cov / nonseparable_cov and the
    reference's own matrices (value of Sigma to <= 1e-14)."""
    from nonstationary_multivariate_gaussian_process_b200 import _lib
    from oracle import nmgp_oracle as O
    lib = _lib.load_library()
    u = _load(name)
    N, M = int(u["N1"]), int(u["M"])
    x = T(u["x1"]).cuda()
    out = torch.empty((N, N), dtype=torch.float64, device="cuda")
    _lib.check(lib.nmgp_rbf_cov(x.data_ptr(), N, None, 0, float(u["alpha"]), float(u["beta"]), out.data_ptr(), None), "rbf")
    assert max_rel(out.cpu(), O.rbf_cov(T(u["x1"]), float(u["alpha"]), float(u["beta"]))) < TOL_ELEMENTWISE
    assert max_rel(out.cpu(), u["rbf_self"]) < TOL_ELEMENTWISE
    ell, sig = T(u["ell1"]).cuda(), T(u["sig1"]).cuda()
    _lib.check(lib.nmgp_gibbs_cov(x.data_ptr(), sig.data_ptr(), ell.data_ptr(), N, None, None, None, 0, out.data_ptr(), None), "gibbs")
    assert max_rel(out.cpu(), O.gibbs_cov(T(u["x1"]), T(u["ell1"]), T(u["sig1"])).numpy()) < TOL_ELEMENTWISE
    assert max_rel(out.cpu(), u["gibbs_self"]) < TOL_ELEMENTWISE
    pars = T(u["pars_svc"]).cuda()
    cov = torch.empty((N * M, N * M), dtype=torch.float64, device="cuda")
    _lib.check(lib.nmgp_nonseparable_cov(x.data_ptr(), pars.data_ptr(), 1, N, M, cov.data_ptr(), None), "svc cov")
    assert max_rel(cov.cpu(), u["svc_cov"]) < TOL_ELEMENTWISE                       # logpos.py:339-352, output-major
    Tt = M * (M + 1) // 2
    p = T(u["pars_svc"])
    oc = O.nonseparable_cov(T(u["x1"]), p[:N], p[N:N + N * Tt], M) + torch.exp(p[-1]) * torch.eye(N * M, dtype=torch.float64)
    assert max_rel(cov.cpu(), oc.numpy()) < TOL_ELEMENTWISE


@pytest.mark.parametrize("name", UNITS)
def test_kronecker_mirror(name, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import kronecker_operation as K
    u = _load(name)
    B, Kx, y, s2 = T(u["B"]), T(u["gibbs_self"]), T(u["y"]), float(u["sigma2"])
    assert max_rel(K.kronecker_product(B, Kx), u["kron"]) < TOL_ELEMENTWISE
    assert max_rel(K.kronecker_product(T(u["Brect"]), T(u["Krect"])), u["kron_rect"]) < TOL_ELEMENTWISE
    assert max_rel(K.kronecker_product_diag(T(u["d1"]), T(u["d2"])), u["kron_diag"]) < TOL_ELEMENTWISE
    assert max_rel(K.kron_mv(B, Kx, y), u["kron_mv"]) < TOL_SUMS
    assert max_rel(K.kron_mv(T(u["Brect"]), T(u["Krect"]), y), u["kron_mv_rect"]) < TOL_SUMS
    assert max_rel(K.kron_inv(torch.tensor(s2, dtype=torch.float64), B, Kx), u["kron_inv"]) < TOL_SOLVE
    ld = K.kron_logdet(torch.tensor(s2, dtype=torch.float64), B, Kx)
    assert abs(float(ld) - float(u["kron_logdet"])) / abs(float(u["kron_logdet"])) < TOL_SOLVE
    # identity (i), kronecker_operation.py:112-115: kron_mv(B, K, y) == mv(kronecker_product(B, K), y)
    lhs = K.kron_mv(B, Kx, y)
    rhs = torch.mv(K.kronecker_product(B, Kx), y)
    assert max_rel(lhs, rhs) < TOL_SUMS
    assert K.kron_mv(B.cuda(), Kx.cuda(), y.cuda()).is_cuda


@pytest.mark.parametrize("name", UNITS)
def test_distributions_mirror(name, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import distributions as D
    from nonstationary_multivariate_gaussian_process_b200 import kronecker_operation as K
    u = _load(name)
    B, Kx, y, mu = T(u["B"]), T(u["gibbs_self"]), T(u["y"]), T(u["mu"])
    s2 = torch.tensor(float(u["sigma2"]), dtype=torch.float64)
    rel = lambda a, b: abs(float(a) - float(b)) / abs(float(b))
    dense = D.multivariate_normal_logpdf(y, mu, T(u["kron_logdet"]), T(u["kron_inv"]))
    assert rel(dense, u["logpdf"]) < 1e-13
    v0 = D.multivariate_normal_logpdf0(y, mu, B, Kx, s2)
    v2 = D.multivariate_normal_logpdf2(y, mu, B, Kx, s2)
    assert rel(v0, u["logpdf0"]) < TOL_SOLVE
    assert rel(v2, u["logpdf2"]) < TOL_SOLVE
    # the reference's "robust" variant adds unseeded uniform jitter of size 1e-6 to the diagonals (distributions.py:66-70):
    # its value moves by O(1e-5) relative between two calls of the reference itself
    assert rel(D.multivariate_normal_logpdf1(y, mu, B, Kx, s2), u["logpdf1"]) < 1e-4
    # identity (ii), distributions.py:163-168: logpdf0 == logpdf(kron_logdet, kron_inv)
    inv, ld = K.kron_inv(s2, B, Kx), K.kron_logdet(s2, B, Kx)
    assert rel(D.multivariate_normal_logpdf(y, mu, ld, inv), v0) < TOL_SOLVE
    # identity (iii), logpos.py:439-441: logpdf0 ("woodbury") ~ logpdf1 ("robust") ~ logpdf2 ("real")
    assert rel(v0, v2) < TOL_SOLVE
    xs = T(u["ig_x"])
    assert max_rel(D.inverse_gamma_logpdf_u(xs, 2.5, 0.7), u["ig_u"]) < 1e-14
    assert max_rel(D.inverse_gamma_logpdf(xs, 2.5, 0.7), u["ig"]) < 1e-14
    assert max_rel(D.gamma_logpdf(xs, 2.5, 0.7), u["gamma"]) < 1e-14


@pytest.mark.parametrize("name", UNITS)
def test_utils_and_logpos_helpers(name, cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import logpos, utils
    u = _load(name)
    N, M = int(u["N1"]), int(u["M"])
    Tt = M * (M + 1) // 2
    assert max_rel(utils.uLvec2Lvec(T(u["uL"]), M), u["uLvec2Lvec"]) < 1e-15
    assert max_rel(utils.Lvec2uLvec(T(u["uLvec2Lvec"]), M), u["Lvec2uLvec"]) < 1e-15
    assert max_rel(utils.uLvecs2Lvecs(T(u["uLs"]), N, M), u["uLvecs2Lvecs"]) < 1e-15
    assert max_rel(utils.Lvecs2uLvecs(T(u["uLvecs2Lvecs"]), N, M), u["Lvecs2uLvecs"]) < 1e-15
    assert np.array_equal(utils.vec2lowtriangle(T(u["uLvec2Lvec"]), M).numpy(), u["vec2lowtriangle"])
    assert np.array_equal(utils.lowtriangle2vec(T(u["vec2lowtriangle"]), M).numpy(), u["lowtriangle2vec"])
    tl, ul, ts = logpos.vec2pars_SVC(T(u["pars_svc"]), N, M)
    assert np.array_equal(np.concatenate([tl.numpy(), ul.numpy(), [float(ts)]]), u["vec2pars_SVC"])
    Lf = [utils.vec2lowtriangle(T(u["uLvecs2Lvecs"][n * Tt:(n + 1) * Tt]), M) for n in range(N)]
    assert max_rel(logpos.generate_K_index_SVC(Lf), u["K_index_SVC"]) < TOL_SUMS
    # deviance / deviance_obj (logpos.py:189-213)
    Y, x, p = T(u["Y"]), T(u["x1"]), T(u["pars_dev"])
    rel = lambda a, b: abs(float(a) - float(b)) / abs(float(b))
    assert rel(logpos.deviance_obj(p, Y, x), u["deviance_obj"]) < TOL_SOLVE
    tl, tsg, Lv, ts2 = logpos.vec2pars(p, N, M)
    assert rel(logpos.deviance(tl, tsg, Lv, ts2, Y, x), u["deviance"]) < TOL_SOLVE
    logpos.clear_plan_cache()


def test_sym_eig_entry_point(cuda_device):
    from nonstationary_multivariate_gaussian_process_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.RandomState(3)
    for M in (1, 2, 5, 10, 16):
        A = rng.randn(M, M)
        B = A @ A.T + 0.1 * np.eye(M)
        Bd = T(B).cuda()
        lam = torch.empty(M, dtype=torch.float64, device="cuda")
        V = torch.empty((M, M), dtype=torch.float64, device="cuda")
        _lib.check(lib.nmgp_sym_eig(Bd.data_ptr(), M, lam.data_ptr(), V.data_ptr(), None), "eig")
        w = np.linalg.eigvalsh(B)
        assert np.abs(lam.cpu().numpy() - w).max() / w.max() < 1e-13                 # ascending, like torch.symeig
        Vn = V.cpu().numpy()
        assert np.abs(Vn @ np.diag(lam.cpu().numpy()) @ Vn.T - B).max() / np.abs(B).max() < 1e-13
        assert np.abs(Vn.T @ Vn - np.eye(M)).max() < 1e-13


def test_not_positive_definite_is_nan_with_a_warning(cuda_device):
    """Failure contract of the drop-in objectives: NaN value and gradient + NmgpNotPositiveDefinite (logpos._single)."""
    from nonstationary_multivariate_gaussian_process_b200 import logpos, synth
    N, M = 20, 2
    x, Y, _ = synth.sample_subject(N, M, 0)
    p = synth.start_point("nonseparable", N, M, 0, 0.0)
    p[0] = 800.0                   # ell_0 = exp(800) = inf: row 0 of K_x is NaN -> the first pivot is not a positive number
    pt = torch.from_numpy(p).requires_grad_(True)
    with pytest.warns(logpos.NmgpNotPositiveDefinite):
        v = logpos.nlogpos_obj_SVC(pt, torch.from_numpy(Y), torch.from_numpy(x), Prior=False)
    assert torch.isnan(v)
    logpos.clear_plan_cache()
