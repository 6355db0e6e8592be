"""Function-level golden vectors of the reference's stand-alone helpers (build container only: needs /root/reference).

    python tests/golden/make_golden_units.py   ->  tests/golden/units_<seed>.npz

Every array is the output of an UNMODIFIED reference function on seeded inputs stored beside it:
`Utility/kernels.py:5-73` (pairwise_distances, RBF_cov, Nonstationary_RBF_cov), `kronecker_operation.py:5-85`
(kronecker_product, kronecker_product_diag, kron_mv, kron_inv, kron_logdet), `distributions.py:10-134`
(multivariate_normal_logpdf / 0 / 1 / 2, inverse_gamma_logpdf(_u), gamma_logpdf), `utils.py:10-88` (parameter transforms),
`logpos.py:17-71` (vec2pars*), `:111-118` (generate_K_index_SVC), `:189-213` (deviance, deviance_obj), and the dense
nonseparable covariance `K + sigma2 I` assembled by lines `logpos.py:339-352`.  The three self-consistency identities the
reference prints in its `__main__` blocks (`kronecker_operation.py:112-115`, `distributions.py:163-168`,
`logpos.py:439-441`) are recorded as the differences the reference itself attains.
"""
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
from Utility import distributions, kernels, kronecker_operation, logpos, settings, utils  # noqa: E402  (the reference)

from nonstationary_multivariate_gaussian_process_b200 import synth  # noqa: E402


def T_(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float64))


def make(seed, N1, N2, M):
    rng = np.random.RandomState(seed)
    out = {"seed": seed, "N1": N1, "N2": N2, "M": M, "jitter": settings.jitter, "precision": settings.precision}
    x1, x2 = np.sort(rng.rand(N1)), np.sort(rng.rand(N2))
    ell1, ell2 = np.exp(0.5 * rng.randn(N1) - 2.0), np.exp(0.5 * rng.randn(N2) - 2.0)
    sig1, sig2 = np.exp(0.3 * rng.randn(N1)), np.exp(0.3 * rng.randn(N2))
    alpha, beta = 1.7, 0.6
    out.update(x1=x1, x2=x2, ell1=ell1, ell2=ell2, sig1=sig1, sig2=sig2, alpha=alpha, beta=beta)
    X1, X2 = T_(x1).view(-1, 1), T_(x2).view(-1, 1)
    # ---- kernels.py
    out["pd_self"] = kernels.pairwise_distances(X1).numpy()
    out["pd_cross"] = kernels.pairwise_distances(X1, X2).numpy()
    out["rbf_self"] = kernels.RBF_cov(X1, alpha=alpha, beta=beta).numpy()
    out["rbf_cross"] = kernels.RBF_cov(X1, X2, alpha=alpha, beta=beta).numpy()
    out["gibbs_self"] = kernels.Nonstationary_RBF_cov(X1, sigma1=T_(sig1), ell1=T_(ell1)).numpy()
    out["gibbs_self_nosigma"] = kernels.Nonstationary_RBF_cov(X1, ell1=T_(ell1)).numpy()
    out["gibbs_cross"] = kernels.Nonstationary_RBF_cov(X1, T_(sig1), T_(ell1), X2, T_(sig2), T_(ell2)).numpy()
    # ---- kronecker_operation.py
    Lb = np.tril(rng.randn(M, M)) + 2.0 * np.eye(M)
    B = Lb @ Lb.T
    K = out["gibbs_self"]
    y = rng.randn(M * N1)
    mu = 0.1 * rng.randn(M * N1)
    sigma2 = 0.37
    d1, d2 = np.exp(rng.randn(M)), np.exp(rng.randn(N1))
    Brect, Krect = rng.randn(3, M), rng.randn(5, N1)
    out.update(B=B, y=y, mu=mu, sigma2=sigma2, d1=d1, d2=d2, Brect=Brect, Krect=Krect)
    out["kron"] = kronecker_operation.kronecker_product(T_(B), T_(K)).numpy()
    out["kron_rect"] = kronecker_operation.kronecker_product(T_(Brect), T_(Krect)).numpy()
    out["kron_diag"] = kronecker_operation.kronecker_product_diag(T_(d1), T_(d2)).numpy()
    out["kron_mv"] = kronecker_operation.kron_mv(T_(B), T_(K), T_(y)).numpy()
    out["kron_mv_rect"] = kronecker_operation.kron_mv(T_(Brect), T_(Krect), T_(y)).numpy()
    out["kron_inv"] = kronecker_operation.kron_inv(torch.tensor(sigma2, dtype=torch.float64), T_(B), T_(K)).numpy()
    out["kron_logdet"] = float(kronecker_operation.kron_logdet(torch.tensor(sigma2, dtype=torch.float64), T_(B), T_(K)))
    # identity (i): kron_mv == mv(kronecker_product)  (kronecker_operation.py:112-115)
    out["identity_kron_mv"] = float(np.abs(out["kron_mv"] - out["kron"] @ y).max())
    # ---- distributions.py
    s2 = torch.tensor(sigma2, dtype=torch.float64)
    out["logpdf"] = float(distributions.multivariate_normal_logpdf(T_(y), T_(mu), torch.tensor(out["kron_logdet"], dtype=torch.float64), T_(out["kron_inv"])))
    out["logpdf0"] = float(distributions.multivariate_normal_logpdf0(T_(y), T_(mu), T_(B), T_(K), s2))
    torch.manual_seed(seed)
    out["logpdf1"] = float(distributions.multivariate_normal_logpdf1(T_(y), T_(mu), T_(B), T_(K), s2))
    out["logpdf2"] = float(distributions.multivariate_normal_logpdf2(T_(y), T_(mu), T_(B), T_(K), s2))
    # identities (ii) distributions.py:163-168 and (iii) logpos.py:439-441, as attained by the reference itself
    out["identity_logpdf0_vs_dense"] = abs(out["logpdf0"] - out["logpdf"])
    out["identity_logpdf0_vs_logpdf2"] = abs(out["logpdf0"] - out["logpdf2"])
    xs = np.exp(rng.randn(6))
    out["ig_x"] = xs
    out["ig_u"] = distributions.inverse_gamma_logpdf_u(T_(xs), 2.5, 0.7).numpy()
    out["ig"] = np.asarray(distributions.inverse_gamma_logpdf(T_(xs), 2.5, 0.7))
    out["gamma"] = np.asarray(distributions.gamma_logpdf(T_(xs), 2.5, 0.7))
    # ---- utils.py
    T = M * (M + 1) // 2
    uL = rng.randn(T)
    uLs = rng.randn(N1 * T)
    out.update(uL=uL, uLs=uLs)
    out["uLvec2Lvec"] = utils.uLvec2Lvec(T_(uL), M).numpy()
    out["Lvec2uLvec"] = utils.Lvec2uLvec(T_(out["uLvec2Lvec"]), M).numpy()
    out["uLvecs2Lvecs"] = utils.uLvecs2Lvecs(T_(uLs), N1, M).numpy()
    out["Lvecs2uLvecs"] = utils.Lvecs2uLvecs(T_(out["uLvecs2Lvecs"]), N1, M).numpy()
    out["vec2lowtriangle"] = utils.vec2lowtriangle(T_(out["uLvec2Lvec"]), M).numpy()
    out["lowtriangle2vec"] = utils.lowtriangle2vec(T_(out["vec2lowtriangle"]), M).numpy()
    # ---- logpos.py helpers
    Lf = [utils.vec2lowtriangle(T_(out["uLvecs2Lvecs"][n * T:(n + 1) * T]), M) for n in range(N1)]
    out["K_index_SVC"] = logpos.generate_K_index_SVC(Lf).numpy()
    p_svc = np.concatenate([np.log(ell1), uLs, [-1.3]])
    out["pars_svc"] = p_svc
    tl, ul, ts = logpos.vec2pars_SVC(T_(p_svc), N1, M)
    out["vec2pars_SVC"] = np.concatenate([tl.numpy(), ul.numpy(), [float(ts)]])
    # dense nonseparable covariance K + sigma2_err I, lines logpos.py:339-352 (output-major ordering)
    l = torch.exp(tl)
    K_x = kernels.Nonstationary_RBF_cov(X1, ell1=l)
    K_i = logpos.generate_K_index_SVC(Lf)
    neworder = torch.arange(N1 * M).view([N1, M]).t().contiguous().view(-1)
    K_i = K_i[:, neworder][neworder]
    Kd = kronecker_operation.kronecker_product(torch.ones([M, M]).type(settings.torchType), K_x) * K_i
    out["svc_cov"] = (Kd + torch.exp(ts) * torch.eye(N1 * M).type(settings.torchType)).numpy()
    # deviance / deviance_obj (logpos.py:189-213): separable layout [tilde_l, tilde_sigma, L_vec (NOT exponentiated), tilde_sigma2_err]
    Y = rng.randn(N1, M)
    Lvec = out["uLvec2Lvec"]
    p_dev = np.concatenate([np.log(ell1), np.log(sig1), Lvec, [-1.3]])
    out.update(Y=Y, pars_dev=p_dev)
    out["deviance_obj"] = float(logpos.deviance_obj(T_(p_dev), T_(Y), T_(x1)))
    out["deviance"] = float(logpos.deviance(T_(np.log(ell1)), T_(np.log(sig1)), T_(Lvec), torch.tensor(-1.3, dtype=torch.float64),
                                            T_(Y), T_(x1)))
    return out


def main():
    for seed, N1, N2, M in [(0, 37, 23, 4), (1, 70, 9, 2), (2, 48, 64, 6)]:
        d = make(seed, N1, N2, M)
        np.savez_compressed(os.path.join(HERE, f"units_s{seed}_N{N1}_M{M}.npz"), **d)
        print(seed, N1, N2, M, "identities:", d["identity_kron_mv"], d["identity_logpdf0_vs_dense"], d["identity_logpdf0_vs_logpdf2"],
              "logpdf0/1:", d["logpdf0"], d["logpdf1"], "deviance", d["deviance"])


if __name__ == "__main__":
    main()
