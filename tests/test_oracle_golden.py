"""Pins the CPU oracle (oracle/nmgp_oracle.py) to golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU-only."""
import numpy as np
import pytest

from conftest import (golden_cases, hadamard_cases, load_golden, load_hadamard_golden, load_predict_golden, max_rel, predict_cases, predictsep_cases,
                      predictstat_cases, rel_err)
from oracle import nmgp_oracle as O

# The oracle restates the reference's algorithm op-for-op, so agreement is at rounding level for the
# likelihood; the GP-prior terms are ill-conditioned (cond 1e8..1e10) and even a
# re-ordered sum moves them at the 1e-10 level.
VAL_TOL = 1e-9
GRAD_TOL = 1e-9


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    vals, grad = O.value_and_grad(g["model"], g["pars"], g["Y"], g["x"], Prior=g["prior"], **g["hyper"])
    vals = vals.numpy()
    ref = g["vals"]
    # golden for stationary Prior=False holds only the scalar objective (the reference cannot return more)
    for k in range(len(ref)):
        assert rel_err(vals[k], ref[k]) < VAL_TOL, (name, k, vals[k], ref[k])
    assert rel_err(grad.numpy(), g["grad"]) < GRAD_TOL, name
    assert grad.shape[0] == O.n_params(g["model"], g["N"], g["M"])


@pytest.mark.parametrize("name", predict_cases())
def test_predict_oracle_matches_reference_golden(name):
    """oracle/nmgp_predict_oracle.py against every draw (loc, scale, value) the unmodified reference made
    (Utility/prediction.py:1038-1262, recorded by tests/golden/make_golden_predict.py), same seed, same generator."""
    import torch
    from oracle import nmgp_predict_oracle as PO
    g = load_predict_golden(name)
    N, M = g["N"], g["M"]
    T = M * (M + 1) // 2
    pars = torch.from_numpy(g["pars"])
    args = (g["n_sample"], pars[:N], pars[N:N + N * T], pars[-1], torch.from_numpy(g["Y"]), torch.from_numpy(g["x"]),
            torch.from_numpy(g["grids"]))
    torch.manual_seed(1000 + g["seed"])
    r = PO.pointwise_predict(*args, **g["hyper"])
    # same operations as the reference; LAPACK's eigh / LU results move at the 1e-11 level with the thread count
    for k in ("l_loc", "l_scale", "l_draw", "u_loc", "u_scale", "u_draw", "y_loc", "y_scale", "y_draw"):
        assert max_rel(r[k], g[k]) < VAL_TOL, (name, k, max_rel(r[k], g[k]))
    q, mean, std = PO.summarise(r["y_draw"])
    assert max_rel(q, g["quantiles"]) < VAL_TOL and max_rel(mean, g["mean"]) < VAL_TOL and max_rel(std, g["std"]) < VAL_TOL
    torch.manual_seed(2000 + g["seed"])
    r = PO.pointwise_predict(*args, mode="smoothness", **g["hyper"])
    assert max_rel(r["l_draw"], g["smooth"]) < VAL_TOL
    torch.manual_seed(3000 + g["seed"])
    r = PO.pointwise_predict(*args, mode="cov", **g["hyper"])
    from oracle import nmgp_oracle as O
    Ls = O.tril_vec_to_matrix(O.unconstrained_to_tril_vec(torch.from_numpy(r["u_draw"]), M), M).numpy()
    assert max_rel(Ls, g["cov"]) < VAL_TOL
    # plug-in variant (prediction.py:912-1012) and posterior-sample variant (:1265-1378)
    pct, Lv = PO.pointwise_predict_plugin(*args[1:], **g["hyper"])
    assert max_rel(pct, g["map_percentiles"]) < VAL_TOL and max_rel(Lv, g["map_Lvecs"]) < VAL_TOL
    hp = torch.from_numpy(g["hist_pars"])
    torch.manual_seed(5000 + g["seed"])
    r = PO.pointwise_predict_history(hp[:, :N], hp[:, N:N + N * T], hp[:, -1], torch.from_numpy(g["Y"]),
                                     torch.from_numpy(g["x"]), torch.from_numpy(g["hist_grids"]),
                                     N_sample=int(g["hist_n_sample"]), **g["hyper"])
    for k in ("l_loc", "l_scale", "l_draw", "u_loc", "u_scale", "u_draw", "y_loc", "y_scale"):
        assert max_rel(r[k], g["hist_" + k]) < VAL_TOL, (name, "hist_" + k, max_rel(r[k], g["hist_" + k]))
    assert max_rel(r["y_draw"], g["hist_y"]) < VAL_TOL


@pytest.mark.parametrize("name", predictsep_cases())
def test_separable_predict_oracle_matches_reference_golden(name):
    """Separable-model predictors (Utility/prediction.py:34-459) restated in oracle/nmgp_predict_oracle.py."""
    import torch
    from oracle import nmgp_predict_oracle as PO
    g = load_predict_golden(name)
    N, M = g["N"], g["M"]
    T = M * (M + 1) // 2
    p = torch.from_numpy(g["pars"])
    Y, x, grids = torch.from_numpy(g["Y"]), torch.from_numpy(g["x"]), torch.from_numpy(g["grids"])
    args = (p[:N], p[N:2 * N], p[2 * N:2 * N + T], p[-1], Y, x, grids)
    assert max_rel(PO.sep_pointwise_plugin(*args, **g["hyper"]), g["map_percentiles"]) < VAL_TOL
    torch.manual_seed(1000 + g["seed"])
    r = PO.sep_pointwise_predict(g["n_sample"], *args, **g["hyper"])
    for k in r:
        assert max_rel(r[k], g[k]) < VAL_TOL, (name, k, max_rel(r[k], g[k]))
    hp = torch.from_numpy(g["hist_pars"])
    torch.manual_seed(5000 + g["seed"])
    r = PO.sep_pointwise_history(hp[:, :N], hp[:, N:2 * N], hp[:, 2 * N:2 * N + T], hp[:, -1], Y, x,
                                 torch.from_numpy(g["hist_grids"]), N_sample=int(g["hist_n_sample"]), **g["hyper"])
    for k in ("l_loc", "l_scale", "l_draw", "s_loc", "s_scale", "s_draw", "y_loc", "y_scale"):
        assert max_rel(r[k], g["hist_" + k]) < VAL_TOL, (name, "hist_" + k, max_rel(r[k], g["hist_" + k]))
    assert max_rel(r["y_draw"], g["hist_y"]) < VAL_TOL


@pytest.mark.parametrize("name", predictstat_cases())
def test_stationary_predict_oracle_matches_reference_golden(name):
    """Stationary-model predictors (Utility/prediction.py:1566-1692)."""
    import torch
    from oracle import nmgp_predict_oracle as PO
    g = load_predict_golden(name)
    M = g["M"]
    T = M * (M + 1) // 2
    p = torch.from_numpy(g["pars"])
    Y, x = torch.from_numpy(g["Y"]), torch.from_numpy(g["x"])
    mu, s2 = PO.stationary_moments(p[0], p[1], p[2:2 + T], p[-1], Y, x, torch.from_numpy(g["grids"]))
    sd = np.sqrt(s2)
    assert max_rel(np.stack([mu - 1.96 * sd, mu, mu + 1.96 * sd], axis=1), g["map_percentiles"]) < VAL_TOL
    mu, s2 = PO.stationary_moments(p[0], p[1], p[2:2 + T], p[-1], Y, x, x[:6])
    assert max_rel(mu, g["test_mean"]) < VAL_TOL and max_rel(np.sqrt(s2), g["test_std"]) < VAL_TOL
    np.random.seed(7000 + g["seed"])
    hp = torch.from_numpy(g["hist_pars"])
    ys = []
    for h in range(hp.shape[0]):
        mu, s2 = PO.stationary_moments(hp[h, 0], hp[h, 1], hp[h, 2:2 + T], hp[h, -1], Y, x, torch.from_numpy(g["grids"]))
        ys.append(np.stack([mu[i] + np.random.randn() * np.sqrt(s2[i]) for i in range(mu.shape[0])]))
    assert max_rel(np.stack(ys), g["hist_y"]) < VAL_TOL


@pytest.mark.parametrize("name", hadamard_cases())
def test_hadamard_oracle_matches_reference_golden(name):
    """Irregular-sampling objectives (Utility/logpos.py:465-716) restated in oracle/nmgp_oracle.py."""
    g = load_hadamard_golden(name)
    vals, grad = O.value_and_grad_hadamard(g["model"], g["pars"], g["x"], g["indx"], g["y"], Prior=g["prior"], **g["hyper"])
    vals = vals.numpy()
    for k in range(len(g["vals"])):
        assert rel_err(vals[k], g["vals"][k]) < VAL_TOL, (name, k, vals[k], g["vals"][k])
    assert rel_err(grad.numpy(), g["grad"]) < GRAD_TOL, name


# ---- function-level pins: the oracle's covariance builders against matrices produced by the reference's own
# kernels.py / logpos.py:339-352 (tests/golden/make_golden_units.py)
import glob as _glob
import torch
import os as _os

_UNITS = sorted(_glob.glob(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "units_*.npz")))


@pytest.mark.parametrize("path", _UNITS, ids=[_os.path.basename(p)[:-4] for p in _UNITS])
def test_oracle_covariances_match_reference_matrices(path):
    u = np.load(path)
    N, M = int(u["N1"]), int(u["M"])
    T = M * (M + 1) // 2
    x = torch.from_numpy(u["x1"])
    rel = lambda a, b: float(np.abs(np.asarray(a) - b).max() / np.abs(b).max())
    assert rel(O.sq_dist(x), u["pd_self"]) < 1e-15
    assert rel(O.rbf_cov(x, float(u["alpha"]), float(u["beta"])), u["rbf_self"]) < 1e-15
    assert rel(O.gibbs_cov(x, torch.from_numpy(u["ell1"]), torch.from_numpy(u["sig1"])), u["gibbs_self"]) < 1e-15
    p = torch.from_numpy(u["pars_svc"])
    cov = O.nonseparable_cov(x, p[:N], p[N:N + N * T], M) + torch.exp(p[-1]) * torch.eye(N * M, dtype=torch.float64)
    assert rel(cov, u["svc_cov"]) < 1e-14
    # the reference's own identities, at the level the reference attains them
    assert float(u["identity_kron_mv"]) < 1e-12
    assert float(u["identity_logpdf0_vs_dense"]) < 1e-10 and float(u["identity_logpdf0_vs_logpdf2"]) < 1e-10
    y, mu = torch.from_numpy(u["y"]), torch.from_numpy(u["mu"])
    v = O.kron_eig_loglik(y - mu, torch.from_numpy(u["B"]), torch.from_numpy(u["gibbs_self"]), torch.tensor(float(u["sigma2"]), dtype=torch.float64))
    assert abs(float(v) - float(u["logpdf0"])) / abs(float(u["logpdf0"])) < 1e-12
