"""CPU-only: the C-ABI shared library builds for sm_100a, loads without a GPU, and exports every function that
include/nmgp_b200.h declares (no compute call is made here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "nmgp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nmgp_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from nonstationary_multivariate_gaussian_process_b200 import _lib
    _lib.build_library()          # no-op when the in-tree .so is newer than its sources
    return _lib.load_library()


def test_header_declares_the_plan_api():
    names = declared_functions()
    for must in ("nmgp_plan_create", "nmgp_plan_destroy", "nmgp_logpost_grad", "nmgp_logpost_grad_host",
                 "nmgp_n_params", "nmgp_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_export_table_matches_header():
    from nonstationary_multivariate_gaussian_process_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_functions()


def test_n_params_needs_no_device(lib):
    # mirrors vec2pars_S / vec2pars / vec2pars_SVC (Utility/logpos.py:17-57)
    assert lib.nmgp_n_params(0, 50, 2) == 3 + 3
    assert lib.nmgp_n_params(1, 200, 5) == 2 * 200 + 15 + 1
    assert lib.nmgp_n_params(2, 100, 6) == 100 + 100 * 21 + 1
    # Hadamard variants: vec2pars / vec2pars_hadamard_SVC / vec2pars_S on N observations (logpos.py:479, 60-72, 657)
    assert lib.nmgp_n_params(3, 10, 2) == 24 and lib.nmgp_n_params(4, 10, 2) == 41 and lib.nmgp_n_params(5, 10, 2) == 6
    assert lib.nmgp_n_params(6, 10, 2) < 0 and lib.nmgp_n_params(2, 0, 2) < 0


def test_bad_arguments_return_codes_not_exceptions(lib):
    out = ctypes.c_void_p()
    hyper = (ctypes.c_double * 9)()
    rc = lib.nmgp_plan_create(ctypes.byref(out), 7, 1, 4, 2, None, None, hyper, 1, 0, None)
    assert rc == -1 and not out.value
    assert b"bad arguments" in lib.nmgp_last_error()
    assert lib.nmgp_plan_destroy(None) == 0
    assert lib.nmgp_logpost_grad(None, None, None, None, None, None) == -1


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from nonstationary_multivariate_gaussian_process_b200 import _lib, kernels, logpos
    x = torch.linspace(0, 1, 5, dtype=torch.float64)
    with pytest.raises(_lib.NmgpError):
        kernels.RBF_cov(x.view(-1, 1))
    with pytest.raises(_lib.NmgpError):
        logpos.nlogpos_obj_SVC(torch.zeros(5 + 5 * 3 + 1, dtype=torch.float64), torch.zeros(5, 2, dtype=torch.float64), x)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nonstationary_multivariate_gaussian_process_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_new_entry_points_validate_arguments_without_a_device(lib):
    """Prediction, HMC and Hadamard entries: bad arguments give NMGP_EINVAL (-1) and a message, never a crash."""
    out = ctypes.c_void_p()
    hyper = (ctypes.c_double * 9)()
    assert lib.nmgp_plan_create_hadamard(ctypes.byref(out), 2, 1, 4, 2, None, None, None, hyper, 1, 0, None) == -1
    assert b"nmgp_plan_create_hadamard" in lib.nmgp_last_error() and not out.value
    assert lib.nmgp_predict_prior_moments(None, None, None, 3, None, None, None, None, None) == -1
    assert lib.nmgp_predict_moments(None, None, None, 3, 2, None, None, 0, None, None, None, None) == -1
    assert lib.nmgp_predict_moments_sep(None, None, None, 3, 2, None, None, None, None, None, None) == -1
    assert lib.nmgp_hmc_kick(None, None, None, 1, 4, 0.1, None) == -1
    assert lib.nmgp_hmc_drift(None, None, 1, 4, 0.1, None) == -1
    assert lib.nmgp_hmc_accept(None, None, None, None, None, None, None, None, None, None, None, 1, 4, None) == -1
    assert lib.nmgp_hmc_kick(None, None, None, 0, 0, 0.1, None) == -1          # P must be positive


def test_round2_entry_points_validate_arguments_without_a_device(lib):
    """The helper entries behind the kernels / kronecker_operation / distributions mirrors and the DMMA probe: bad
    arguments give NMGP_EINVAL (-1) and a message before any CUDA call."""
    assert lib.nmgp_pairwise_sqdist(None, 3, None, 0, None, None) == -1
    assert b"nmgp_pairwise_sqdist" in lib.nmgp_last_error()
    assert lib.nmgp_kron(None, 2, 2, None, 2, 2, None, None) == -1
    assert lib.nmgp_kron_mv(None, 2, 2, None, 2, 2, None, None, None, None) == -1
    assert lib.nmgp_gram(None, 4, 2, None, None) == -1
    assert lib.nmgp_sym_eig(None, 3, None, None, None) == -1
    assert lib.nmgp_kron_eig_solve(None, 2, None, 5, 0.1, None, None, None, None, None) == -1
    tf = ctypes.c_double()
    assert lib.nmgp_fp64_dmma_probe(-1.0, ctypes.byref(tf), None, None) == -1
    assert lib.nmgp_fp64_dmma_probe(100.0, ctypes.byref(tf), None, None) == -1
    assert lib.nmgp_fp64_dmma_probe(0.1, None, None, None) == -1
