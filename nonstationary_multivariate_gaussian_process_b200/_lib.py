"""Loader / builder of the C-ABI CUDA library `libnmgp_b200.so` (include/nmgp_b200.h).

There is deliberately no CPU fallback: if the library is missing or no CUDA device is present, every
compute entry point of this package raises.  `build_library()` compiles the sources in `csrc/` with nvcc for
sm_100a into the package directory (in-tree, so the built .so travels to the GPU box).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libnmgp_b200.so")
# A/B timing of kernel variants (tools/build_variant.py): NMGP_B200_LIB points at another build of the SAME sources with
# different -D switches.  Not a fallback: the file must exist, and it is this package's own CUDA library either way.
if os.environ.get("NMGP_B200_LIB"):
    LIB_PATH = os.path.abspath(os.environ["NMGP_B200_LIB"])
SOURCES = ["engine.cu", "engine_ll.cu", "diag.cu", "models.cu", "predict.cu", "hadamard.cu", "hyper.cu", "kron.cu", "probe.cu", "api.cu"]
HEADERS = ["common.cuh", "engine.cuh", "models.cuh", "jacobi.cuh", os.path.join("..", "..", "include", "nmgp_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]

NVALS = 6
NHYPER = 9
NPHASES = 5
PHASE_NAMES = ("build", "potrf", "potri", "prior_solves", "contract")
STATIONARY, SEPARABLE, NONSEPARABLE = 0, 1, 2
HADAMARD, HADAMARD_SVC, HADAMARD_S = 3, 4, 5
PRED_RAW_FACTOR = 1

_lock = threading.Lock()
_lib = None


class NmgpError(RuntimeError):
    pass


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, extra_flags=(), out_path: str | None = None,
                  csrc: str | None = None) -> str:
    """Compile csrc/*.cu -> libnmgp_b200.so for sm_100a (nvcc cross-compiles without a GPU).
    `extra_flags` / `out_path` / `csrc` build a variant beside it (tools/build_variant.py)."""
    lib_path = out_path or os.path.join(_HERE, "libnmgp_b200.so")
    csrc = csrc or CSRC
    if not force and not out_path and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(_HERE, "build", os.path.basename(lib_path)[:-3] if out_path else "")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        res = subprocess.run(cmd, cwd=csrc, capture_output=True, text=True)
        if res.returncode != 0:
            raise NmgpError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:      # one nvcc per translation unit, in parallel
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", lib_path] + objs
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, cwd=csrc, capture_output=True, text=True)
    if res.returncode != 0:
        raise NmgpError("nvcc link failed:\n" + res.stdout + res.stderr)
    return lib_path


def _declare(lib):
    c = ctypes
    dp, ip, vp = c.c_void_p, c.c_void_p, c.c_void_p   # raw device/host addresses
    lib.nmgp_last_error.restype = c.c_char_p
    lib.nmgp_last_error.argtypes = []
    lib.nmgp_n_params.restype = c.c_int
    lib.nmgp_n_params.argtypes = [c.c_int, c.c_int, c.c_int]
    lib.nmgp_plan_create.restype = c.c_int
    lib.nmgp_plan_create.argtypes = [c.POINTER(c.c_void_p), c.c_int, c.c_int, c.c_int, c.c_int, dp, dp,
                                     c.POINTER(c.c_double), c.c_int, c.c_size_t, vp]
    lib.nmgp_plan_create_hadamard.restype = c.c_int
    lib.nmgp_plan_create_hadamard.argtypes = [c.POINTER(c.c_void_p), c.c_int, c.c_int, c.c_int, c.c_int, dp, ip, dp,
                                              c.POINTER(c.c_double), c.c_int, c.c_size_t, vp]
    lib.nmgp_plan_destroy.restype = c.c_int
    lib.nmgp_plan_destroy.argtypes = [c.c_void_p]
    lib.nmgp_logpost_grad.restype = c.c_int
    lib.nmgp_logpost_grad.argtypes = [c.c_void_p, dp, dp, dp, ip, vp]
    lib.nmgp_hyper_grad.restype = c.c_int
    lib.nmgp_hyper_grad.argtypes = [c.c_void_p, dp, dp, vp]
    lib.nmgp_logpost_grad_hyper.restype = c.c_int
    lib.nmgp_logpost_grad_hyper.argtypes = [c.c_void_p, dp, dp, dp, dp, ip, vp]
    lib.nmgp_sweep_reduce.restype = c.c_int
    lib.nmgp_sweep_reduce.argtypes = [dp, dp, ip, c.c_long, dp, vp]
    lib.nmgp_plan_set_hyper.restype = c.c_int
    lib.nmgp_plan_set_hyper.argtypes = [c.c_void_p, c.POINTER(c.c_double), vp]
    lib.nmgp_logpost_grad_host.restype = c.c_int
    lib.nmgp_logpost_grad_host.argtypes = [c.c_void_p, dp, dp, dp, ip, vp]
    lib.nmgp_logpost_grad_profile.restype = c.c_int
    lib.nmgp_logpost_grad_profile.argtypes = [c.c_void_p, dp, dp, dp, ip, c.POINTER(c.c_float), vp]
    lib.nmgp_plan_set_engine.restype = c.c_int
    lib.nmgp_plan_set_engine.argtypes = [c.c_void_p, c.c_int]
    lib.nmgp_plan_set_graph.restype = c.c_int
    lib.nmgp_plan_set_graph.argtypes = [c.c_void_p, c.c_int]
    lib.nmgp_plan_graph_replays.restype = c.c_long
    lib.nmgp_plan_graph_replays.argtypes = [c.c_void_p]
    lib.nmgp_plan_last_launches.restype = c.c_long
    lib.nmgp_plan_last_launches.argtypes = [c.c_void_p]
    lib.nmgp_plan_device_bytes.restype = c.c_size_t
    lib.nmgp_plan_device_bytes.argtypes = [c.c_void_p]
    lib.nmgp_plan_chunk.restype = c.c_int
    lib.nmgp_plan_chunk.argtypes = [c.c_void_p]
    lib.nmgp_plan_block.restype = c.c_int
    lib.nmgp_plan_block.argtypes = [c.c_void_p]
    lib.nmgp_adam_step.restype = c.c_int
    lib.nmgp_adam_step.argtypes = [dp, dp, dp, dp, ip, dp, c.c_long, c.c_long, c.c_double, c.c_double, c.c_double,
                                   c.c_double, c.c_long, vp]
    lib.nmgp_hmc_kick.restype = c.c_int
    lib.nmgp_hmc_kick.argtypes = [dp, dp, ip, c.c_long, c.c_long, c.c_double, vp]
    lib.nmgp_hmc_drift.restype = c.c_int
    lib.nmgp_hmc_drift.argtypes = [dp, dp, c.c_long, c.c_long, c.c_double, vp]
    lib.nmgp_hmc_accept.restype = c.c_int
    lib.nmgp_hmc_accept.argtypes = [dp, dp, dp, dp, dp, dp, dp, dp, ip, dp, ip, c.c_long, c.c_long, vp]
    lib.nmgp_predict_prior_moments.restype = c.c_int
    lib.nmgp_predict_prior_moments.argtypes = [c.c_void_p, dp, dp, c.c_int, dp, dp, dp, dp, vp]
    lib.nmgp_predict_moments.restype = c.c_int
    lib.nmgp_predict_moments.argtypes = [c.c_void_p, dp, dp, c.c_int, c.c_int, dp, dp, c.c_int, dp, dp, ip, vp]
    lib.nmgp_predict_moments_sep.restype = c.c_int
    lib.nmgp_predict_moments_sep.argtypes = [c.c_void_p, dp, dp, c.c_int, c.c_int, dp, dp, dp, dp, ip, vp]
    lib.nmgp_rbf_cov.restype = c.c_int
    lib.nmgp_rbf_cov.argtypes = [dp, c.c_int, dp, c.c_int, c.c_double, c.c_double, dp, vp]
    lib.nmgp_gibbs_cov.restype = c.c_int
    lib.nmgp_gibbs_cov.argtypes = [dp, dp, dp, c.c_int, dp, dp, dp, c.c_int, dp, vp]
    lib.nmgp_nonseparable_cov.restype = c.c_int
    lib.nmgp_nonseparable_cov.argtypes = [dp, dp, c.c_int, c.c_int, c.c_int, dp, vp]
    lib.nmgp_fp64_dmma_probe.restype = c.c_int
    lib.nmgp_fp64_dmma_probe.argtypes = [c.c_double, c.POINTER(c.c_double), c.POINTER(c.c_double), vp]
    lib.nmgp_pairwise_sqdist.restype = c.c_int
    lib.nmgp_pairwise_sqdist.argtypes = [dp, c.c_int, dp, c.c_int, dp, vp]
    lib.nmgp_kron.restype = c.c_int
    lib.nmgp_kron.argtypes = [dp, c.c_int, c.c_int, dp, c.c_int, c.c_int, dp, vp]
    lib.nmgp_kron_mv.restype = c.c_int
    lib.nmgp_kron_mv.argtypes = [dp, c.c_int, c.c_int, dp, c.c_int, c.c_int, dp, dp, dp, vp]
    lib.nmgp_gram.restype = c.c_int
    lib.nmgp_gram.argtypes = [dp, c.c_int, c.c_int, dp, vp]
    lib.nmgp_sym_eig.restype = c.c_int
    lib.nmgp_sym_eig.argtypes = [dp, c.c_int, dp, dp, vp]
    lib.nmgp_kron_eig_solve.restype = c.c_int
    lib.nmgp_kron_eig_solve.argtypes = [dp, c.c_int, dp, c.c_int, c.c_double, dp, dp, dp, ip, vp]
    lib.nmgp_potrf_batched.restype = c.c_int
    lib.nmgp_potrf_batched.argtypes = [dp, c.c_int, c.c_int, dp, ip, vp]
    lib.nmgp_potrf_potri_batched.restype = c.c_int
    lib.nmgp_potrf_potri_batched.argtypes = [dp, c.c_int, c.c_int, dp, ip, vp]
    return lib


EXPORTS = ["nmgp_last_error", "nmgp_n_params", "nmgp_plan_create", "nmgp_plan_create_hadamard", "nmgp_plan_destroy", "nmgp_logpost_grad", "nmgp_hyper_grad", "nmgp_plan_set_hyper", "nmgp_sweep_reduce", "nmgp_logpost_grad_hyper",
           "nmgp_logpost_grad_host", "nmgp_logpost_grad_profile", "nmgp_plan_set_engine", "nmgp_plan_set_graph", "nmgp_plan_graph_replays", "nmgp_plan_last_launches", "nmgp_plan_device_bytes", "nmgp_plan_chunk",
           "nmgp_plan_block", "nmgp_adam_step", "nmgp_hmc_kick", "nmgp_hmc_drift", "nmgp_hmc_accept", "nmgp_predict_prior_moments", "nmgp_predict_moments", "nmgp_predict_moments_sep", "nmgp_rbf_cov", "nmgp_gibbs_cov", "nmgp_nonseparable_cov", "nmgp_fp64_dmma_probe", "nmgp_pairwise_sqdist", "nmgp_kron", "nmgp_kron_mv", "nmgp_gram", "nmgp_sym_eig", "nmgp_kron_eig_solve", "nmgp_potrf_batched",
           "nmgp_potrf_potri_batched"]


def load_library():
    """dlopen the C-ABI library (no CUDA call is made here).  Raises NmgpError if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NmgpError(
                    f"{LIB_PATH} is missing: the CUDA library has not been built "
                    "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
            _lib = _declare(ctypes.CDLL(LIB_PATH))
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load_library().nmgp_last_error().decode("utf-8", "replace")
        raise NmgpError(f"{what} failed (code {rc}): {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise NmgpError("no CUDA device is available: this package has no CPU fallback (B200 / sm_100a only)")
    return torch
