"""Batched evaluation plans: the host-side handle on `nmgp_plan` (include/nmgp_b200.h).

A plan fixes everything the reference's MAP / HMC loops keep constant between calls -- the model, the subjects'
time stamps `x` and observations `Y`, the hyper-parameters (`**hyper_pars` of the drivers) -- and evaluates
`-log posterior`, its components and its gradient for all subjects in one call.  The single-subject functions of
`logpos.py` (reference signatures) are thin wrappers over a plan with S = 1.
"""
from __future__ import annotations

import ctypes
from typing import Mapping

import numpy as np

from . import _lib

MODELS = {"stationary": _lib.STATIONARY, "separable": _lib.SEPARABLE, "nonseparable": _lib.NONSEPARABLE,
          "hadamard": _lib.HADAMARD, "hadamard_svc": _lib.HADAMARD_SVC, "hadamard_s": _lib.HADAMARD_S}
HADAMARD_MODELS = ("hadamard", "hadamard_svc", "hadamard_s")

# keyword order and defaults of the reference signatures (Utility/logpos.py:383, :216, :299)
HYPER_SPEC = {
    "stationary": (("mu_tilde_l", None), ("sigma_tilde_l", None), ("a", 1), ("b", 1), ("c", 10)),
    "separable": (("mu_tilde_l", 0.0), ("alpha_tilde_l", 1.0), ("beta_tilde_l", 1.0), ("mu_tilde_sigma", 0.0),
                  ("alpha_tilde_sigma", 1.0), ("beta_tilde_sigma", 1.0), ("a", 1), ("b", 1), ("c", 10)),
    "nonseparable": (("mu_tilde_l", 0.0), ("alpha_tilde_l", 5.0), ("beta_tilde_l", 1.0), ("mu_L", 0.0),
                     ("alpha_L", 5.0), ("beta_L", 1.0), ("a", 1), ("b", 1)),
    # irregularly sampled variants (Utility/logpos.py:465, :561, :640): same keywords, their own defaults
    "hadamard": (("mu_tilde_l", 0.0), ("alpha_tilde_l", 1.0), ("beta_tilde_l", 1.0), ("mu_tilde_sigma", 0.0),
                 ("alpha_tilde_sigma", 1.0), ("beta_tilde_sigma", 1.0), ("a", 1), ("b", 1), ("c", 10)),
    "hadamard_svc": (("mu_tilde_l", 0.0), ("alpha_tilde_l", 1.0), ("beta_tilde_l", 1.0), ("mu_L", 0.0),
                     ("alpha_L", 1.0), ("beta_L", 1.0), ("a", 1), ("b", 1)),
    "hadamard_s": (("mu_tilde_l", None), ("sigma_tilde_l", None), ("a", 1), ("b", 1), ("c", 10)),
}
# number of entries of the reference's verbose tuple (value + components)
N_VERBOSE = {"stationary": 5, "separable": 6, "nonseparable": 5, "hadamard": 6, "hadamard_svc": 5, "hadamard_s": 5}


def hyper_vector(model: str, hyper: Mapping[str, float]) -> np.ndarray:
    spec = HYPER_SPEC[model]
    unknown = set(hyper) - {k for k, _ in spec}
    if unknown:
        raise TypeError(f"unexpected hyper-parameter(s) for the {model} model: {sorted(unknown)}")
    out = np.zeros(_lib.NHYPER, dtype=np.float64)
    for i, (k, default) in enumerate(spec):
        v = hyper.get(k, default)
        if v is None:
            raise TypeError(f"missing required hyper-parameter '{k}' for the {model} model")
        out[i] = float(v)
    return out


def n_params(model: str, N: int, M: int) -> int:
    T = M * (M + 1) // 2
    return {"stationary": T + 3, "separable": 2 * N + T + 1, "nonseparable": N + N * T + 1,
            "hadamard": 2 * N + T + 1, "hadamard_svc": N + N * T + 1, "hadamard_s": T + 3}[model]


class LogPosteriorPlan:
    """Evaluation plan for S subjects of equal shape (N time points, M outputs) on one GPU."""

    def __init__(self, model: str, x, Y, hyper: Mapping[str, float] | None = None, prior: bool = True,
                 device=None, workspace_limit_bytes: int = 0, indx=None, M: int | None = None):
        """x [S,N] (or [N]); Y [S,N,M] (or [N,M]).  Hadamard models ('hadamard', 'hadamard_svc', 'hadamard_s': one
        observation per row): Y is y [S,N] (or [N]), indx [S,N] the output index of every observation, M the number of
        outputs (default: the number of distinct indices, as the reference counts it, logpos.py:478)."""
        torch = _lib.require_cuda()
        if model not in MODELS:
            raise ValueError(f"unknown model '{model}' (expected one of {sorted(MODELS)})")
        self.model = model
        self.lib = _lib.load_library()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        x_t = torch.as_tensor(x, dtype=torch.float64)
        Y_t = torch.as_tensor(Y, dtype=torch.float64)
        had = model in HADAMARD_MODELS
        ix_t = None
        if x_t.dim() == 1:
            x_t, Y_t = x_t.unsqueeze(0), Y_t.unsqueeze(0)
        if had:
            if indx is None:
                raise ValueError("the Hadamard models need indx (the output index of every observation)")
            ix_t = torch.as_tensor(indx).to(torch.int64).reshape(x_t.shape)
            if Y_t.shape != x_t.shape:
                raise ValueError(f"expected x, indx, y of equal shape [S,N]; got {tuple(x_t.shape)} and {tuple(Y_t.shape)}")
            n_out = int(M) if M is not None else int(torch.unique(ix_t).numel())
            if int(ix_t.min()) < 0 or int(ix_t.max()) >= n_out:
                raise ValueError("indx must hold output indices 0 .. M-1")
            self.S, self.N, self.M = int(x_t.shape[0]), int(x_t.shape[1]), n_out
        else:
            if x_t.dim() != 2 or Y_t.dim() != 3 or Y_t.shape[:2] != x_t.shape:
                raise ValueError(f"expected x [S,N] and Y [S,N,M]; got {tuple(x_t.shape)} and {tuple(Y_t.shape)}")
            self.S, self.N, self.M = int(Y_t.shape[0]), int(Y_t.shape[1]), int(Y_t.shape[2])
        if self.N < 1 or self.M < 1 or self.M > 16:
            raise ValueError("need N >= 1 and 1 <= M <= 16")
        self.P = n_params(model, self.N, self.M)
        self.hyper = dict(hyper or {})
        self.prior = bool(prior)
        hv = hyper_vector(model, self.hyper)
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            xd = x_t.to(self.device).contiguous()
            Yd = Y_t.to(self.device).contiguous()
            stream = torch.cuda.current_stream(self.device).cuda_stream
            if had:
                ixd = ix_t.to(torch.int32).to(self.device).contiguous()
                rc = self.lib.nmgp_plan_create_hadamard(
                    ctypes.byref(self._handle), MODELS[model], self.S, self.N, self.M, xd.data_ptr(), ixd.data_ptr(),
                    Yd.data_ptr(), hv.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(self.prior),
                    int(workspace_limit_bytes), ctypes.c_void_p(stream))
                _lib.check(rc, "nmgp_plan_create_hadamard")
            else:
                rc = self.lib.nmgp_plan_create(
                    ctypes.byref(self._handle), MODELS[model], self.S, self.N, self.M, xd.data_ptr(), Yd.data_ptr(),
                    hv.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(self.prior), int(workspace_limit_bytes),
                    ctypes.c_void_p(stream))
                _lib.check(rc, "nmgp_plan_create")
        self._pinned = None

    # ------------------------------------------------------------------ properties
    @property
    def chunk(self) -> int:
        return self.lib.nmgp_plan_chunk(self._handle)

    @property
    def device_bytes(self) -> int:
        return self.lib.nmgp_plan_device_bytes(self._handle)

    @property
    def last_launches(self) -> int:
        return self.lib.nmgp_plan_last_launches(self._handle)

    @property
    def graph_replays(self) -> int:
        """evaluations served by a CUDA-graph replay so far (launch-bound single-chunk plans)"""
        return self.lib.nmgp_plan_graph_replays(self._handle)

    def set_graph(self, enabled: bool):
        """CUDA-graph replay of launch-bound evaluations: on (automatic, default) / off (A/B timing)."""
        _lib.check(self.lib.nmgp_plan_set_graph(self._handle, 0 if enabled else 1), "nmgp_plan_set_graph")

    def set_engine(self, mode: str):
        """'auto' | 'right' (right-looking tile tasks) | 'left' (left-looking / Takahashi): tests and A/B timing."""
        _lib.check(self.lib.nmgp_plan_set_engine(self._handle, {"auto": 0, "right": 1, "left": 2, "left_stable": 3, "recursive": 4}[mode]),
                   "nmgp_plan_set_engine")

    # ------------------------------------------------------------------ evaluation
    def value_and_grad(self, pars, need_grad: bool = True, out=None):
        """Device path: pars [S,P] CUDA float64 -> (vals [S,6], grad [S,P] or None, info [S] int32), all CUDA.
        Stream-ordered on torch's current stream; no host synchronisation.  `out` = (vals, grad, info) reuses the caller's
        buffers: a loop that evaluates into the same buffers from the same `pars` storage is replayed as one CUDA graph when
        the plan is launch-bound (nmgp_plan_set_graph)."""
        torch = _lib.require_cuda()
        if not (isinstance(pars, torch.Tensor) and pars.is_cuda):
            raise TypeError("value_and_grad expects a CUDA tensor; use value_and_grad_host for host buffers")
        p = pars.detach().to(torch.float64).reshape(self.S, self.P).contiguous()
        if out is not None:
            vals, grad, info = out
            if not need_grad:
                grad = None
            for t, shape, dt in ((vals, (self.S, _lib.NVALS), torch.float64), (grad, (self.S, self.P), torch.float64),
                                 (info, (self.S,), torch.int32)):
                if t is not None and not (isinstance(t, torch.Tensor) and t.device == self.device and t.dtype == dt
                                          and tuple(t.shape) == shape and t.is_contiguous()):
                    raise ValueError(f"out buffers must be contiguous {dt} CUDA tensors on {self.device} of shapes "
                                     f"[S,{_lib.NVALS}], [S,P], [S]")
        else:
            vals = torch.empty((self.S, _lib.NVALS), dtype=torch.float64, device=self.device)
            grad = torch.empty((self.S, self.P), dtype=torch.float64, device=self.device) if need_grad else None
            info = torch.empty((self.S,), dtype=torch.int32, device=self.device)
        if self.S == 0:
            return vals, grad, info
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.nmgp_logpost_grad(self._handle, p.data_ptr(), vals.data_ptr(),
                                            grad.data_ptr() if need_grad else None, info.data_ptr(),
                                            ctypes.c_void_p(stream))
        _lib.check(rc, "nmgp_logpost_grad")
        return vals, grad, info

    def hyper_grad(self, pars):
        """d(-log posterior)/d(hyper-parameters) of every subject: pars [S,P] CUDA -> [S,9] CUDA float64 in the keyword order of
        HYPER_SPEC[model] (unused slots 0).  The reference keeps the hyper-parameters fixed per subject
        (Nonseparable_model_mpisim.py:311-312); a caller that ties them across subjects sums this over its subjects and
        all-reduces it (sharding.all_reduce_hyper_grad)."""
        torch = _lib.require_cuda()
        if not (isinstance(pars, torch.Tensor) and pars.is_cuda):
            raise TypeError("hyper_grad expects a CUDA tensor")
        p = pars.detach().to(torch.float64).reshape(self.S, self.P).contiguous()
        out = torch.zeros((self.S, _lib.NHYPER), dtype=torch.float64, device=self.device)
        if self.S == 0:
            return out
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.nmgp_hyper_grad(self._handle, p.data_ptr(), out.data_ptr(), ctypes.c_void_p(stream))
        _lib.check(rc, "nmgp_hyper_grad")
        return out

    def value_grad_and_hyper_grad(self, pars, need_grad: bool = True):
        """value_and_grad + hyper_grad in one pass (nmgp_logpost_grad_hyper): -> (vals, grad, hgrad [S,9], info)."""
        torch = _lib.require_cuda()
        if not (isinstance(pars, torch.Tensor) and pars.is_cuda):
            raise TypeError("value_grad_and_hyper_grad expects a CUDA tensor")
        p = pars.detach().to(torch.float64).reshape(self.S, self.P).contiguous()
        vals = torch.empty((self.S, _lib.NVALS), dtype=torch.float64, device=self.device)
        grad = torch.empty((self.S, self.P), dtype=torch.float64, device=self.device) if need_grad else None
        hgrad = torch.zeros((self.S, _lib.NHYPER), dtype=torch.float64, device=self.device)
        info = torch.empty((self.S,), dtype=torch.int32, device=self.device)
        if self.S == 0:
            return vals, grad, hgrad, info
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.nmgp_logpost_grad_hyper(self._handle, p.data_ptr(), vals.data_ptr(),
                                                  grad.data_ptr() if need_grad else None, hgrad.data_ptr(),
                                                  info.data_ptr(), ctypes.c_void_p(stream))
        _lib.check(rc, "nmgp_logpost_grad_hyper")
        return vals, grad, hgrad, info

    def set_hyper(self, hyper: Mapping[str, float]):
        """Replace hyper-parameters in place (nmgp_plan_set_hyper): keywords not given keep their current value.  Only the
        prior covariances whose (alpha, beta) changed are factored again."""
        torch = _lib.require_cuda()
        new = dict(self.hyper)
        new.update(hyper)
        hv = hyper_vector(self.model, new)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.nmgp_plan_set_hyper(self._handle, hv.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                              ctypes.c_void_p(stream))
        _lib.check(rc, "nmgp_plan_set_hyper")
        self.hyper = new

    def hyper_names(self):
        return tuple(k for k, _ in HYPER_SPEC[self.model])

    def profile(self, pars):
        """One evaluation with CUDA events between phases: returns ({phase: ms}, vals, grad, info) (device tensors)."""
        torch = _lib.require_cuda()
        p = pars.detach().to(torch.float64).reshape(self.S, self.P).contiguous()
        vals = torch.empty((self.S, _lib.NVALS), dtype=torch.float64, device=self.device)
        grad = torch.empty((self.S, self.P), dtype=torch.float64, device=self.device)
        info = torch.empty((self.S,), dtype=torch.int32, device=self.device)
        ms = (ctypes.c_float * _lib.NPHASES)()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.nmgp_logpost_grad_profile(self._handle, p.data_ptr(), vals.data_ptr(), grad.data_ptr(),
                                                    info.data_ptr(), ms, ctypes.c_void_p(stream))
        _lib.check(rc, "nmgp_logpost_grad_profile")
        return dict(zip(_lib.PHASE_NAMES, [float(v) for v in ms])), vals, grad, info

    def _pinned_buffers(self):
        torch = _lib.require_cuda()
        if self._pinned is None:
            self._pinned = (
                torch.empty((self.S, self.P), dtype=torch.float64).pin_memory(),
                torch.empty((self.S, _lib.NVALS), dtype=torch.float64).pin_memory(),
                torch.empty((self.S, self.P), dtype=torch.float64).pin_memory(),
                torch.empty((self.S,), dtype=torch.int32).pin_memory(),
            )
        return self._pinned

    def value_and_grad_host(self, pars, need_grad: bool = True, pinned_io: bool = False):
        """Host path (what the reference's drivers see): pars [S,P] CPU float64 -> CPU (vals, grad, info).
        Host->device copy of pars, evaluation, device->host copy of the results and a stream synchronise all
        happen inside the C call `nmgp_logpost_grad_host`.  With pinned_io=True `pars` is staged through (and the
        results are returned in) the plan's pinned buffers, which are reused by the next call."""
        torch = _lib.require_cuda()
        p = torch.as_tensor(pars, dtype=torch.float64).detach().reshape(self.S, self.P)
        if p.is_cuda:
            raise TypeError("value_and_grad_host expects host memory")
        if pinned_io:
            pin_p, vals, grad, info = self._pinned_buffers()
            if p.data_ptr() != pin_p.data_ptr():
                pin_p.copy_(p)
            p = pin_p
        else:
            p = p.contiguous()
            vals = torch.empty((self.S, _lib.NVALS), dtype=torch.float64)
            grad = torch.empty((self.S, self.P), dtype=torch.float64)
            info = torch.empty((self.S,), dtype=torch.int32)
        if self.S == 0:
            return vals, (grad if need_grad else None), info
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.nmgp_logpost_grad_host(self._handle, p.data_ptr(), vals.data_ptr(),
                                                 grad.data_ptr() if need_grad else None, info.data_ptr(),
                                                 ctypes.c_void_p(stream))
        _lib.check(rc, "nmgp_logpost_grad_host")
        return vals, (grad if need_grad else None), info

    def pinned_pars(self):
        """The plan's pinned staging buffer for `pars` ([S,P]); fill it in place and call
        value_and_grad_host(buf, pinned_io=True) to avoid an extra host copy."""
        return self._pinned_buffers()[0]

    # ------------------------------------------------------------------ device-resident MAP loop
    def map_fit(self, pars0, steps: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, frozen=None,
                record_every: int = 1):
        """The drivers' MAP loop (Adam on `nlogpos_obj*`: Stationary_model.py:106-131, Separable_model.py:149-231,
        Nonseparable_model_mpisim.py:163-207) for all S subjects at once, entirely on the device: per iteration one
        batched value+gradient evaluation and one `nmgp_adam_step`; `pars` and `grad` never cross PCIe.
        pars0 [S,P] (host or device) -> (pars [S,P] CUDA, trace [steps/record_every, S, 6] CUDA of the value tuples
        *before* each recorded update, info [S])."""
        torch = _lib.require_cuda()
        p = torch.as_tensor(pars0, dtype=torch.float64).to(self.device).reshape(self.S, self.P).clone().contiguous()
        m = torch.zeros_like(p)
        v = torch.zeros_like(p)
        fz = None
        if frozen is not None:
            fz = torch.as_tensor(frozen, dtype=torch.bool).reshape(self.P).to(self.device).to(torch.uint8).contiguous()
        trace = []
        # fixed output buffers: with unchanged pointers a launch-bound plan replays each evaluation as one CUDA graph
        buf = (torch.empty((self.S, _lib.NVALS), dtype=torch.float64, device=self.device), torch.empty_like(p),
               torch.empty((self.S,), dtype=torch.int32, device=self.device))
        info = None
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            for it in range(1, steps + 1):
                vals, grad, info = self.value_and_grad(p, out=buf)
                if (it - 1) % record_every == 0:
                    trace.append(vals.clone())
                rc = self.lib.nmgp_adam_step(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), info.data_ptr(),
                                             fz.data_ptr() if fz is not None else None, self.S, self.P, float(lr),
                                             float(betas[0]), float(betas[1]), float(eps), it, ctypes.c_void_p(stream))
                _lib.check(rc, "nmgp_adam_step")
        return p, (torch.stack(trace) if trace else torch.empty((0, self.S, _lib.NVALS), device=self.device)), info

    # ------------------------------------------------------------------ device-resident HMC
    def hmc_sample(self, pars0, sample_size: int, step_size: float, num_steps_in_leap: int, momenta=None, log_uniforms=None,
                   seed: int | None = None, keep_every: int = 1):
        """Hamiltonian Monte Carlo with potential `nlogpos_obj*` for all S subjects at once (independent chains), entirely on
        the device -- what the drivers' `HMC_Sampler.HMC_sampler.sampler(sample_size, potential_func, init_position, step_size,
        num_steps_in_leap, duplicate_samples=True, ...).main_hmc_loop()` does per subject (Separable_model.py:209-210,
        Nonseparable_model_mpiKAISER.py:267-270): per sample a standard-normal momentum, `num_steps_in_leap` leapfrog steps of
        size `step_size` (half kicks at both ends), Metropolis accept with a rejected proposal duplicating the current state.
        One batched value+gradient evaluation per leapfrog step; positions, momenta and gradients never cross PCIe.
        momenta [sample_size,S,P] / log_uniforms [sample_size,S] may be supplied (tests); else drawn on the device from `seed`.
        Returns (samples [sample_size // keep_every, S, P] CUDA, accept_rate [S] CUDA, potential [S] CUDA)."""
        torch = _lib.require_cuda()
        f64 = torch.float64
        q = torch.as_tensor(pars0, dtype=f64).to(self.device).reshape(self.S, self.P).clone().contiguous()
        gen = None
        if momenta is None or log_uniforms is None:
            gen = torch.Generator(device=self.device)
            gen.manual_seed(0 if seed is None else int(seed))
        vals, grad, info = self.value_and_grad(q)
        if int(info.abs().sum()) != 0:
            raise _lib.NmgpError("hmc_sample: the initial position of some subject is not valid (info != 0)")
        U = vals[:, 0].clone().contiguous()
        grad = grad.clone().contiguous()
        accepted = torch.empty((self.S,), dtype=torch.int32, device=self.device)
        n_acc = torch.zeros((self.S,), dtype=f64, device=self.device)
        samples = []
        eps, L = float(step_size), int(num_steps_in_leap)
        if L < 1 or keep_every < 1:
            raise ValueError("hmc_sample needs num_steps_in_leap >= 1 and keep_every >= 1")
        lib, S, P = self.lib, self.S, self.P
        # fixed proposal / output buffers: with unchanged pointers a launch-bound plan replays every leapfrog evaluation as
        # one CUDA graph (nmgp_plan_set_graph)
        p, qp = torch.empty_like(q), torch.empty_like(q)
        buf = (torch.empty((S, _lib.NVALS), dtype=f64, device=self.device), torch.empty_like(q),
               torch.empty((S,), dtype=torch.int32, device=self.device))
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            for it in range(sample_size):
                if momenta is not None:
                    p0 = torch.as_tensor(momenta[it], dtype=f64).to(self.device).reshape(S, P).contiguous()
                else:
                    p0 = torch.randn((S, P), dtype=f64, device=self.device, generator=gen)
                if log_uniforms is not None:
                    lu = torch.as_tensor(log_uniforms[it], dtype=f64).to(self.device).reshape(S).contiguous()
                else:
                    lu = torch.log(torch.rand((S,), dtype=f64, device=self.device, generator=gen))
                p.copy_(p0)
                qp.copy_(q)
                failed = torch.zeros((S,), dtype=torch.int32, device=self.device)
                _lib.check(lib.nmgp_hmc_kick(p.data_ptr(), grad.data_ptr(), None, S, P, 0.5 * eps, stream), "nmgp_hmc_kick")
                vp = gp = None
                for l in range(L):
                    _lib.check(lib.nmgp_hmc_drift(qp.data_ptr(), p.data_ptr(), S, P, eps, stream), "nmgp_hmc_drift")
                    vp, gp, ip = self.value_and_grad(qp, out=buf)
                    failed |= (ip != 0).to(torch.int32)
                    _lib.check(lib.nmgp_hmc_kick(p.data_ptr(), gp.data_ptr(), ip.data_ptr(), S, P,
                                                 eps if l + 1 < L else 0.5 * eps, stream), "nmgp_hmc_kick")
                _lib.check(lib.nmgp_hmc_accept(q.data_ptr(), qp.data_ptr(), grad.data_ptr(), gp.data_ptr(), U.data_ptr(),
                                               vp.data_ptr(), p0.data_ptr(), p.data_ptr(), failed.data_ptr(), lu.data_ptr(),
                                               accepted.data_ptr(), S, P, stream), "nmgp_hmc_accept")
                n_acc += accepted.to(f64)
                if (it + 1) % keep_every == 0:
                    samples.append(q.clone())
        out = torch.stack(samples) if samples else torch.empty((0, S, P), dtype=f64, device=self.device)
        return out, n_acc / max(sample_size, 1), U

    # ------------------------------------------------------------------ posterior prediction (nonseparable model)
    def _xstar(self, xstar):
        torch = _lib.require_cuda()
        xs = torch.as_tensor(xstar, dtype=torch.float64).to(self.device)
        if xs.dim() == 1:
            xs = xs.unsqueeze(0).expand(self.S, -1)
        if xs.dim() != 2 or xs.shape[0] != self.S:
            raise ValueError(f"expected new inputs [G] or [S,G]; got {tuple(xs.shape)}")
        return xs.contiguous()

    def predict_prior_moments(self, pars, xstar):
        """Conditional moments of the two GP priors at new inputs (Utility/prediction.py:1060-1092; separable model: :53-69).
        pars [S,P], xstar [G] or [S,G] -> (mu_l [S,G], s2_l [S,G], mu_uL [S,G,T], s2_uL [S,G]), CUDA tensors; for a
        separable plan the second pair is the conditional of tilde_sigma ([S,G,1], [S,G])."""
        torch = _lib.require_cuda()
        if self.model == "stationary":
            raise ValueError("the stationary model has no GP priors to condition")
        p = torch.as_tensor(pars, dtype=torch.float64).detach().to(self.device).reshape(self.S, self.P).contiguous()
        xs = self._xstar(xstar)
        G, T = int(xs.shape[1]), (self.M * (self.M + 1) // 2 if self.model == "nonseparable" else 1)
        kw = dict(dtype=torch.float64, device=self.device)
        mu_l, s2_l = torch.empty((self.S, G), **kw), torch.empty((self.S, G), **kw)
        mu_u, s2_u = torch.empty((self.S, G, T), **kw), torch.empty((self.S, G), **kw)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.nmgp_predict_prior_moments(self._handle, p.data_ptr(), xs.data_ptr(), G, mu_l.data_ptr(),
                                                     s2_l.data_ptr(), mu_u.data_ptr(), s2_u.data_ptr(),
                                                     ctypes.c_void_p(stream))
        _lib.check(rc, "nmgp_predict_prior_moments")
        return mu_l, s2_l, mu_u, s2_u

    def predict_moments(self, pars, xstar, tl_star, uL_star, raw_factor: bool = False):
        """Predictive mean and variance of the M outputs for sampled (tilde_l*, uL*) (Utility/prediction.py:1130-1165).
        pars [S,P], xstar [G] or [S,G], tl_star [S,G,ns], uL_star [S,G,ns,T] ->
        (mu_f [S,G,ns,M], s2_y [S,G,ns,M], info [S]), CUDA tensors.  raw_factor=True: uL_star is the factor's triangle
        itself (no exp on the diagonal), as in point_predsample_inhomogeneous (prediction.py:1310-1311)."""
        torch = _lib.require_cuda()
        if self.model != "nonseparable":
            raise ValueError("prediction is implemented for the nonseparable model")
        p = torch.as_tensor(pars, dtype=torch.float64).detach().to(self.device).reshape(self.S, self.P).contiguous()
        xs = self._xstar(xstar)
        G, T = int(xs.shape[1]), self.M * (self.M + 1) // 2
        tl = torch.as_tensor(tl_star, dtype=torch.float64).to(self.device).reshape(self.S, G, -1).contiguous()
        ns = int(tl.shape[2])
        ul = torch.as_tensor(uL_star, dtype=torch.float64).to(self.device).reshape(self.S, G, ns, T).contiguous()
        kw = dict(dtype=torch.float64, device=self.device)
        mu_f, s2_y = torch.empty((self.S, G, ns, self.M), **kw), torch.empty((self.S, G, ns, self.M), **kw)
        info = torch.zeros((self.S,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.nmgp_predict_moments(self._handle, p.data_ptr(), xs.data_ptr(), G, ns, tl.data_ptr(),
                                               ul.data_ptr(), _lib.PRED_RAW_FACTOR if raw_factor else 0, mu_f.data_ptr(),
                                               s2_y.data_ptr(), info.data_ptr(),
                                               ctypes.c_void_p(stream))
        _lib.check(rc, "nmgp_predict_moments")
        return mu_f, s2_y, info

    def predict_moments_sep(self, pars, xstar, tl_star, ts_star):
        """Separable / stationary predictive moments (Utility/prediction.py:82-118, 226-266, 372-398, 1587-1596).
        pars [S,P], xstar [G] or [S,G], tl_star / ts_star [S,G,ns] (tilde_l*, tilde_sigma* per new input and sample) ->
        (mu_f [S,G,ns,M] = k_f^T Sigma^-1 y, quad [S,G,ns,M] = diag(k_f^T Sigma^-1 k_f), info [S]), CUDA tensors."""
        torch = _lib.require_cuda()
        if self.model == "nonseparable":
            raise ValueError("use predict_moments for the nonseparable model")
        p = torch.as_tensor(pars, dtype=torch.float64).detach().to(self.device).reshape(self.S, self.P).contiguous()
        xs = self._xstar(xstar)
        G = int(xs.shape[1])
        tl = torch.as_tensor(tl_star, dtype=torch.float64).to(self.device).reshape(self.S, G, -1).contiguous()
        ns = int(tl.shape[2])
        ts = torch.as_tensor(ts_star, dtype=torch.float64).to(self.device).reshape(self.S, G, ns).contiguous()
        kw = dict(dtype=torch.float64, device=self.device)
        mu_f, quad = torch.empty((self.S, G, ns, self.M), **kw), torch.empty((self.S, G, ns, self.M), **kw)
        info = torch.zeros((self.S,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.nmgp_predict_moments_sep(self._handle, p.data_ptr(), xs.data_ptr(), G, ns, tl.data_ptr(),
                                                   ts.data_ptr(), mu_f.data_ptr(), quad.data_ptr(), info.data_ptr(),
                                                   ctypes.c_void_p(stream))
        _lib.check(rc, "nmgp_predict_moments_sep")
        return mu_f, quad, info

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self.lib.nmgp_plan_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------- subjects of unequal shape
def group_by_shape(shapes):
    """[(N_0, M_0), (N_1, M_1), ...] -> {(N, M): [subject indices in input order]} (insertion-ordered by first occurrence)."""
    groups: dict = {}
    for i, sh in enumerate(shapes):
        groups.setdefault((int(sh[0]), int(sh[1])), []).append(i)
    return groups


class RaggedPlans:
    """Subjects with their OWN numbers of time points (and outputs): the personalized / real-data drivers fit one patient per
    process, every patient with a different N (Nonseparable_model_personalized.py, Nonseparable_model_mpiKAISER.py:39-60,
    Separable_model_personalized.py:26-27).  A `LogPosteriorPlan` batches subjects of one shape; this groups the subjects by
    (N, M), keeps one plan per shape and presents them in the caller's subject order.  Parameter vectors have different
    lengths, so they travel as a list of 1-D tensors (host or device) or, for device-resident loops, as the packed
    per-shape matrices of `pack`."""

    def __init__(self, model: str, xs, Ys, hyper: Mapping[str, float] | None = None, prior: bool = True, device=None,
                 workspace_limit_bytes: int = 0):
        torch = _lib.require_cuda()
        if model in HADAMARD_MODELS:
            raise ValueError("RaggedPlans covers the regularly sampled models")
        if len(xs) != len(Ys):
            raise ValueError(f"got {len(xs)} time-stamp vectors and {len(Ys)} observation matrices")
        self.model, self.S = model, len(xs)
        x_t = [torch.as_tensor(x, dtype=torch.float64).reshape(-1) for x in xs]
        Y_t = [torch.as_tensor(Y, dtype=torch.float64) for Y in Ys]
        for i, (x, Y) in enumerate(zip(x_t, Y_t)):
            if Y.dim() != 2 or Y.shape[0] != x.shape[0]:
                raise ValueError(f"subject {i}: expected x [N] and Y [N,M]; got {tuple(x.shape)} and {tuple(Y.shape)}")
        self.shapes = [(int(Y.shape[0]), int(Y.shape[1])) for Y in Y_t]
        self.groups = group_by_shape(self.shapes)
        self.plans = {}
        for sh, idx in self.groups.items():
            self.plans[sh] = LogPosteriorPlan(model, torch.stack([x_t[i] for i in idx]), torch.stack([Y_t[i] for i in idx]),
                                              hyper, prior=prior, device=device, workspace_limit_bytes=workspace_limit_bytes)
        self.device = next(iter(self.plans.values())).device if self.plans else torch.device(
            device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.P = [n_params(model, N, M) for N, M in self.shapes]

    def pack(self, pars_list):
        """list of S parameter vectors (subject order) -> {shape: [S_shape, P_shape] CUDA float64}."""
        torch = _lib.require_cuda()
        if len(pars_list) != self.S:
            raise ValueError(f"expected {self.S} parameter vectors, got {len(pars_list)}")
        out = {}
        for sh, idx in self.groups.items():
            rows = [torch.as_tensor(pars_list[i], dtype=torch.float64).detach().reshape(-1) for i in idx]
            for i, r in zip(idx, rows):
                if r.numel() != self.P[i]:
                    raise ValueError(f"subject {i}: expected {self.P[i]} parameters for shape {sh}, got {r.numel()}")
            out[sh] = torch.stack([r.to(self.device) for r in rows]).contiguous()
        return out

    def unpack(self, packed):
        """{shape: [S_shape, ...]} -> list of S rows in subject order."""
        rows = [None] * self.S
        for sh, idx in self.groups.items():
            for k, i in enumerate(idx):
                rows[i] = packed[sh][k]
        return rows

    def value_and_grad_packed(self, packed, need_grad: bool = True):
        """{shape: pars} -> {shape: (vals, grad, info)}: one batched evaluation per shape, all stream-ordered, no host sync."""
        return {sh: self.plans[sh].value_and_grad(packed[sh], need_grad=need_grad) for sh in self.groups}

    def value_and_grad(self, pars_list, need_grad: bool = True):
        """list of S parameter vectors -> (vals [S,6] CUDA, list of S gradient vectors (CUDA) or None, info [S] CUDA),
        all in subject order."""
        torch = _lib.require_cuda()
        res = self.value_and_grad_packed(self.pack(pars_list), need_grad=need_grad)
        vals = torch.empty((self.S, _lib.NVALS), dtype=torch.float64, device=self.device)
        info = torch.empty((self.S,), dtype=torch.int32, device=self.device)
        for sh, idx in self.groups.items():
            ix = torch.as_tensor(idx, device=self.device)
            vals[ix] = res[sh][0]
            info[ix] = res[sh][2]
        grads = self.unpack({sh: r[1] for sh, r in res.items()}) if need_grad else None
        return vals, grads, info

    def map_fit(self, pars_list, steps: int, lr: float, **kw):
        """`LogPosteriorPlan.map_fit` per shape -> (list of S fitted vectors, info [S]) in subject order."""
        torch = _lib.require_cuda()
        packed = self.pack(pars_list)
        fitted, info = {}, torch.empty((self.S,), dtype=torch.int32, device=self.device)
        for sh, idx in self.groups.items():
            fitted[sh], _, inf = self.plans[sh].map_fit(packed[sh], steps, lr, **kw)
            info[torch.as_tensor(idx, device=self.device)] = inf
        return self.unpack(fitted), info

    def close(self):
        for p in self.plans.values():
            p.close()
