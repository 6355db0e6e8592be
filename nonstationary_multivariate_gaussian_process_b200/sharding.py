"""Subject sharding across the GPUs of one box: one process per GPU, block partition, one tiny all-reduce.

The reference runs one OS process per subject (`srun -n 1000 python Nonseparable_model_mpisim.py`,
Nonseparable_Model/sim_job:9) and uses MPI only as a rank -> subject index (Nonseparable_model_mpisim.py:39-43);
nothing is exchanged between subjects.  Here rank r owns subjects [lo, hi) of the S_total; `x`, `Y`, `pars` and
`grad` of a subject never leave its GPU.  The only collective is an all-reduce(sum) of a short float64 vector of
sweep totals (sum of -log posterior, of each component, failure count, subject count) -- NCCL over NVLink on the
GPU box, gloo in the CPU tests.

BASELINE.json's north star additionally ties the hyper-priors across the subjects of a sharded run: the gradient of the
summed -log posterior with respect to the shared hyper-parameters is the sum over subjects of the per-subject gradients
(`LogPosteriorPlan.hyper_grad`, C entry point nmgp_hyper_grad), so it rides in the same all-reduce: `local_sweep_vector`
packs [summary (8) | hyper-gradient (9)] into one 17-double vector, `all_reduce_sweep` sums it across ranks.
"""
from __future__ import annotations

SUMMARY_FIELDS = ("neg_logpost", "c1", "c2", "c3", "c4", "c5", "n_failed", "n_subjects")


def shard_range(S_total: int, rank: int, world: int):
    """Block partition of S_total subjects: the first S_total % world ranks hold one extra subject."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, extra = divmod(S_total, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def local_summary(vals, info):
    """vals [S_local,6], info [S_local] -> float64 vector [8] of this rank's partial sums.
    Subjects whose factorisation failed (info != 0, NaN values) are counted, not summed."""
    import torch
    ok = (info == 0)
    v = torch.where(ok.unsqueeze(1), vals, torch.zeros_like(vals))
    out = torch.zeros(len(SUMMARY_FIELDS), dtype=torch.float64, device=vals.device)
    out[:6] = v.sum(0)
    out[6] = (~ok).sum().to(torch.float64)
    out[7] = float(vals.shape[0])
    return out


def all_reduce_summary(summary, group=None):
    """Sum the per-rank summaries over all ranks (in place) and return a dict.  No-op without a process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(summary, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(SUMMARY_FIELDS, summary.tolist()))


def local_hyper_grad(hgrad, info):
    """hgrad [S_local,9] (per-subject d(-logpost)/d hyper), info [S_local] -> float64 [9]: the sum over this rank's
    subjects whose evaluation succeeded (the same subjects local_summary sums)."""
    import torch
    ok = (info == 0).unsqueeze(1)
    return torch.where(ok, hgrad, torch.zeros_like(hgrad)).sum(0)


def local_sweep_vector(vals, info, hgrad=None):
    """One rank's contribution to the sweep's single all-reduce: [summary (8) | shared-hyper-parameter gradient (9)].
    CUDA tensors: one launch of the library's own reduction (nmgp_sweep_reduce, fixed summation order); host tensors (the
    gloo tests): the same sums with torch."""
    import torch
    if vals.is_cuda:
        import ctypes

        from . import _lib
        v = vals.detach().to(torch.float64).contiguous()
        i = info.to(torch.int32).contiguous()
        h = hgrad.detach().to(torch.float64).contiguous() if hgrad is not None else None
        out = torch.empty(len(SUMMARY_FIELDS) + 9, dtype=torch.float64, device=vals.device)
        with torch.cuda.device(vals.device):
            stream = torch.cuda.current_stream(vals.device).cuda_stream
            rc = _lib.load_library().nmgp_sweep_reduce(v.data_ptr(), h.data_ptr() if h is not None else None, i.data_ptr(),
                                                       int(v.shape[0]), out.data_ptr(), ctypes.c_void_p(stream))
        _lib.check(rc, "nmgp_sweep_reduce")
        return out
    s = local_summary(vals, info)
    h = local_hyper_grad(hgrad, info) if hgrad is not None else torch.zeros(9, dtype=torch.float64, device=vals.device)
    return torch.cat([s, h])


def all_reduce_sweep(vec, hyper_names=(), group=None):
    """Sum the [17] sweep vectors over all ranks (in place; one collective) -> (summary dict, hyper-gradient dict)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    v = vec.tolist()
    n = len(SUMMARY_FIELDS)
    return dict(zip(SUMMARY_FIELDS, v[:n])), dict(zip(hyper_names, v[n:n + len(hyper_names)]))


# hyper-parameters that must stay positive are stepped in log space
POSITIVE_HYPER = ("alpha_tilde_l", "beta_tilde_l", "alpha_tilde_sigma", "beta_tilde_sigma", "alpha_L", "beta_L", "sigma_tilde_l",
                  "a", "b", "c")


class HyperAdam:
    """Adam on the tied hyper-parameters (host scalars).  Every rank feeds it the same all-reduced gradient, so every rank
    takes the same step: the shared values never need a broadcast.  Positive hyper-parameters move in log space."""

    def __init__(self, hyper, tied, lr, betas=(0.9, 0.999), eps=1e-8):
        import math
        self.tied = tuple(tied)
        unknown = [k for k in self.tied if k not in hyper]
        if unknown:
            raise KeyError(f"tied hyper-parameters {unknown} are not hyper-parameters of this plan")
        self.lr, self.b1, self.b2, self.eps, self.t = float(lr), float(betas[0]), float(betas[1]), float(eps), 0
        self.m = {k: 0.0 for k in self.tied}
        self.v = {k: 0.0 for k in self.tied}
        self._log, self._exp = math.log, math.exp

    def step(self, hyper, grad):
        """hyper: current values; grad: d(sum of -log posterior)/d hyper -> new dict (untied keys unchanged)."""
        self.t += 1
        new = dict(hyper)
        for k in self.tied:
            pos = k in POSITIVE_HYPER
            h = float(hyper[k])
            g = float(grad[k]) * (h if pos else 1.0)              # chain rule for theta = log h
            if g != g:                                            # a sweep without a valid subject: leave the value alone
                continue
            self.m[k] = self.b1 * self.m[k] + (1.0 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1.0 - self.b2) * g * g
            mhat = self.m[k] / (1.0 - self.b1 ** self.t)
            vhat = self.v[k] / (1.0 - self.b2 ** self.t)
            theta = (self._log(h) if pos else h) - self.lr * mhat / (vhat ** 0.5 + self.eps)
            new[k] = self._exp(theta) if pos else theta
        return new


def tied_map_fit(plan, pars0, steps, lr, tied, hyper_lr, group=None, betas=(0.9, 0.999), eps=1e-8):
    """MAP fit of a sharded population whose subjects SHARE the hyper-parameters in `tied` (BASELINE.json's north star; the
    reference fixes one dictionary per subject, Nonseparable_model_mpisim.py:311-312): per iteration one fused sweep of this
    rank's subjects (value, gradient, hyper-gradient), the device-resident Adam step on their parameters
    (Nonseparable_model_mpisim.py:177-190), ONE all-reduce of 17 doubles, and the same Adam step on the shared
    hyper-parameters on every rank (`nmgp_plan_set_hyper` re-factors only the prior covariances that moved).
    Returns (pars [S_local,P] CUDA, hyper dict, trace: list of the all-reduced sweep summaries)."""
    import ctypes

    import torch

    from . import _lib
    p = torch.as_tensor(pars0, dtype=torch.float64).to(plan.device).reshape(plan.S, plan.P).clone().contiguous()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    from .batched import hyper_vector
    hyper = dict(zip(plan.hyper_names(), (float(x) for x in hyper_vector(plan.model, plan.hyper))))
    opt = HyperAdam(hyper, tied, hyper_lr, betas, eps)
    names, trace = plan.hyper_names(), []
    with torch.cuda.device(plan.device):
        stream = ctypes.c_void_p(torch.cuda.current_stream(plan.device).cuda_stream)
        for it in range(1, int(steps) + 1):
            vals, grad, hgrad, info = plan.value_grad_and_hyper_grad(p)
            if plan.S > 0:
                _lib.check(plan.lib.nmgp_adam_step(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), info.data_ptr(), None,
                                                   plan.S, plan.P, float(lr), float(betas[0]), float(betas[1]), float(eps), it,
                                                   stream), "nmgp_adam_step")
            totals, shared = all_reduce_sweep(local_sweep_vector(vals, info, hgrad), names, group=group)
            trace.append(totals)
            hyper = opt.step(hyper, shared)
            plan.set_hyper(hyper)
    return p, hyper, trace
