"""Objective functions with the reference's names, signatures and return conventions (Utility/logpos.py),
evaluated by the B200 CUDA library through the C ABI.

    nlogpos_obj_S   (pars, Y, x, mu_tilde_l, sigma_tilde_l, a=1, b=1, c=10, verbose=False, Prior=True)     logpos.py:383
    nlogpos_obj     (pars, Y, x, mu_tilde_l=0., ..., a=1, b=1, c=10, verbose=False, Prior=True)            logpos.py:216
    nlogpos_obj_SVC (pars, Y, x, mu_tilde_l=0., alpha_tilde_l=5., ..., a=1, b=1, verbose=False, Prior=True) logpos.py:299

`pars` is the flat float64 parameter vector the drivers build with `torch.cat([...leaf tensors...])`; the returned
value is a 0-dim tensor on the same device that participates in autograd, so `NegLog.backward()` fills the
leaves' `.grad` exactly as in the drivers' MAP loops (e.g. Nonseparable_model_mpisim.py:183-190) -- the gradient
is the analytic one computed on the GPU in the same call.  With verbose=True the extra tuple entries (loglik and
the prior components) are returned detached, as the drivers only print them.

A plan (cached loop-invariant device state: x, Y, hyper-parameters, factored GP-prior covariances, workspace) is
kept per (model, data, hyper-parameters) so that repeated calls from a MAP / HMC loop only move `pars` in and
`value, components, grad` out.  Batched entry points for many subjects at once: `nlogpos_obj*_batched`.

There is no CPU fallback: without a CUDA device or without the built library these functions raise.
"""
from __future__ import annotations

import collections
import warnings

import numpy as np

from . import _lib, batched
from .batched import LogPosteriorPlan

_PLAN_CACHE: "collections.OrderedDict[tuple, LogPosteriorPlan]" = collections.OrderedDict()
_PLAN_CACHE_SIZE = 8


# ------------------------------------------------------------------------------------- parameter slicing
def vec2pars(pars, N, M):
    """(logpos.py:17-29)"""
    T = M * (M + 1) // 2
    return pars[:N], pars[N:2 * N], pars[2 * N:2 * N + T], pars[-1]


def vec2pars_SVC(pars, N, M):
    """(logpos.py:32-43)"""
    T = M * (M + 1) // 2
    return pars[:N], pars[N:N + N * T], pars[-1]


def vec2pars_S(pars, M):
    """(logpos.py:46-57)"""
    T = M * (M + 1) // 2
    return pars[0], pars[1], pars[2:2 + T], pars[-1]


def generate_K_index_SVC(L_f_list):
    """Stack the per-time-point factors and form L L^T, time-major (logpos.py:111-118), by `nmgp_gram`."""
    import ctypes
    torch = _lib.require_cuda()
    dev = L_f_list[0].device
    L = torch.cat([torch.as_tensor(l, dtype=torch.float64).detach().cuda() for l in L_f_list], dim=0).contiguous()
    out = torch.empty((L.shape[0], L.shape[0]), dtype=torch.float64, device=L.device)
    _lib.check(_lib.load_library().nmgp_gram(L.data_ptr(), L.shape[0], L.shape[1], out.data_ptr(),
                                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "nmgp_gram")
    return out.to(dev)


# ------------------------------------------------------------------------------------- plan cache
# A MAP / HMC loop calls the objective thousands of times with the SAME Y and x tensors.  Hashing their contents needs them
# on the host (a D2H copy + synchronisation for CUDA tensors), so the content hash is remembered per
# (storage address, version counter, shape, dtype, device): in-place edits bump the version and re-hash.  CPU tensors are
# additionally spot-checked (up to 32 strided elements) on every call, which catches a buffer rewritten behind torch's back
# (e.g. through a numpy view); CUDA tensors are trusted between version bumps -- pass `plan=` to
# `nlogpos_obj*_batched` for explicit control.
_FP_CACHE: "collections.OrderedDict[tuple, tuple]" = collections.OrderedDict()
_FP_CACHE_SIZE = 64


def _spot(t):
    flat = t.detach().reshape(-1)
    n = flat.numel()
    if n == 0:
        return ()
    step = max(1, n // 32)
    return tuple(flat[::step][:32].tolist())


def _fingerprint(t):
    key = (t.data_ptr(), t._version, tuple(t.shape), t.dtype, str(t.device))
    hit = _FP_CACHE.get(key)
    spot = _spot(t) if not t.is_cuda else None
    if hit is not None and hit[1] == spot:
        _FP_CACHE.move_to_end(key)
        return hit[0]
    a = t.detach().cpu().contiguous().numpy()
    fp = (a.shape, hash(a.tobytes()))
    _FP_CACHE[key] = (fp, spot)
    while len(_FP_CACHE) > _FP_CACHE_SIZE:
        _FP_CACHE.popitem(last=False)
    return fp


def _get_plan(model, Y, x, hyper, prior, indx=None):
    torch = _lib.require_cuda()
    Y = torch.as_tensor(Y, dtype=torch.float64)
    x = torch.as_tensor(x, dtype=torch.float64).reshape(-1)
    key = (model, _fingerprint(Y), _fingerprint(x), tuple(sorted((k, float(v)) for k, v in hyper.items())), bool(prior),
           torch.cuda.current_device(), None if indx is None else _fingerprint(torch.as_tensor(indx)))
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        plan = LogPosteriorPlan(model, x, Y, hyper, prior=prior, indx=indx)
        _PLAN_CACHE[key] = plan
        while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
            _, old = _PLAN_CACHE.popitem(last=False)
            old.close()
    else:
        _PLAN_CACHE.move_to_end(key)
    return plan


def clear_plan_cache():
    while _PLAN_CACHE:
        _, p = _PLAN_CACHE.popitem()
        p.close()
    _FP_CACHE.clear()


class NmgpNotPositiveDefinite(RuntimeWarning):
    """The covariance of a subject was not positive definite at the given parameters (Cholesky pivot `info` failed)."""


# ------------------------------------------------------------------------------------- autograd bridge
def _make_function():
    torch = _lib.require_cuda()

    class _NegLogPosterior(torch.autograd.Function):
        """vals[S,6] = plan(pars[S,P]); only vals[:,0] (= -log posterior) is differentiable."""

        @staticmethod
        def forward(ctx, pars, plan):
            need_grad = pars.requires_grad
            if pars.is_cuda:
                vals, grad, info = plan.value_and_grad(pars, need_grad=need_grad)
            else:
                vals, grad, info = plan.value_and_grad_host(pars, need_grad=need_grad)
            ctx.save_for_backward(grad if need_grad else None)
            ctx.shape = pars.shape
            ctx.mark_non_differentiable(info)
            return vals, info

        @staticmethod
        def backward(ctx, gvals, _ginfo):
            (grad,) = ctx.saved_tensors
            if grad is None:
                return None, None
            return (gvals[:, :1] * grad).reshape(ctx.shape), None

    return _NegLogPosterior


_FN = None


def _evaluate(plan, pars):
    """pars [S,P] or [P] tensor -> (vals [S,6] with autograd on column 0, info [S])."""
    global _FN
    torch = _lib.require_cuda()
    if _FN is None:
        _FN = _make_function()
    p = torch.as_tensor(pars)
    if p.dtype != torch.float64:
        p = p.to(torch.float64)
    return _FN.apply(p.reshape(plan.S, plan.P), plan)


def _single(model, pars, Y, x, hyper, verbose, Prior, indx=None):
    """Failure contract: if the covariance is not positive definite at `pars` the value and its
    gradient are NaN -- the in-band signal the reference's own callers test for (logpos.py:267) and wrap in try/except
    (Nonseparable_model_mpisim.py:330-334) -- and a NmgpNotPositiveDefinite warning names the failing pivot.  (The
    reference's LU-based inverse / logdet return NaN or garbage without any signal in that case.)  With CUDA `pars` the
    call stays asynchronous and only the NaN is reported."""
    plan = _get_plan(model, Y, x, hyper, Prior, indx)
    vals, info = _evaluate(plan, pars)
    if not info.is_cuda and int(info[0]) != 0:
        warnings.warn(f"{model} objective: covariance not positive definite (Cholesky pivot {int(info[0])}); "
                      "value and gradient are NaN", NmgpNotPositiveDefinite, stacklevel=3)
    if not verbose:
        return vals[0, 0]
    n = batched.N_VERBOSE[model]
    return (vals[0, 0],) + tuple(vals[0, k].detach() for k in range(1, n))


# ------------------------------------------------------------------------------------- reference signatures
def nlogpos_obj_S(pars, Y, x, mu_tilde_l, sigma_tilde_l, a=1, b=1, c=10, verbose=False, Prior=True):
    """Stationary model: -log posterior [, loglik, lp_tilde_l, lp_uL, lp_sigma2]  (logpos.py:383-402)."""
    hyper = dict(mu_tilde_l=mu_tilde_l, sigma_tilde_l=sigma_tilde_l, a=a, b=b, c=c)
    return _single("stationary", pars, Y, x, hyper, verbose, Prior)


def nlogpos_obj(pars, Y, x, mu_tilde_l=0., alpha_tilde_l=1., beta_tilde_l=1., mu_tilde_sigma=0., alpha_tilde_sigma=1.,
                beta_tilde_sigma=1., a=1, b=1, c=10, verbose=False, Prior=True):
    """Separable model: -log posterior [, loglik, lp_tilde_l, lp_tilde_sigma, lp_uL, lp_sigma2]  (logpos.py:216-234)."""
    hyper = dict(mu_tilde_l=mu_tilde_l, alpha_tilde_l=alpha_tilde_l, beta_tilde_l=beta_tilde_l,
                 mu_tilde_sigma=mu_tilde_sigma, alpha_tilde_sigma=alpha_tilde_sigma,
                 beta_tilde_sigma=beta_tilde_sigma, a=a, b=b, c=c)
    return _single("separable", pars, Y, x, hyper, verbose, Prior)


def nlogpos_obj_SVC(pars, Y, x, mu_tilde_l=0., alpha_tilde_l=5., beta_tilde_l=1., mu_L=0., alpha_L=5., beta_L=1., a=1,
                    b=1, verbose=False, Prior=True):
    """Nonseparable model: -log posterior [, loglik, lp_tilde_l, lp_uL, lp_sigma2]  (logpos.py:299-323)."""
    hyper = dict(mu_tilde_l=mu_tilde_l, alpha_tilde_l=alpha_tilde_l, beta_tilde_l=beta_tilde_l, mu_L=mu_L,
                 alpha_L=alpha_L, beta_L=beta_L, a=a, b=b)
    return _single("nonseparable", pars, Y, x, hyper, verbose, Prior)


def _positive(out, verbose):
    if not verbose:
        return -out
    return (-out[0],) + tuple(out[1:])


def logpos_S(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, mu_tilde_l, sigma_tilde_l, a, b, c, verbose=False,
             Prior=True):
    """log posterior of the stationary model from its separate parameters (logpos.py:405-462)."""
    torch = _lib.require_cuda()
    pars = torch.cat([tilde_l.reshape(1), tilde_sigma.reshape(1), uL_vec.reshape(-1), tilde_sigma2_err.reshape(1)])
    return _positive(nlogpos_obj_S(pars, Y, x, mu_tilde_l, sigma_tilde_l, a, b, c, verbose, Prior), verbose)


def logpos(tilde_l, tilde_sigma, uL_vec, tilde_sigma2_err, Y, x, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
           mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, a, b, c, verbose=False, Prior=True):
    """log posterior of the separable model (logpos.py:237-296)."""
    torch = _lib.require_cuda()
    pars = torch.cat([tilde_l.reshape(-1), tilde_sigma.reshape(-1), uL_vec.reshape(-1), tilde_sigma2_err.reshape(1)])
    return _positive(nlogpos_obj(pars, Y, x, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma,
                                 alpha_tilde_sigma, beta_tilde_sigma, a, b, c, verbose, Prior), verbose)


def logpos_SVC(tilde_l, uL_vecs, tilde_sigma2_err, Y, x, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L,
               a, b, verbose=False, Prior=True):
    """log posterior of the nonseparable model (logpos.py:326-380)."""
    torch = _lib.require_cuda()
    pars = torch.cat([tilde_l.reshape(-1), uL_vecs.reshape(-1), tilde_sigma2_err.reshape(1)])
    return _positive(nlogpos_obj_SVC(pars, Y, x, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L, a, b,
                                     verbose, Prior), verbose)


# ------------------------------------------------------------------------------------- deviance (logpos.py:189-213)
def deviance(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, Y, x):
    """-2 x the (un-normalised) separable log-likelihood, `L_vec` being the CONSTRAINED row-major triangle (no exp on
    its diagonal, unlike `logpos`: logpos.py:189-213 forms L L^T from it directly).  The reference goes through
    kron_inv + kron_logdet (two symeig); here it is the separable plan evaluated with Prior=False -- B = L L^T does
    not depend on the signs of L's diagonal, so log|diag| is the equivalent unconstrained vector."""
    torch = _lib.require_cuda()
    Y = torch.as_tensor(Y, dtype=torch.float64)
    M = Y.shape[1]
    Lv = torch.as_tensor(L_vec, dtype=torch.float64).reshape(-1)
    diag = torch.zeros(Lv.shape[0], dtype=torch.bool, device=Lv.device)
    diag[torch.cumsum(torch.arange(1, M + 1), 0) - 1] = True
    uL = torch.where(diag, torch.log(torch.abs(torch.where(diag, Lv, torch.ones_like(Lv)))), Lv)
    pars = torch.cat([torch.as_tensor(tilde_l, dtype=torch.float64).reshape(-1),
                      torch.as_tensor(tilde_sigma, dtype=torch.float64).reshape(-1), uL,
                      torch.as_tensor(tilde_sigma2_err, dtype=torch.float64).reshape(1)])
    return 2.0 * nlogpos_obj(pars, Y, x, verbose=False, Prior=False)


def deviance_obj(pars, Y, x):
    """(logpos.py:189-199)"""
    N, M = Y.shape
    tilde_l, tilde_sigma, L_vec, tilde_sigma2_err = vec2pars(pars, N, M)
    return deviance(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, Y, x)


# ------------------------------------------------------------------------------------- irregular sampling ("Hadamard")
def vec2pars_hadamard_SVC(pars, N, M):
    """(logpos.py:60-72)"""
    T = M * (M + 1) // 2
    return pars[:N], pars[N:N + N * T], pars[-1]


def generate_vectorized_indexes(indx1, indx2):
    """All (i, j) index pairs of two index vectors, row-major (logpos.py:75-86)."""
    torch = _lib.require_cuda()
    n1, n2 = indx1.shape[0], indx2.shape[0]
    return (indx1.reshape(-1, 1).repeat(1, n2).reshape(-1).to(torch.int64), indx2.repeat(n1).to(torch.int64))


def generate_K_index(B_f, indx):
    """K_i[n,n'] = B_f[indx_n, indx_n'] (logpos.py:89-99)."""
    torch = _lib.require_cuda()
    ix = torch.as_tensor(indx).to(torch.int64)
    return B_f[ix][:, ix]


def generate_K_index_SVC_hadamard0(L_f_list, indexes):
    """K_i[n,n'] = <L_n[indx_n,:], L_n'[indx_n',:]> (logpos.py:114-116)."""
    torch = _lib.require_cuda()
    L = torch.stack([L_f[int(i), :] for L_f, i in zip(L_f_list, indexes)])
    return L @ L.t()


generate_K_index_SVC_hadamard = generate_K_index_SVC_hadamard0   # same matrix, entry by entry (logpos.py:119-131)


def nlogpos_obj_hadamard(pars, x, indx, y, mu_tilde_l=0., alpha_tilde_l=1., beta_tilde_l=1., mu_tilde_sigma=0.,
                         alpha_tilde_sigma=1., beta_tilde_sigma=1., a=1, b=1, c=10, verbose=False, Prior=True):
    """Separable model, irregular sampling: -log posterior [, loglik, lp_tilde_l, lp_tilde_sigma, lp_L, lp_sigma2]
    (logpos.py:465-500)."""
    hyper = dict(mu_tilde_l=mu_tilde_l, alpha_tilde_l=alpha_tilde_l, beta_tilde_l=beta_tilde_l,
                 mu_tilde_sigma=mu_tilde_sigma, alpha_tilde_sigma=alpha_tilde_sigma, beta_tilde_sigma=beta_tilde_sigma,
                 a=a, b=b, c=c)
    return _single("hadamard", pars, y, x, hyper, verbose, Prior, indx)


def nlogpos_obj_hadamard_SVC(pars, x, indx, y, mu_tilde_l=0., alpha_tilde_l=1., beta_tilde_l=1., mu_L=0., alpha_L=1.,
                             beta_L=1., a=1, b=1, verbose=False, Prior=True):
    """Nonseparable model, irregular sampling: -log posterior [, loglik, lp_tilde_l, lp_L, lp_sigma2] (logpos.py:561-579)."""
    hyper = dict(mu_tilde_l=mu_tilde_l, alpha_tilde_l=alpha_tilde_l, beta_tilde_l=beta_tilde_l, mu_L=mu_L, alpha_L=alpha_L,
                 beta_L=beta_L, a=a, b=b)
    return _single("hadamard_svc", pars, y, x, hyper, verbose, Prior, indx)


def nlogpos_obj_hadamard_S(pars, x, indx, y, mu_tilde_l, sigma_tilde_l, a=1, b=1, c=10, verbose=False, Prior=True):
    """Stationary model, irregular sampling: -log posterior [, loglik, lp_tilde_l, lp_L, lp_sigma2] (logpos.py:640-652)."""
    hyper = dict(mu_tilde_l=mu_tilde_l, sigma_tilde_l=sigma_tilde_l, a=a, b=b, c=c)
    return _single("hadamard_s", pars, y, x, hyper, verbose, Prior, indx)


def logpos_hadamard(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, mu_tilde_l, alpha_tilde_l, beta_tilde_l,
                    mu_tilde_sigma, alpha_tilde_sigma, beta_tilde_sigma, a, b, c, verbose=False, Prior=True):
    """(logpos.py:503-558)"""
    torch = _lib.require_cuda()
    pars = torch.cat([tilde_l.reshape(-1), tilde_sigma.reshape(-1), L_vec.reshape(-1), tilde_sigma2_err.reshape(1)])
    return _positive(nlogpos_obj_hadamard(pars, x, indx, y, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_tilde_sigma,
                                          alpha_tilde_sigma, beta_tilde_sigma, a, b, c, verbose, Prior), verbose)


def logpos_hadamard_SVC(tilde_l, L_vecs, tilde_sigma2_err, x, indx, y, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L,
                        beta_L, a, b, verbose=False, Prior=True):
    """(logpos.py:582-637)"""
    torch = _lib.require_cuda()
    pars = torch.cat([tilde_l.reshape(-1), L_vecs.reshape(-1), tilde_sigma2_err.reshape(1)])
    return _positive(nlogpos_obj_hadamard_SVC(pars, x, indx, y, mu_tilde_l, alpha_tilde_l, beta_tilde_l, mu_L, alpha_L, beta_L,
                                              a, b, verbose, Prior), verbose)


def logpos_hadamard_S(tilde_l, tilde_sigma, L_vec, tilde_sigma2_err, x, indx, y, mu_tilde_l, sigma_tilde_l, a, b, c,
                      verbose=False, Prior=True):
    """(logpos.py:655-716)"""
    torch = _lib.require_cuda()
    pars = torch.cat([tilde_l.reshape(1), tilde_sigma.reshape(1), L_vec.reshape(-1), tilde_sigma2_err.reshape(1)])
    return _positive(nlogpos_obj_hadamard_S(pars, x, indx, y, mu_tilde_l, sigma_tilde_l, a, b, c, verbose, Prior), verbose)


# ------------------------------------------------------------------------------------- batched entry points
def nlogpos_obj_batched(model, pars, plan):
    """Batched objective over all S subjects of `plan` (model in {'stationary','separable','nonseparable'}):
    pars [S,P] -> (neg_logpost [S] (differentiable), components [S,5] detached, info [S]).
    This is what the subject-sharded drivers (sharding.py, bench.py) call."""
    if plan.model != model:
        raise ValueError(f"plan was built for the {plan.model} model, not {model}")
    vals, info = _evaluate(plan, pars)
    return vals[:, 0], vals[:, 1:].detach(), info
