"""GPU (needs two devices; skipped otherwise): plans on cuda:0 and cuda:1 inside ONE process.  Function attributes (dynamic
shared memory above 48 KB) and occupancy / SM-count caches are per device (common.cuh: NMGP_SMEM_ATTR_PER_DEVICE), the TMA
descriptors belong to the plan -- round 1 kept all three per process, so the second device's 50-70 KB launches failed."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HYPER = {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0,
         "a": 1e-2, "b": 1e-2}


def test_plans_on_two_devices_in_one_process(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from nonstationary_multivariate_gaussian_process_b200 import synth
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
    N, M, S = 40, 3, 6
    subs = [synth.sample_subject(N, M, s)[:2] + (synth.start_point("nonseparable", N, M, s, 0.02),) for s in range(S)]
    xs, Ys, ps = (np.stack(a) for a in zip(*subs))
    out = {}
    for engine in ("auto", "left", "left_stable", "recursive"):     # every kernel family with > 48 KB of shared memory
        for dev in (0, 1):
            with torch.cuda.device(dev):
                plan = LogPosteriorPlan("nonseparable", xs, Ys, HYPER, device=f"cuda:{dev}")
                plan.set_engine(engine)
                vals, grad, info = plan.value_and_grad(torch.from_numpy(ps).to(f"cuda:{dev}"))
                torch.cuda.synchronize(dev)
                assert int(info.abs().sum()) == 0
                out[(engine, dev)] = (vals.cpu(), grad.cpu())
                plan.close()
        assert torch.equal(out[(engine, 0)][0], out[(engine, 1)][0]) and torch.equal(out[(engine, 0)][1], out[(engine, 1)][1])
    # separable + prediction kernels (pred_prior_kernel requests its shared memory by N) on the second device
    with torch.cuda.device(1):
        plan = LogPosteriorPlan("separable", xs, Ys, device="cuda:1")
        p = np.stack([synth.start_point("separable", N, M, s, 0.02) for s in range(S)])
        vals, grad, info = plan.value_and_grad(torch.from_numpy(p).to("cuda:1"))
        torch.cuda.synchronize(1)
        assert int(info.abs().sum()) == 0 and bool(torch.isfinite(vals).all())
        plan.close()
