"""Tied-hyper-prior MAP fit of a subject-sharded population under torchrun (one rank per GPU, NCCL).
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/run_tied.py S_TOTAL [steps]
Checks that every rank ends with the same shared hyper-parameters (they are never broadcast: every rank takes the same Adam
step on the same all-reduced gradient) and prints the per-iteration time and the population objective."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from nonstationary_multivariate_gaussian_process_b200 import sharding, synth
from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan

S_total = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
N, M = 100, 6
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lo, hi = sharding.shard_range(S_total, rank, world)
base = [synth.sample_subject(N, M, s)[:2] + (synth.start_point("nonseparable", N, M, s, 0.05),) for s in range(16)]
xs, Ys, ps = (np.stack([base[s % 16][k] for s in range(lo, hi)]) for k in range(3))
hyper = {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0, "beta_L": 1.0, "a": 1e-2, "b": 1e-2}
plan = LogPosteriorPlan("nonseparable", xs, Ys, hyper)
tied = ("mu_tilde_l", "mu_L", "alpha_tilde_l", "alpha_L")
sharding.tied_map_fit(plan, ps, 2, 0.01, tied, 0.02)           # warm-up (then reset)
plan.set_hyper(hyper)
torch.cuda.synchronize(); t0 = time.perf_counter()
pars, hy, trace = sharding.tied_map_fit(plan, ps, steps, 0.01, tied, 0.02)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
hv = torch.tensor([hy[k] for k in tied], dtype=torch.float64, device="cuda")
same = True
if world > 1:
    all_h = [torch.empty_like(hv) for _ in range(world)]
    dist.all_gather(all_h, hv)
    same = all(torch.equal(a, all_h[0]) for a in all_h)
if rank == 0:
    print(json.dumps({"world": world, "S_total": S_total, "steps": steps, "ms_per_iteration": 1e3 * dt / steps,
                      "objective_first": trace[0]["neg_logpost"], "objective_last": trace[-1]["neg_logpost"],
                      "n_subjects": trace[-1]["n_subjects"], "n_failed": trace[-1]["n_failed"],
                      "hyper": {k: hy[k] for k in tied}, "identical_on_all_ranks": bool(same)}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
assert same
