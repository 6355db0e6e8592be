"""Per-evaluation time of one plan with fixed output buffers, with / without CUDA-graph replay (GPU box).
usage: python tools/run_latency2.py MODEL N M S [reps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nonstationary_multivariate_gaussian_process_b200 import synth
from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan
from bench import HYPER
model, N, M, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
rng = np.random.RandomState(0)
xs = np.stack([synth.truth(N, M, s)[0] for s in range(S)])
ps = np.stack([synth.start_point(model, N, M, s, 0.02) for s in range(S)])
Y = rng.standard_normal((S, N, M))
out = {}
for graph in (True, False):
    plan = LogPosteriorPlan(model, xs, Y, HYPER[model])
    plan.set_graph(graph)
    p = torch.from_numpy(ps).cuda()
    buf = (torch.empty((S, 6), dtype=torch.float64, device="cuda"), torch.empty((S, plan.P), dtype=torch.float64, device="cuda"),
           torch.empty((S,), dtype=torch.int32, device="cuda"))
    for _ in range(3):
        plan.value_and_grad(p, out=buf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.value_and_grad(p, out=buf)
    e1.record(); torch.cuda.synchronize()
    out["graph" if graph else "direct"] = {"ms": e0.elapsed_time(e1) / reps, "replays": plan.graph_replays, "launches": plan.last_launches}
    plan.close()
print(json.dumps({"model": model, "N": N, "M": M, "S": S, **out}))
