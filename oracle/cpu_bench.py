"""CPU worker for the reference arm / cpu_baseline / in-run parity of bench.py.  TEST/BENCH INFRASTRUCTURE ONLY.

Evaluates the oracle (oracle/nmgp_oracle.py: the reference's algorithm -- dense inverse + logdet, double eigh, autograd
backward) and prints one JSON line with the seconds it took.  bench.py starts one of these per host core, single-threaded,
which is how the reference is deployed (`srun -n 1000`, one process per subject: Nonseparable_Model/sim_job:9), or one
process with all cores for a single large subject.

  --data IN.npz [--out OUT.npz]   evaluate the subjects of IN.npz (x [S,N], Y [S,N,M], pars [S,P]: the bench's OWN inputs) and
                                  write vals [S,k] / grad [S,P] for the in-run parity figure
  (default)                       `count` synthetic subjects drawn from the model (synth.sample_subject, seeds first..)
  --map-iters K                   instead of single evaluations: the drivers' K-iteration Adam MAP loop on the first subject
                                  (Stationary_model.py:106-131), timed as K evaluations
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="nonseparable")
    ap.add_argument("--N", type=int, default=100)
    ap.add_argument("--M", type=int, default=6)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--count", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=1)
    ap.add_argument("--data", default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--map-iters", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--lr", type=float, default=0.1)
    a = ap.parse_args()
    import numpy as np
    import torch
    torch.set_num_threads(a.threads)
    from oracle import nmgp_oracle as O
    from nonstationary_multivariate_gaussian_process_b200 import synth
    from bench import HYPER  # the benchmark's hyper-parameters
    hyper = HYPER[a.model]
    if a.data:
        z = np.load(a.data)
        subjects = [(z["x"][s], z["Y"][s], z["pars"][s]) for s in range(z["x"].shape[0])]
    else:
        subjects = []
        for s in range(a.first, a.first + a.count):
            x, Y, _ = synth.sample_subject(a.N, a.M, s)
            subjects.append((x, Y, synth.start_point(a.model, a.N, a.M, s)))
    if a.map_iters > 0:
        x, Y, p0 = subjects[0]
        O.value_and_grad(a.model, p0, Y, x, **hyper)
        p = torch.from_numpy(np.array(p0)).clone().requires_grad_(True)
        opt = torch.optim.Adam([p], lr=a.lr)
        xt, Yt = torch.from_numpy(x), torch.from_numpy(Y)
        fn = {"stationary": O.logpost_stationary, "separable": O.logpost_separable, "nonseparable": O.logpost_nonseparable}[a.model]
        trace = []
        t0 = time.perf_counter()
        for _ in range(a.map_iters):
            opt.zero_grad()
            neg = -fn(p, Yt, xt, **hyper)[0]
            neg.backward()
            opt.step()
            trace.append(float(neg))
        dt = time.perf_counter() - t0
        if a.out:
            np.savez(a.out, trace=np.array(trace), pars=p.detach().numpy())
        print(json.dumps({"seconds": dt, "count": a.map_iters, "threads": a.threads}))
        return
    for i in range(a.warmup):
        x, Y, p = subjects[i % len(subjects)]
        O.value_and_grad(a.model, p, Y, x, **hyper)
    vals, grads = [], []
    t0 = time.perf_counter()
    for x, Y, p in subjects:
        for _ in range(a.repeat):
            v, g = O.value_and_grad(a.model, p, Y, x, **hyper)
        if a.out:
            vals.append(np.array([float(t) for t in v]))
            grads.append(g.numpy())
    dt = time.perf_counter() - t0
    if a.out:
        np.savez(a.out, vals=np.stack(vals), grad=np.stack(grads))
    print(json.dumps({"seconds": dt, "count": len(subjects) * a.repeat, "threads": a.threads}))


if __name__ == "__main__":
    main()
