"""CPU timing worker for the reference arm / cpu_baseline of bench.py.  TEST/BENCH INFRASTRUCTURE ONLY.

Evaluates the oracle (oracle/nmgp_oracle.py: the reference's algorithm -- dense inverse + logdet, autograd
backward) for `count` synthetic subjects, single-threaded, and prints one JSON line with the seconds it took.
bench.py starts one of these per host core, which is how the reference is deployed
(`srun -n 1000`, one single-threaded process per subject: Nonseparable_Model/sim_job:9).
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
    a = ap.parse_args()
    import numpy as np
    import torch
    torch.set_num_threads(a.threads)
    from oracle import nmgp_oracle as O
    from nonstationary_multivariate_gaussian_process_b200 import synth
    from bench import HYPER  # the benchmark's hyper-parameters
    rng = np.random.RandomState(a.first)
    subjects = []
    for s in range(a.first, a.first + a.count):
        x, _, _, _ = synth.truth(a.N, a.M, s)
        Y = rng.standard_normal((a.N, a.M))          # timing does not depend on the data values
        subjects.append((x, Y, synth.start_point(a.model, a.N, a.M, s)))
    for i in range(a.warmup):
        x, Y, p = subjects[i % len(subjects)]
        O.value_and_grad(a.model, p, Y, x, **HYPER[a.model])
    t0 = time.perf_counter()
    for x, Y, p in subjects:
        O.value_and_grad(a.model, p, Y, x, **HYPER[a.model])
    dt = time.perf_counter() - t0
    print(json.dumps({"seconds": dt, "count": a.count, "threads": a.threads}))


if __name__ == "__main__":
    main()
