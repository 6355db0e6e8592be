"""bench.py's synthetic observations do not depend on the sharding: compare make_inputs over [0, S) with two shards of it."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
a = types.SimpleNamespace(model="nonseparable", N=100, M=6, subjects=1000)
dev = torch.device("cuda:0")
x, Y, p = bench.make_inputs(a, 0, 1000, dev)
for lo, hi in ((0, 125), (375, 500), (333, 667), (875, 1000)):
    xs, Ys, ps = bench.make_inputs(a, lo, hi, dev)
    assert torch.equal(xs, x[lo:hi]) and torch.equal(ps, p[lo:hi]) and torch.equal(Ys, Y[lo:hi]), (lo, hi)
print("inputs identical for every shard; |Y| mean", float(Y.abs().mean()))
