"""CPU-only, world_size 2 over gloo: the N>1 host path of bench.py / sharding.py -- block partition of the subjects,
per-rank partial sums, one all-reduce of the 8-double summary -- reproduces the single-process totals."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_sweep(S):
    rng = np.random.RandomState(0)
    vals = torch.from_numpy(rng.standard_normal((S, 6)))
    info = torch.from_numpy((rng.rand(S) < 0.1).astype(np.int32) * 5)
    vals[info != 0] = float("nan")
    return vals, info


def _worker(rank, world, port, S, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from nonstationary_multivariate_gaussian_process_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vals, info = _fake_sweep(S)
    lo, hi = sharding.shard_range(S, rank, world)
    summ = sharding.all_reduce_summary(sharding.local_summary(vals[lo:hi], info[lo:hi]))
    hg = torch.from_numpy(np.random.RandomState(1).standard_normal((S, 9)))
    names = ("mu_tilde_l", "alpha_tilde_l", "beta_tilde_l", "mu_L", "alpha_L", "beta_L", "a", "b")
    summ2, hyp = sharding.all_reduce_sweep(sharding.local_sweep_vector(vals[lo:hi], info[lo:hi], hg[lo:hi]), names)
    torch.save({"summary": summ, "summary2": summ2, "hyper": hyp, "range": (lo, hi)}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_all_reduce_equals_single_process_sum(tmp_path):
    from nonstationary_multivariate_gaussian_process_b200 import sharding
    S, world = 101, 2
    mp.spawn(_worker, args=(world, _free_port(), S, str(tmp_path)), nprocs=world, join=True)
    vals, info = _fake_sweep(S)
    want = dict(zip(sharding.SUMMARY_FIELDS, sharding.local_summary(vals, info).tolist()))
    got = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert got[0]["range"] == (0, 51) and got[1]["range"] == (51, 101)
    for g in got:
        for k in sharding.SUMMARY_FIELDS:
            assert abs(g["summary"][k] - want[k]) < 1e-12, (k, g["summary"][k], want[k])
    assert want["n_failed"] > 0 and want["n_subjects"] == S
    # the shared-hyper-parameter gradient rides in the same collective: sum over the successful subjects of all ranks
    hg = torch.from_numpy(np.random.RandomState(1).standard_normal((S, 9)))
    hwant = sharding.local_hyper_grad(hg, info).tolist()
    for g in got:
        assert g["summary2"] == g["summary"]
        assert list(g["hyper"]) == ["mu_tilde_l", "alpha_tilde_l", "beta_tilde_l", "mu_L", "alpha_L", "beta_L", "a", "b"]
        for i, k in enumerate(g["hyper"]):
            assert abs(g["hyper"][k] - hwant[i]) < 1e-12
