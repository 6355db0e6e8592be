#!/usr/bin/env python
"""Benchmark of the NMGP batched log-posterior + gradient hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm on the host cores

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): the nonseparable model batched over
10 000 synthetic subjects of M=6 outputs x N=100 time points (n = NM = 600 per subject), subject-sharded over the
ranks ("strong" scaling: the 10 000 subjects are split over the GPUs).  One step = one evaluation of
-log posterior, its components and its gradient for every subject.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "log-posterior+grad evals/sec (batched subjects)"
UNIT = "evals/s"

# hyper-parameters of the reference's simulation drivers (Stationary_model_mpisim.py:86,
# Separable_model_mpisim.py:296-297, Nonseparable_model_mpisim.py:311-312)
HYPER = {
    "stationary": {"mu_tilde_l": 0.0, "sigma_tilde_l": 10.0, "a": 1e-6, "b": 1e-6, "c": 1.0},
    "separable": {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_tilde_sigma": 0.0,
                  "alpha_tilde_sigma": 1.0, "beta_tilde_sigma": 1.0, "a": 1e-2, "b": 1e-2, "c": 0.1},
    "nonseparable": {"mu_tilde_l": 0.0, "alpha_tilde_l": 10.0, "beta_tilde_l": 1.0, "mu_L": 0.0, "alpha_L": 1.0,
                     "beta_L": 1.0, "a": 1e-2, "b": 1e-2},
}
FP64_PEAK_TFLOPS = 37.15   # measured on this pool's B200 with tools/fp64_peak.cu (DMMA.8x8x4), profiles/r01_fp64_peak.txt


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="nonseparable", choices=list(HYPER))
    ap.add_argument("--subjects", type=int, default=10000)
    ap.add_argument("--N", type=int, default=100)
    ap.add_argument("--M", type=int, default=6)
    ap.add_argument("--cpu-subjects-per-core", type=int, default=160)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def algorithmic_flops(model, N, M):
    """Per evaluation: potrf n^3/3 and potri 2n^3/3 on each factorised matrix."""
    if model == "nonseparable":
        n, nmat = N * M, 1
    else:
        n, nmat = N, M
    return nmat * n ** 3 / 3.0, nmat * 2.0 * n ** 3 / 3.0


# ------------------------------------------------------------------------------------------ reference arm
def run_cpu_workers(model, N, M, per_core, cores, first=0):
    """One single-threaded oracle process per host core (the reference's deployment mode); returns evals/s."""
    procs = []
    t0 = time.perf_counter()
    for c in range(cores):
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "cpu_bench.py"), "--model", model, "--N", str(N), "--M",
               str(M), "--first", str(first + c * per_core), "--count", str(per_core), "--threads", "1"]
        env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT))
    secs = []
    for p in procs:
        out, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("cpu worker failed: " + err[-2000:])
        secs.append(json.loads(out.strip().splitlines()[-1])["seconds"])
    wall = time.perf_counter() - t0
    n_eval = per_core * cores
    return n_eval / max(secs), max(secs), wall, n_eval


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per = a.cpu_subjects_per_core
    for _ in range(min(a.warmup, 1)):
        run_cpu_workers(a.model, a.N, a.M, 1, cores)
    step_secs, n_eval = [], per * cores
    for s in range(a.steps):
        _, sec, _, _ = run_cpu_workers(a.model, a.N, a.M, per, cores, first=s * n_eval)
        step_secs.append(sec)
    total = sum(step_secs)
    value = n_eval * a.steps / total
    sample = (f"{n_eval} of the {a.subjects} subjects per step, {cores} single-threaded processes x {per} subjects "
              f"(oracle/nmgp_oracle.py: dense inverse + logdet + autograd, the reference's algorithm)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, a.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(a, world):
    n = a.N * a.M if a.model == "nonseparable" else a.N
    return {"workload": f"{a.model} model, {a.subjects} synthetic subjects (SIM_code/sim.py recipe) x (M={a.M}, N={a.N}), "
                        f"value+gradient per subject, subject-sharded over {world} GPU(s)",
            "subjects": a.subjects, "M": a.M, "N": a.N, "matrix_dim": n,
            "cache": "per-step working set (covariance workspace) far exceeds the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------ inputs
def make_inputs(a, lo, hi, device):
    """x [S,N], pars [S,P] (driver-like start point) on the host; Y [S,N,M] drawn on the GPU from the generating
    covariance (SIM_code/sim.py:256-264) with library Cholesky -- input generation, outside every timed region."""
    import ctypes

    import numpy as np
    import torch

    from nonstationary_multivariate_gaussian_process_b200 import _lib, synth
    S, N, M = hi - lo, a.N, a.M
    T = M * (M + 1) // 2
    x = np.empty((S, N))
    pars = np.empty((S, {"stationary": T + 3, "separable": 2 * N + T + 1, "nonseparable": N + N * T + 1}[a.model]))
    for i, s in enumerate(range(lo, hi)):
        x[i] = synth.truth(N, M, s)[0]
        pars[i] = synth.start_point(a.model, N, M, s)
    lib = _lib.load_library()
    n = N * M
    Y = torch.empty((S, N, M), dtype=torch.float64, device=device)
    gen = torch.Generator(device=device)
    stream = torch.cuda.current_stream().cuda_stream
    # Observations are drawn in fixed GLOBAL blocks of subjects, one seed per block, whatever the sharding: every subject gets
    # the same Y at 1, 2, 4 or 8 ranks, so `sweep_summary` of the runs is comparable (a rank generates the whole block it
    # overlaps and keeps its part).
    blk = max(1, min(250, int((2 << 30) // (n * n * 8))))
    for g0 in range((lo // blk) * blk, hi, blk):
        g1 = min(g0 + blk, a.subjects)
        cs = g1 - g0
        xb = np.empty((cs, N))
        tb = np.empty((cs, N + N * T + 1))
        for i, s in enumerate(range(g0, g1)):
            xs, tl, uL, ts2 = synth.truth(N, M, s)
            xb[i] = xs
            tb[i] = np.concatenate([tl, uL.reshape(-1), [ts2]])
        xd = torch.from_numpy(xb).to(device)
        td = torch.from_numpy(tb).to(device)
        cov = torch.empty((cs, n, n), dtype=torch.float64, device=device)
        _lib.check(lib.nmgp_nonseparable_cov(xd.data_ptr(), td.data_ptr(), cs, N, M, cov.data_ptr(),
                                             ctypes.c_void_p(stream)), "nmgp_nonseparable_cov")
        Lc = torch.linalg.cholesky(cov)
        gen.manual_seed(12345 + g0)
        z = torch.randn((cs, n, 1), dtype=torch.float64, device=device, generator=gen)
        y = (Lc @ z).squeeze(-1).view(cs, M, N).transpose(1, 2)    # output-major (m*N+i), sim.py:263-264
        a0, a1 = max(g0, lo), min(g1, hi)
        Y[a0 - lo:a1 - lo] = y[a0 - g0:a1 - g0]
        del cov, Lc
    return torch.from_numpy(x), Y, torch.from_numpy(pars)


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        # median of the samples taken under load (upper half: the sampler also sees the idle edges)
        med = sm[(3 * len(sm)) // 4] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ our arm
def ours(a):
    import torch
    import torch.distributed as dist

    from nonstationary_multivariate_gaussian_process_b200 import sharding
    from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    lo, hi = sharding.shard_range(a.subjects, rank, world)
    S = hi - lo
    x, Y, pars_h = make_inputs(a, lo, hi, device)
    plan = LogPosteriorPlan(a.model, x, Y, HYPER[a.model], prior=True, device=device)
    pars_d = pars_h.to(device)
    P = plan.P

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    hyper_names = plan.hyper_names()

    def step_device():
        # one sweep: value + gradient of every local subject and, in the same pass, the gradient with respect to the
        # hyper-parameters shared by all subjects; the only collective is one all-reduce of 17 doubles
        vals, grad, hgrad, info = plan.value_grad_and_hyper_grad(pars_d)
        vec = sharding.local_sweep_vector(vals, info, hgrad)
        if world > 1:
            dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        return vec

    # ---- device-resident timing (value)
    for _ in range(a.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        summary = step_device()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    launches = plan.last_launches * a.steps
    sv = summary.tolist()
    summary = dict(zip(sharding.SUMMARY_FIELDS, sv[:8]))
    shared_hyper_grad = dict(zip(hyper_names, sv[8:8 + len(hyper_names)]))

    # ---- end to end through the host-buffer C call (pinned host memory in, results out)
    e2e = None
    if not a.no_e2e:
        pin = plan.pinned_pars()
        pin.copy_(pars_h)
        for _ in range(min(a.warmup, 3)):
            plan.value_and_grad_host(pin, pinned_io=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            vals_h, grad_h, info_h = plan.value_and_grad_host(pin, pinned_io=True)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e = {"value": a.subjects * a.steps / dt, "unit": UNIT, "h2d_bytes_per_step": S * P * 8,
               "d2h_bytes_per_step": S * (P * 8 + 6 * 8 + 4), "ms_per_step": 1e3 * dt / a.steps,
               "path": "nmgp_logpost_grad_host (pinned host pars -> vals, grad, info on the host)"}

    # ---- live per-phase timing for the roofline (CUDA events inside the C call, separate from `value`)
    phases = None
    for _ in range(3):
        phases, _, _, _ = plan.profile(pars_d)
    f_potrf, f_potri = algorithmic_flops(a.model, a.N, a.M)
    t_fact = (phases["potrf"] + phases["potri"]) * 1e-3
    achieved = S * (f_potrf + f_potri) / t_fact / 1e12 if t_fact > 0 else 0.0
    # DRAM traffic of the same kernels from the committed ncu pass (profiles/r01_launches_final.txt): 353 GB per sweep of
    # 10 000 subjects over 168 engine launches (3 chunks x 56) -> bytes per launch, scaled to this rank's subjects
    traffic = 353e9 / 168.0 * (S / 10000.0) if (a.model == "nonseparable" and a.N == 100 and a.M == 6) else None
    roofline = {"bound": "tensor", "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": achieved / FP64_PEAK_TFLOPS, "traffic": traffic,
                "traffic_note": "average dram__bytes_read+write per launch over the 168 panel_gemm/diag64 launches of one "
                                "sweep (ncu, profiles/r01_launches_final.txt); algorithmic minimum (every tile read once "
                                "per use from L2-missing operands) is ~250 GB per sweep",
                "kernel": "panel_gemm_kernel<mode> (left-looking potrf + Takahashi inverse: 64x64 DMMA.8x8x4 tiles fed by "
                          "TMA, cp.async.bulk.tensor + 128B swizzle + mbarrier ring) + diag64_kernel; achieved = S*n^3 flop / "
                          "(t_potrf + t_potri), phase times from CUDA events inside nmgp_logpost_grad_profile",
                "peak_source": "FP64 DMMA peak measured on this pool's B200 (tools/fp64_peak.cu, "
                               "profiles/r01_fp64_peak.txt); MEASURED_PEAKS.json has no FP64 entry",
                "algorithmic_flops_per_eval": f_potrf + f_potri,
                "potrf_tflops": S * f_potrf / (phases["potrf"] * 1e-3) / 1e12 if phases["potrf"] > 0 else None,
                "phase_ms": phases}

    if rank == 0:
        line = {
            "metric": METRIC, "value": a.subjects * a.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_total / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline,
            "sweep_summary": {k: summary[k] for k in ("neg_logpost", "n_failed", "n_subjects")},
            "shared_hyper_grad": shared_hyper_grad,
            "plan": {"chunk": plan.chunk, "device_bytes": plan.device_bytes, "subjects_this_rank": S},
        }
        if world == 1 and not a.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, sec, wall, n_eval = run_cpu_workers(a.model, a.N, a.M, a.cpu_subjects_per_core, cores)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{n_eval} subjects of the same workload: {cores} single-threaded processes x "
                          f"{a.cpu_subjects_per_core} (oracle/nmgp_oracle.py, the reference's algorithm), {sec:.1f} s"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours(a)


if __name__ == "__main__":
    main()
