"""Per-component parity of the CUDA path against every golden vector, for every engine (run on the GPU box).

    python tools/report_parity.py [--json OUT.json] [--big]

Prints one row per (case, engine): relative errors of the total, the likelihood, the prior components, the gradient
(2-norm), the Prior=False gradient against the reference's Prior=False gradient (`grad_noprior` of the fixture), and -- for
the cases with a 50-digit mpmath truth (tests/golden/truth/) -- the distance of the GP-prior value / gradient from the
exact answer, for this repository and for the reference.  `--json` also writes the numbers; tools/update_tolerances.py
turns them into the per-fixture bounds of tests/golden/MANIFEST.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from conftest import GOLDEN_DIR, big_cases, golden_cases, load_golden, prior_blocks, rel_err
from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan

ENGINES = ["auto", "left", "left_stable", "recursive"]


def run(g, engine, prior):
    plan = LogPosteriorPlan(g["model"], g["x"], g["Y"], g["hyper"], prior=prior)
    plan.set_engine(engine)
    vals, grad, info = plan.value_and_grad_host(torch.from_numpy(g["pars"]))
    plan.close()
    return vals.numpy()[0], grad.numpy()[0], int(info[0])


def main():
    out_path = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    names = golden_cases() + (big_cases() if "--big" in sys.argv else [])
    rows = []
    print(f"{'case':34s} {'engine':11s} {'total':>8s} {'loglik':>8s} {'priors':>8s} {'grad':>8s} {'grad_np':>8s} "
          f"{'lp:ours':>8s} {'lp:ref':>8s} {'dlp:ours':>8s} {'dlp:ref':>8s}")
    for name in names:
        g = load_golden(name)
        truth_path = os.path.join(GOLDEN_DIR, "truth", name + ".npz")
        truth = np.load(truth_path) if os.path.exists(truth_path) else None
        for engine in ENGINES:
            vals, grad, info = run(g, engine, g["prior"])
            ref = g["vals"]
            row = {"case": name, "engine": engine, "info": info, "total": rel_err(vals[0], ref[0]),
                   "loglik": rel_err(vals[1], ref[1]) if len(ref) > 1 else None,
                   "prior": max([rel_err(vals[k], ref[k]) for k in range(2, len(ref))], default=None),
                   "grad": rel_err(grad, g["grad"]), "grad_noprior": None}
            if g["prior"] and "grad_noprior" in g:
                v0, g0, i0 = run(g, engine, False)
                row["grad_noprior"] = rel_err(g0, g["grad_noprior"])
                row["val_noprior"] = rel_err(v0[0], g["val_noprior"])
                row["info"] = info or i0
                if truth is not None:
                    blocks = prior_blocks(g)
                    for k, (sl, key) in enumerate(blocks):
                        t = -truth["dlp%d" % k].reshape(-1)
                        row["dlp%d_ours" % k] = rel_err((grad - g0)[sl], t)
                        row["dlp%d_ref" % k] = rel_err((g["grad"] - g["grad_noprior"])[sl], t)
                        row["lp%d_ours" % k] = rel_err(vals[2 + k], float(truth["lp%d" % k]))
                        row["lp%d_ref" % k] = rel_err(ref[2 + k], float(truth["lp%d" % k]))
            rows.append(row)

            def f(v):
                return "     nan" if v is None else f"{v:8.1e}"
            tl = [max(row.get(f"{q}{k}_{w}", 0.0) for k in (0, 1)) if f"{q}0_{w}" in row else None
                  for q in ("lp", "dlp") for w in ("ours", "ref")]
            print(f"{name:34s} {engine:11s} {f(row['total'])} {f(row['loglik'])} {f(row['prior'])} {f(row['grad'])} "
                  f"{f(row['grad_noprior'])} {f(tl[0])} {f(tl[1])} {f(tl[2])} {f(tl[3])}" + ("" if row["info"] == 0 else "  INFO!=0"),
                  flush=True)
    if out_path:
        with open(out_path, "w") as fh:
            json.dump(rows, fh, indent=0)


if __name__ == "__main__":
    main()
