"""Per-slot parity of the shared-hyper-parameter gradient against the reference's autograd (tests/golden/hyper_*.npz) and of
smoke()'s checks (run on the GPU box).  usage: python tools/report_hyper_parity.py"""
import glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nonstationary_multivariate_gaussian_process_b200.batched import LogPosteriorPlan, HYPER_SPEC

for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "hyper_*.npz"))):
    z = np.load(path, allow_pickle=False)
    model, hyper = str(z["model"]), json.loads(str(z["hyper"]))
    plan = LogPosteriorPlan(model, z["x"], z["Y"], hyper)
    hg = plan.hyper_grad(torch.from_numpy(z["pars"]).cuda()).cpu().numpy()
    names = [k for k, _ in HYPER_SPEC[model]]
    ref = z["hgrad"]
    err = np.abs(hg[:, :len(names)] - ref[:, :len(names)]).max(0) / np.maximum(np.abs(ref[:, :len(names)]).max(0), 1e-300)
    print(os.path.basename(path)[:-4], " ".join(f"{n}={e:.1e}" for n, e in zip(names, err)))
    plan.close()
