"""Full-size golden vectors from the UNMODIFIED reference (build container only: needs /root/reference).

    python tests/golden/make_golden_big.py [c3] [n4096] [n8192] [n16384]

BASELINE.json configs[2] (C3: nonseparable M=10, N=500, n = NM = 5000; Prior=True and Prior=False) and the C5 shape
(M=8) at n = 4096, 8192 and the full 16 384 (N=2048): `Utility.logpos.nlogpos_obj_SVC(..., verbose=True)` +
`.backward()`, exactly as `Nonseparable_model_mpisim.py:183-186` calls it.  ~7 s (n=5000) to several minutes and
~30 GB (n=16 384) of the reference's dense inverse + logdet + autograd per case, so every case is written once and
never regenerated.  Each Prior=True file also carries the Prior=False value and gradient (`val_noprior`,
`grad_noprior`) so that the likelihood part of the gradient is pinned at 1e-9 separately from the ill-conditioned
GP-prior terms.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (imports the reference + the symeig shim)
from nonstationary_multivariate_gaussian_process_b200 import synth  # noqa: E402

# key -> (N, M, seed, noise, hyper index)
BIG = {
    "c3": (500, 10, 11, 0.05, 0),
    "n4096": (512, 8, 12, 0.05, 0),
    "n8192": (1024, 8, 13, 0.05, 0),
    "n16384": (2048, 8, 14, 0.05, 0),
}


def main():
    torch.set_num_threads(8)
    want = [a for a in sys.argv[1:] if a in BIG] or ["c3", "n4096", "n8192"]
    manifest = G.load_manifest()
    big = manifest.setdefault("big_cases", [])
    for key in want:
        N, M, seed, noise, hidx = BIG[key]
        name = G.case_name("nonseparable", N, M, seed, hidx, True)
        path = os.path.join(HERE, name + ".npz")
        if not os.path.exists(path):
            t0 = time.time()
            x, Y, _ = synth.sample_subject(N, M, seed)
            pars = synth.start_point("nonseparable", N, M, seed, noise)
            hyper = G.HYPER["nonseparable"][hidx]
            t1 = time.time()
            vals, grad = G.reference_eval("nonseparable", pars, Y, x, hyper, True)
            t2 = time.time()
            vnp, gnp = G.reference_eval("nonseparable", pars, Y, x, hyper, False)
            t3 = time.time()
            G.write_case(name, "nonseparable", N, M, x, Y, pars, hyper, True, vals, grad, val_noprior=vnp[0],
                         grad_noprior=gnp, reference_seconds=np.array([t2 - t1, t3 - t2]))
            print(f"{name}: data {t1 - t0:.1f}s, reference Prior=True {t2 - t1:.1f}s, Prior=False {t3 - t2:.1f}s, "
                  f"-logpost {vals[0]:.6f} loglik {vals[1]:.6f}", flush=True)
        if name not in big:
            big.append(name)
        G.save_manifest(manifest)


if __name__ == "__main__":
    main()
