"""Golden vectors for the empirical initialisation, produced by running the UNMODIFIED reference
`Utility/empirical_estimation.py:70-133` (`local_estimation`) and `:62-67` (`global_estimation`).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_empirical.py

The reference module imports matplotlib at the top (absent here): a stub package is registered in sys.modules before the
import, nothing else is touched.
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

mpl = types.ModuleType("matplotlib")
mpl.use = lambda *a, **k: None
sys.modules["matplotlib"] = mpl
sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
mpl.pyplot = sys.modules["matplotlib.pyplot"]
from Utility import empirical_estimation as ref  # noqa: E402  (the reference)

from nonstationary_multivariate_gaussian_process_b200 import synth  # noqa: E402

CASES = [(40, 2, 0, 8), (60, 3, 1, 30), (35, 4, 2, 12), (100, 6, 3, 30)]   # N, M, seed, window_size


def main():
    for N, M, seed, w in CASES:
        x, Y, _ = synth.sample_subject(N, M, seed)
        out = ref.local_estimation(x, Y, window_size=w)
        S, L_vec = ref.global_estimation(x, Y)
        lag, sv = ref.SV(x[:9], Y[:9], M - 1)
        name = f"empirical_N{N}_M{M}_s{seed}_w{w}"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), N=N, M=M, x=x, Y=Y, window_size=w, est_sigmas=out[0], est_ls=out[1],
                            smooth_ls=out[2], est_stds=out[3], est_R=out[4], est_B=out[5], est_L_vecs=out[6],
                            est_tilde_sigma2_err=out[7], global_S=S, global_L_vec=np.asarray(L_vec), sv_lag=lag, sv_val=sv)
        print(name, out[1][:3], out[6].shape)


if __name__ == "__main__":
    main()
