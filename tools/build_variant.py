"""Build a variant of libnmgp_b200.so for A/B timing on the GPU box.
usage: python tools/build_variant.py NAME [-DFLAG ...] [--rev GITREV]
-> nonstationary_multivariate_gaussian_process_b200/variants/libnmgp_b200_NAME.so  (git-ignored like every .so; travels with gpurun)
Select it with NMGP_B200_LIB=<that path>.  --rev builds the csrc/ of another commit (e.g. the previous round's engine)."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nonstationary_multivariate_gaussian_process_b200 import _lib

args = sys.argv[1:]
name, flags, rev = args[0], [a for a in args[1:] if a.startswith("-D")], None
if "--rev" in args:
    rev = args[args.index("--rev") + 1]
outdir = os.path.join(ROOT, "nonstationary_multivariate_gaussian_process_b200", "variants")
os.makedirs(outdir, exist_ok=True)
out = os.path.join(outdir, f"libnmgp_b200_{name}.so")
csrc = None
if rev:
    tmp = tempfile.mkdtemp(prefix="nmgp_rev_")
    subprocess.check_call(f"git -C {ROOT} archive {rev} nonstationary_multivariate_gaussian_process_b200/csrc include | tar -x -C {tmp}", shell=True)
    csrc = os.path.join(tmp, "nonstationary_multivariate_gaussian_process_b200", "csrc")
print(_lib.build_library(force=True, extra_flags=flags, out_path=out, csrc=csrc))
