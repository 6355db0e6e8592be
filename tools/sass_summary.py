"""Per-kernel counts of the SASS instructions that identify the Blackwell-native code paths in libnmgp_b200.so
(cuobjdump -sass): DMMA (FP64 tensor core, mma.sync.m8n8k4.f64), UTMALDG (TMA tensor load), SYNCS (mbarrier),
DFMA / MUFU.RSQ64H (FP64 pipe), LDS / STS, and that no UTCMMA / LDTM / STTM (tcgen05) appears -- tcgen05 has no f64 kind.
usage: python tools/sass_summary.py [path/to/lib.so] > profiles/r02_sass_summary.txt   (runs without a GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "nonstationary_multivariate_gaussian_process_b200", "libnmgp_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pats = collections.OrderedDict([("DMMA", r"\bDMMA\b|\bDMMA\."), ("UTMALDG", r"\bUTMALDG"), ("SYNCS", r"\bSYNCS"), ("DFMA", r"\bDFMA\b"),
                                ("RSQ64", r"MUFU\.RSQ64H"), ("LDS", r"\bLDS\b|\bLDS\."), ("STS", r"\bSTS\b|\bSTS\."), ("LDG", r"\bLDG\b|\bLDG\."),
                                ("STG", r"\bSTG\b|\bSTG\."), ("tcgen05", r"UTC.?MMA|\bLDTM|\bSTTM")])
cur, counts = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pats.items():
            if re.search(p, ln):
                counts[cur][k] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines() if counts else []
names = dict(zip(counts, dem)) if len(dem) == len(counts) else {k: k for k in counts}
print(f"# {os.path.basename(lib)}: instruction counts per kernel (cuobjdump -sass, sm_100a); arch lines: " +
      ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", out)))))
print(f"{'DMMA':>6} {'UTMALDG':>8} {'SYNCS':>6} {'DFMA':>6} {'RSQ64':>6} {'LDS':>6} {'STS':>5} {'LDG':>5} {'STG':>5} {'tcgen05':>8}  kernel")
tot = collections.Counter()
for k, c in counts.items():
    nm = re.sub(r"nmgp::\(anonymous namespace\)::|\(anonymous namespace\)::", "", names[k]).split("(")[0].replace("void ", "")
    print(f"{c['DMMA']:6d} {c['UTMALDG']:8d} {c['SYNCS']:6d} {c['DFMA']:6d} {c['RSQ64']:6d} {c['LDS']:6d} {c['STS']:5d} {c['LDG']:5d} {c['STG']:5d} {c['tcgen05']:8d}  {nm}")
    tot.update(c)
print(f"{tot['DMMA']:6d} {tot['UTMALDG']:8d} {tot['SYNCS']:6d} {tot['DFMA']:6d} {tot['RSQ64']:6d} {tot['LDS']:6d} {tot['STS']:5d} {tot['LDG']:5d} {tot['STG']:5d} {tot['tcgen05']:8d}  TOTAL ({len(counts)} kernels)")
