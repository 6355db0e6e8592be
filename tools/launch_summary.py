"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launch count per kernel.
usage: python tools/launch_summary.py launches.csv [skip_first_n_launches]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1 + skip:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = r[ki].replace("nmgp::<unnamed>::", "").split("(")[0]
    agg[name][0] += 1
    agg[name][1] += v * scale
tot = sum(v[1] for v in agg.values())
print(f"{'ms':>10} {'launches':>8} {'share':>6}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.3f} {v[0]:8d} {100 * v[1] / tot:5.1f}%  {k}")
print(f"{tot:10.3f} total")
