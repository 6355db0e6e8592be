"""Per-kernel time / launches / DRAM traffic from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` launch list.  usage: python tools/launch_table.py launches.csv n_evaluations"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nev = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].replace("nmgp::<unnamed>::", "").replace("void ", "").split("(")[0]
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    if r[mi] == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[r[ui]]
        cnt[name] += 1
    else:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ui]]
    agg[name][r[mi]] += v
tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
print(f"{'ms/eval':>8} {'launches':>8} {'share':>6} {'GB/eval':>8} {'TB/s':>6}  kernel   ({nev:g} evaluations captured, averaged; "
      "times are cold-cache and serialised under ncu: compare shares)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    t = a["gpu__time_duration.sum"]
    b = a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]
    if t / tot < 0.0005:
        continue
    print(f"{t / nev:8.2f} {cnt[k] / nev:8.1f} {100 * t / tot:5.1f}% {b / nev / 1e9:8.2f} {b / t / 1e9 if t else 0:6.2f}  {k}")
print(f"{tot / nev:8.2f} ms per evaluation (sum of kernel times)")
