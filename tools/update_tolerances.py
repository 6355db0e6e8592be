"""Turn a parity report (tools/report_parity.py --json) into the per-fixture bounds of tests/golden/MANIFEST.json.

    python tools/update_tolerances.py gpurun_out/parity_r02.json

For every Prior=True fixture: bound = 3 x the largest error measured over all engines, for the total, the prior components
and the gradient.  tests/conftest.py:tolerances() never lets a bound drop below north_star's 1e-9."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = json.load(open(sys.argv[1]))
path = os.path.join(ROOT, "tests", "golden", "MANIFEST.json")
m = json.load(open(path))
worst = {}
for r in rows:
    w = worst.setdefault(r["case"], {"total": 0.0, "prior": 0.0, "grad": 0.0})
    for k in w:
        if r.get(k) is not None:
            w[k] = max(w[k], float(r[k]))
m["tolerances"] = {c: {k: float(f"{3.0 * v:.2e}") for k, v in w.items()} for c, w in sorted(worst.items())
                   if c.endswith("_p")}
m["tolerances_source"] = ("3 x the largest relative error over the engines auto/left/left_stable/recursive measured on B200 "
                          f"({os.path.basename(sys.argv[1])}, tools/report_parity.py); floor 1e-9 applied by tests/conftest.py")
json.dump(m, open(path, "w"), indent=1)
print(json.dumps(m["tolerances"], indent=1))
