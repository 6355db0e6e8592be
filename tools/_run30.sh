#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_v3.json 2> gpurun_out/bench_err_r02_v3.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_v3.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['phase_ms'], d['clocks'], d['parity'])
PY
for w in C3 C2 C1; do python bench.py --workload $w > gpurun_out/bench_${w}_r02_v3.json 2>/dev/null; python - <<PY
import json
d=json.load(open('gpurun_out/bench_${w}_r02_v3.json'))
print('$w', d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('frac'), d.get('parity'))
PY
done
