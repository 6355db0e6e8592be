#!/bin/bash
# One GPU-box pass: parity tests, the three large configurations, the bench line, then the ncu launch list of the bench command.
# usage (under gpurun): bash tools/gpu_check.sh TAG
TAG=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.txt
cat gpurun_out/pytest_$TAG.txt
{ python tools/run_config.py nonseparable 500 10 1 5
  python tools/run_config.py nonseparable 2048 8 1 2
  python tools/run_config.py nonseparable 2048 8 8 1
  python tools/run_config.py separable 200 5 1 20
  python tools/run_config.py stationary 50 2 1 20
  python tools/run_config.py nonseparable 100 6 10000 3 ; } 2>&1 | tee gpurun_out/configs_$TAG.txt
{ python tools/run_predict.py 100 6 1 201 100
  python tools/run_predict.py 100 6 64 201 100
  python tools/run_predict.py 500 10 1 201 100 ; } 2>&1 | grep "^{" | tee gpurun_out/predict_$TAG.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_err_$TAG.txt
cat gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_err_$TAG.txt
