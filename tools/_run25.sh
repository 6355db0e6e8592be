#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for v in 0 1; do
  echo "NMGP_DIAG_MMA=$v"
  NMGP_DIAG_MMA=$v timeout 300 python tools/run_config.py nonseparable 100 6 10000 3 2>&1 | grep "^{"
  NMGP_DIAG_MMA=$v timeout 300 python tools/run_config.py nonseparable 500 10 1 5 2>&1 | grep "^{"
  NMGP_DIAG_MMA=$v timeout 300 python tools/run_config.py nonseparable 2048 8 1 2 2>&1 | grep "^{"
  NMGP_DIAG_MMA=$v timeout 300 python tools/run_config.py nonseparable 100 6 1250 5 2>&1 | grep "^{"
done 2>&1 | tee gpurun_out/diag_mma_ab.txt
