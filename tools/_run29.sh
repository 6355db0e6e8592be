#!/bin/bash
NMGP_DIAG128=1 timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
  echo "NMGP_DIAG128=$v"
  NMGP_DIAG128=$v timeout 300 python tools/run_config.py nonseparable 500 10 1 5 2>&1 | grep "^{" | cut -c1-330
  NMGP_DIAG128=$v timeout 300 python tools/run_config.py nonseparable 2048 8 1 2 2>&1 | grep "^{" | cut -c1-330
done 2>&1 | tee gpurun_out/diag128_mma.txt
