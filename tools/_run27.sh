#!/bin/bash
for v in 1 2; do
  echo "NMGP_DIAG_MMA=$v"
  NMGP_DIAG_MMA=$v timeout 300 python tools/run_config.py nonseparable 100 6 10000 3 2>&1 | grep "^{"
  NMGP_DIAG_MMA=$v timeout 300 python tools/run_config.py nonseparable 500 10 1 5 2>&1 | grep "^{"
  NMGP_DIAG_MMA=$v timeout 300 python tools/run_config.py nonseparable 2048 8 1 2 2>&1 | grep "^{"
  NMGP_DIAG_MMA=$v timeout 300 python tools/run_config.py nonseparable 100 6 1250 5 2>&1 | grep "^{"
done 2>&1 | tee gpurun_out/diag_mma_ab3.txt
