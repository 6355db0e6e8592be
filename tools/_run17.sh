mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_edge_cases.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for st in 0 1; do
  if [ $st = 1 ]; then export NMGP_LL_STATIC=1; else unset NMGP_LL_STATIC; fi
  echo "== static=$st" | tee -a gpurun_out/dyn_r02.txt
  for S in 10000 1250; do timeout 300 python tools/run_config.py nonseparable 100 6 $S 5 2>&1 | grep "^{" | tee -a gpurun_out/dyn_r02.txt; done
done
