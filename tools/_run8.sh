mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_full_size.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity_golden.py -m gpu -q -x -k "full_size" 2>&1 | tail -4
for nd in 0 1; do
  if [ $nd = 1 ]; then export NMGP_NO_DIAG128=1; else unset NMGP_NO_DIAG128; fi
  echo "== NO_DIAG128=$nd" | tee -a gpurun_out/diag128_r02.txt
  timeout 300 python tools/run_config.py nonseparable 500 10 1 10 2>&1 | grep "^{" | tee -a gpurun_out/diag128_r02.txt
  timeout 300 python tools/run_config.py nonseparable 2048 8 1 3 2>&1 | grep "^{" | tee -a gpurun_out/diag128_r02.txt
  timeout 300 python tools/run_config.py nonseparable 1024 8 1 5 2>&1 | grep "^{" | tee -a gpurun_out/diag128_r02.txt
done
unset NMGP_NO_DIAG128
timeout 600 python tools/lib_baselines.py 2>&1 | tee gpurun_out/lib_baselines_r02.txt | tail -20
