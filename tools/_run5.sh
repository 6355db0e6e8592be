mkdir -p gpurun_out
for mid in 3 5 99; do
  echo "== pipeline mid=$mid" | tee -a gpurun_out/pipe2_r02.txt
  NMGP_PIPE_MID=$mid timeout 300 python tools/run_config.py nonseparable 100 6 10000 5 2>&1 | grep "^{" | tee -a gpurun_out/pipe2_r02.txt
done
for eng in left left_stable recursive; do
  echo "== no pipeline engine=$eng" | tee -a gpurun_out/pipe2_r02.txt
  NMGP_NO_PIPELINE=1 timeout 300 python tools/run_config.py nonseparable 100 6 10000 5 $eng 2>&1 | grep "^{" | tee -a gpurun_out/pipe2_r02.txt
done
timeout 120 python -m pytest tests/test_gpu_units_golden.py -m gpu -q -k positive 2>&1 | tail -5
