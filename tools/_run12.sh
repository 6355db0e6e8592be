mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hadamard.py tests/test_gpu_parity_golden.py -m gpu -q -x 2>&1 | tail -3
for mx in 256 100000; do
  echo "== prior warp kernel max N = $mx" | tee -a gpurun_out/hadamard_r02.txt
  NMGP_PRIOR_WARP_MAXN=$mx timeout 300 python tools/run_hadamard.py hadamard_svc 600 6 3000 2>&1 | grep "^{" | tee -a gpurun_out/hadamard_r02.txt
done
timeout 300 python tools/run_hadamard.py hadamard 100 3 10000 2>&1 | grep "^{" | tee -a gpurun_out/hadamard_r02.txt
timeout 300 python tools/run_hadamard.py hadamard_s 200 5 1 2>&1 | grep "^{" | tee -a gpurun_out/hadamard_r02.txt
