mkdir -p gpurun_out
for l2 in 128 256 0; do
  echo "== TMA L2 promotion $l2" | tee -a gpurun_out/tma_l2_r02.txt
  NMGP_TMA_L2=$l2 timeout 300 python tools/run_config.py nonseparable 100 6 10000 5 2>&1 | grep "^{" | tee -a gpurun_out/tma_l2_r02.txt
  NMGP_TMA_L2=$l2 timeout 300 python tools/run_config.py nonseparable 500 10 1 10 2>&1 | grep "^{" | tee -a gpurun_out/tma_l2_r02.txt
done
