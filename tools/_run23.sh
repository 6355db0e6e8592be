mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
V=nonstationary_multivariate_gaussian_process_b200/variants
for name in invsync main invsync main; do
  if [ "$name" = "main" ]; then unset NMGP_B200_LIB; else export NMGP_B200_LIB=$V/libnmgp_b200_$name.so; fi
  echo "== variant $name" | tee -a gpurun_out/ab_r02f.txt
  timeout 300 python tools/run_config.py nonseparable 500 10 1 10 2>&1 | grep "^{" | tee -a gpurun_out/ab_r02f.txt
  timeout 300 python tools/run_config.py nonseparable 2048 8 2 2 2>&1 | grep "^{" | tee -a gpurun_out/ab_r02f.txt
done
