mkdir -p gpurun_out
echo "# raw Takahashi sweep (NMGP_TAKAHASHI_GUARD=0) vs W^T W ('left_stable'), both against the CPU oracle" > gpurun_out/takahashi_stress_r02.txt
NMGP_TAKAHASHI_GUARD=0 timeout 600 python tools/run_takahashi_stress.py 2>&1 | tee -a gpurun_out/takahashi_stress_r02.txt | tail -3
echo "# guarded sweep (default)" >> gpurun_out/takahashi_stress_r02.txt
timeout 600 python tools/run_takahashi_stress.py 2>&1 | tee -a gpurun_out/takahashi_stress_r02.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_r02e.txt
for g in default 0; do
  echo "== guard=$g" | tee -a gpurun_out/guard_r02.txt
  if [ $g = 0 ]; then export NMGP_TAKAHASHI_GUARD=0; else unset NMGP_TAKAHASHI_GUARD; fi
  for S in 10000 1250; do timeout 300 python tools/run_config.py nonseparable 100 6 $S 5 2>&1 | grep "^{" | tee -a gpurun_out/guard_r02.txt; done
done
