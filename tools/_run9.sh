mkdir -p gpurun_out
for w in 2 4 6 8; do
  echo "== W=$w" | tee -a gpurun_out/diag128w_r02.txt
  NMGP_POTRF_W=$w timeout 300 python tools/run_config.py nonseparable 500 10 1 10 2>&1 | grep "^{" | tee -a gpurun_out/diag128w_r02.txt
  NMGP_POTRF_W=$w timeout 300 python tools/run_config.py nonseparable 2048 8 1 3 2>&1 | grep "^{" | tee -a gpurun_out/diag128w_r02.txt
done
# kernel durations of one C3 evaluation (warm): ncu launch list
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/c3_launches_r02.csv python tools/run_config.py nonseparable 500 10 1 1 > /dev/null 2>&1
python tools/launch_table.py gpurun_out/c3_launches_r02.csv 4 2>&1 | head -30 | tee gpurun_out/c3_launch_table_r02.txt
