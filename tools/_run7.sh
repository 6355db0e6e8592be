mkdir -p gpurun_out
for g in default 0; do
  if [ $g = 0 ]; then export NMGP_TAKAHASHI_GUARD=0; else unset NMGP_TAKAHASHI_GUARD; fi
  echo "== guard=$g" | tee -a gpurun_out/lat2_r02.txt
  timeout 300 python tools/run_latency2.py nonseparable 100 6 1250 10 2>&1 | grep "^{" | tee -a gpurun_out/lat2_r02.txt
  timeout 300 python tools/run_latency2.py nonseparable 100 6 2500 10 2>&1 | grep "^{" | tee -a gpurun_out/lat2_r02.txt
done
unset NMGP_TAKAHASHI_GUARD
for w in C1 C2 C3; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_${w}_r02a.json 2> gpurun_out/bench_${w}_err.txt; tail -c 1500 gpurun_out/bench_${w}_r02a.json; tail -3 gpurun_out/bench_${w}_err.txt
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_C4_r02a.json 2> gpurun_out/bench_C4_err.txt; tail -c 2500 gpurun_out/bench_C4_r02a.json; tail -3 gpurun_out/bench_C4_err.txt
