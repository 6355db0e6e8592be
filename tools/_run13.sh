mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final_err.txt; tail -c 600 gpurun_out/bench_r02_final.json; tail -3 gpurun_out/bench_r02_final_err.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.json 2>&1; tail -c 400 gpurun_out/bench_r02_reference.json
# ncu launch list of the bench command (times + DRAM bytes), only after the plain run above exited 0
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
python tools/launch_table.py gpurun_out/r02_launches_final.csv 4 > gpurun_out/r02_launches_final.txt 2>&1; cat gpurun_out/r02_launches_final.txt | head -24
bash tools/ncu_kernels.sh r02 "panel_gemm_kernel:32" "panel_gemm_kernel:9" "panel_gemm_kernel:8" "diag64_kernel:8" "svc_contract_kernel" "svc_build_kernel"
python tools/ncu_summary.py gpurun_out/prof_r02_*.ncu-rep > gpurun_out/r02_ncu_full_summary.txt 2>&1; head -60 gpurun_out/r02_ncu_full_summary.txt
timeout 300 ./tools/fp64_peak > gpurun_out/r02_fp64_peak.txt 2>&1; tail -5 gpurun_out/r02_fp64_peak.txt
