#!/bin/bash
bash tools/gpu_check.sh r02_v2
K='regex:^(panel_gemm_kernel|inverse_kernel|diag64|svc_[a-z_]*kernel|kx_kernel|prior_|trace_rows|tile_kernel|trsm_panel|extract_factor)'
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" --csv --log-file gpurun_out/launches_r02_v2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launches_r02_v2.log 2>&1
python tools/launch_table.py gpurun_out/launches_r02_v2.csv 6 | head -30
bash tools/ncu_kernels.sh r02v2 "diag64_mma_kernel:8"
python tools/ncu_summary.py gpurun_out/prof_r02v2_*.ncu-rep 2>&1 | head -40
timeout 300 python tools/lib_baselines.py 2>&1 | tail -40 > gpurun_out/lib_baselines_r02_v2.txt; tail -5 gpurun_out/lib_baselines_r02_v2.txt
