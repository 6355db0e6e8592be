mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_r02f.txt
for S in 10000 1250; do timeout 300 python tools/run_config.py nonseparable 100 6 $S 5 2>&1 | grep "^{" | tee -a gpurun_out/lds_r02.txt; done
timeout 300 python tools/run_config.py nonseparable 500 10 1 10 2>&1 | grep "^{" | tee -a gpurun_out/lds_r02.txt
timeout 300 python tools/run_config.py nonseparable 2048 8 4 2 2>&1 | grep "^{" | tee -a gpurun_out/lds_r02.txt
