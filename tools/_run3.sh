mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_golden.py 2>&1 | tail -15 | tee gpurun_out/pytest_r02c.txt
timeout 900 python tools/report_parity.py --big --json gpurun_out/parity_r02.json 2>&1 | tee gpurun_out/parity_r02.txt | tail -120
