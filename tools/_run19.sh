mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_edge_cases.py tests/test_gpu_parity_golden.py -m gpu -q -x 2>&1 | tail -3
bash tools/ab_variants.sh r02c "10000 1250" main sa3 main sa3
