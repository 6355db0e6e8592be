mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
bash tools/ab_variants.sh r02e "10000" t0prod main t0prod main
