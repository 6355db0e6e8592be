mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
bash tools/ab_variants.sh r02d "10000" noragged main noragged main
