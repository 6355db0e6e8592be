mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final_err.txt; tail -c 300 gpurun_out/bench_r02_final.json; tail -3 gpurun_out/bench_r02_final_err.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
