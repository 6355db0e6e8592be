mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload C5 --steps 3 --warmup 3 > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8_err.txt
tail -c 3000 gpurun_out/bench_c5_n8.json; tail -5 gpurun_out/bench_c5_n8_err.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_c4_n8.json 2> gpurun_out/bench_c4_n8_err.txt
tail -c 2500 gpurun_out/bench_c4_n8.json; tail -5 gpurun_out/bench_c4_n8_err.txt
