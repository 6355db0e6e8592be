mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_c4_n8.json 2> gpurun_out/bench_c4_n8_err.txt
tail -c 400 gpurun_out/bench_c4_n8.json; tail -3 gpurun_out/bench_c4_n8_err.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2_err.txt
tail -c 300 gpurun_out/bench_c4_n2.json
timeout 300 python -m pytest tests/test_gpu_two_devices.py -m gpu -q 2>&1 | tail -2
