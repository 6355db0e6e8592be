mkdir -p gpurun_out
timeout 600 python tools/run_takahashi_stress.py 2>&1 | tee gpurun_out/takahashi_stress_r02.txt
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_golden.py --deselect tests/test_gpu_edge_cases.py::test_takahashi_sweep_on_ill_conditioned_covariances_at_its_size_limit 2>&1 | tail -15 | tee gpurun_out/pytest_r02d.txt
for nc in 4096 2048 1300; do for np_ in 0 1; do
  echo "== chunk_matrices=$nc no_pipeline=$np_" | tee -a gpurun_out/pipe_r02.txt
  if [ $np_ = 1 ]; then export NMGP_NO_PIPELINE=1; else unset NMGP_NO_PIPELINE; fi
  NMGP_CHUNK_MATRICES=$nc timeout 300 python tools/run_config.py nonseparable 100 6 10000 5 2>&1 | grep "^{\|plan:" | tee -a gpurun_out/pipe_r02.txt
done; done
unset NMGP_NO_PIPELINE
