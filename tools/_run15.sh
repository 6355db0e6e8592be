python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python tools/report_hyper_parity.py 2>&1 | tail -10
timeout 300 python -m pytest tests/test_gpu_units_golden.py tests/test_gpu_two_devices.py -m gpu -q 2>&1 | tail -3
