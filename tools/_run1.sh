python tools/quick_bench.py 2>&1 | tail -1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -x -q 2>&1 | tail -2
