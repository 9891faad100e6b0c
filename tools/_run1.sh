python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/quick_bench.py 2>&1 | tail -1
