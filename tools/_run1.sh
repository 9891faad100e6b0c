python -m pytest tests/test_gpu_romis.py -m gpu -x -q 2>&1 | tail -1
for i in 1 2; do python bench.py --config romis --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('romis', d['ms_per_step'], d['roofline']['stages_ms_per_frame'])"; done
