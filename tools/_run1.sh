python -m pytest tests/test_gpu_rmis.py tests/test_gpu_romis.py -m gpu -x -q 2>&1 | tail -2
python bench.py --config rmis --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rmis', d['ms_per_step'], d['roofline']['stages_ms_per_frame'])"
