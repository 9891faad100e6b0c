pr() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['ms_per_step'],3), d['roofline']['stages_ms_per_frame'])" $1 $2; }
python -m pytest tests/test_gpu_rmis.py tests/test_gpu_romis.py tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -3
python tools/quick_bench.py 2>&1 | tail -1
python bench.py --config romis --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/t_romis.json 2>/dev/null; pr gpurun_out/t_romis.json romis
python bench.py --config rmis --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/t_rmis.json 2>/dev/null; pr gpurun_out/t_rmis.json rmis
