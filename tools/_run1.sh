python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/quick_bench.py 2>&1 | tail -3
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_new.json 2> gpurun_out/bench_err.log; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_new.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:v['ms_per_launch'] for k,v in d['roofline']['passes'].items()})"
