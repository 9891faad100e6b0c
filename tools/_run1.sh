python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/quick_bench.py 2>&1 | tail -2
python bench.py --config romis --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_romis2.json 2> gpurun_out/bench_romis.err; python -c "
import json; d=json.load(open('gpurun_out/bench_romis2.json')); print(d['value'], d['ms_per_step'], d['roofline']['stages_ms_per_frame'])"
