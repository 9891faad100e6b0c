# last check of a round: GPU tests, smoke, the default bench line
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_err.log
python -c "
import json; d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['traffic_source'], 'cpu', d['cpu_baseline']['value'], d['clocks'])"
