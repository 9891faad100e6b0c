# usage: bash tools/_ncu.sh a|b|final   (gpurun brings back at most 64 MiB per call: the captures are split over calls)
set -e
R=${ROUND_TAG:-r01f}
pr() { for c in "$@"; do f=gpurun_out/bench_${c}_n1.json; [ -f $f ] || f=gpurun_out/bench_$c.json; python -c "
import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$c', round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), d['roofline'].get('stages_ms_per_frame') or {k:v['ms_per_launch'] for k,v in d['roofline']['passes'].items()})"; done; }
if [ "$1" = "final" ]; then      # tests, smoke, the default bench line, launch list and the ReSTIR capture of the final state
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_err.log
python bench.py --config rmis --steps 10 --warmup 3 > gpurun_out/bench_rmis.json 2> gpurun_out/bench_rmis.err
python bench.py --config romis --steps 10 --warmup 3 > gpurun_out/bench_romis.json 2> gpurun_out/bench_romis.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$R.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 28 --launch-count 9 -k regex:'primary_kernel|initial_kernel|temporal_kernel|spatial_kernel|shade_kernel' -o gpurun_out/prof_$R -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
pr c2 rmis romis
elif [ "$1" = "a" ]; then
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_err.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$R.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 28 --launch-count 9 -k regex:'primary_kernel|initial_kernel|temporal_kernel|spatial_kernel|shade_kernel' -o gpurun_out/prof_$R -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
pr c2
else
python bench.py --config rmis --steps 10 --warmup 3 > gpurun_out/bench_rmis.json 2> gpurun_out/bench_rmis.err
python bench.py --config romis --steps 10 --warmup 3 > gpurun_out/bench_romis.json 2> gpurun_out/bench_romis.err
for c in c2u c3 c4 c4k; do python bench.py --config $c --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/bench_${c}_n1.json 2> gpurun_out/bench_${c}.err; done
ncu --set full --clock-control none --import-source on --launch-skip 20 --launch-count 4 -k regex:'rmis_neighbours_kernel|romis_accumulate_kernel|romis_solve_kernel' -o gpurun_out/prof_${R}_romis -f python bench.py --config romis --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_romis.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 10 --launch-count 2 -k regex:'rmis_gather_kernel' -o gpurun_out/prof_${R}_rmis -f python bench.py --config rmis --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rmis.log 2>&1
pr c2u c3 c4 c4k rmis romis
fi
