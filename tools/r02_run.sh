# usage (on the GPU box): bash tools/r02_run.sh <tag> [tests] [bench] [ncu] [ref]
# tests: pytest -m gpu; bench: default bench line; ncu: launch list + --set full capture of the five pass kernels; ref: reference arm
R=$1; shift
for what in "$@"; do
case $what in
tests) python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests_$R.log; tail -3 gpurun_out/tests_$R.log ;;
smoke) python __graft_entry__.py smoke 2>&1 | tail -2 ;;
bench) python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -c 3000 gpurun_out/bench_$R.json ;;
quick) python bench.py --steps 30 --warmup 4 --no-cpu-baseline > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; python tools/show_bench.py gpurun_out/bench_$R.json ;;
ref) python bench.py --impl reference --steps 3 --warmup 2 > gpurun_out/ref_$R.json 2> gpurun_out/ref_$R.err; cat gpurun_out/ref_$R.json ;;
rmis) python bench.py --config rmis --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rmis_$R.json 2> gpurun_out/bench_rmis_$R.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_rmis_$R.json').read().strip().splitlines()[-1]); print('rmis', round(d['ms_per_step'],3), d['roofline']['stages_ms_per_frame'])" ;;
romis) python bench.py --config romis --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_romis_$R.json 2> gpurun_out/bench_romis_$R.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_romis_$R.json').read().strip().splitlines()[-1]); print('romis', round(d['ms_per_step'],3), d['roofline']['stages_ms_per_frame'])" ;;
ncu)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$R.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain_$R.log; continue; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1_$R.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 28 --launch-count 9 -k regex:'primary_kernel|initial_kernel|temporal_kernel|spatial_kernel|shade_kernel' -o gpurun_out/prof_$R -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2_$R.log 2>&1
ls -la gpurun_out/prof_$R.ncu-rep ;;
esac
done
