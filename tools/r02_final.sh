# usage (GPU box): bash tools/r02_final.sh <tag> -- last check of the round: GPU tests, smoke, both bench arms, R-MIS / R-OMIS lines and captures, drop-in e2e
R=$1
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/tests_$R.log; tail -2 gpurun_out/tests_$R.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python bench.py --impl reference --steps 2 --warmup 2 > gpurun_out/ref_$R.json 2> gpurun_out/ref_$R.err; python -c "
import json; d=json.loads(open('gpurun_out/ref_$R.json').read().strip().splitlines()[-1]); print('reference arm', d['value'], d['cpu_baseline']['sample'])" || tail -3 gpurun_out/ref_$R.err
timeout 300 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; python tools/show_bench.py gpurun_out/bench_$R.json || tail -3 gpurun_out/bench_$R.err
bash tools/r02_run.sh $R rmis romis
timeout 200 ncu --set full --clock-control none --import-source on --launch-skip 20 --launch-count 4 -k regex:'rmis_neighbours_kernel|romis_accumulate_kernel|romis_solve_kernel' -o gpurun_out/prof_${R}_romis -f python bench.py --config romis --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_romis_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on --launch-skip 10 --launch-count 2 -k regex:'rmis_gather_kernel' -o gpurun_out/prof_${R}_rmis -f python bench.py --config rmis --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rmis_$R.log 2>&1
ls -la gpurun_out/prof_${R}_r*mis.ncu-rep
timeout 200 python tools/dropin_e2e.py 2>&1 | tail -1
