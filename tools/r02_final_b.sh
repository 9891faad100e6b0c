# usage (GPU box): bash tools/r02_final_b.sh <tag> -- R-MIS / R-OMIS lines, ncu captures of their kernels, the drop-in's own e2e
R=$1
bash tools/r02_run.sh $R rmis romis
ncu --set full --clock-control none --import-source on --launch-skip 20 --launch-count 4 -k regex:'rmis_neighbours_kernel|romis_accumulate_kernel|romis_solve_kernel' -o gpurun_out/prof_${R}_romis -f python bench.py --config romis --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_romis_$R.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 10 --launch-count 2 -k regex:'rmis_gather_kernel' -o gpurun_out/prof_${R}_rmis -f python bench.py --config rmis --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rmis_$R.log 2>&1
ls -la gpurun_out/prof_${R}_r*mis.ncu-rep
timeout 200 python tools/dropin_e2e.py 2>&1 | tail -1
for c in c2u c3 c4k; do timeout 120 python bench.py --config $c --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/bench_${R}_${c}_n1.json 2> gpurun_out/bench_${R}_${c}_n1.err; python tools/show_bench.py gpurun_out/bench_${R}_${c}_n1.json; done
