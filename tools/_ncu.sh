set -e
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_err.log
python bench.py --config rmis --steps 10 --warmup 3 > gpurun_out/bench_rmis.json 2> gpurun_out/bench_rmis.err
python bench.py --config romis --steps 10 --warmup 3 > gpurun_out/bench_romis.json 2> gpurun_out/bench_romis.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r01d.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 28 --launch-count 9 -k regex:'primary_kernel|initial_kernel|temporal_kernel|spatial_kernel|shade_kernel' -o gpurun_out/prof_r01d -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 20 --launch-count 4 -k regex:'rmis_neighbours_kernel|romis_accumulate_kernel|romis_solve_kernel' -o gpurun_out/prof_r01d_romis -f python bench.py --config romis --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_romis.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 10 --launch-count 2 -k regex:'rmis_gather_kernel' -o gpurun_out/prof_r01d_rmis -f python bench.py --config rmis --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rmis.log 2>&1
cut -c1-300 gpurun_out/bench_c2.json
