set -e
python bench.py --config romis --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_romis.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 20 --launch-count 4 -k regex:'rmis_neighbours_kernel|romis_accumulate_kernel|romis_solve_kernel' -o gpurun_out/prof_r01d_romis -f python bench.py --config romis --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_romis.log 2>&1
python bench.py --config rmis --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_rmis.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 10 --launch-count 2 -k regex:'rmis_gather_kernel' -o gpurun_out/prof_r01d_rmis -f python bench.py --config rmis --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rmis.log 2>&1
ls -la gpurun_out/*.ncu-rep
