set -e
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r01d.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 28 --launch-count 9 -k regex:'primary_kernel|initial_kernel|temporal_kernel|spatial_kernel|shade_kernel' -o gpurun_out/prof_r01d -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/
