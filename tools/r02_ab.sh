# usage (GPU box): bash tools/r02_ab.sh <tag> [tests] -- parity subset, then frame times (full frame and a 135-row band) of the product
# library under the environment toggles in $ROMIS_AB_TOGGLES and of every tuning build in romis_b200/build/lib_*.so
R=$1
if [ "$2" = tests ]; then
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py tests/test_gpu_light_edits.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/tests_$R.log; tail -4 gpurun_out/tests_$R.log
fi
q() { echo "== $*"; env "$@" timeout 120 python tools/quick_bench.py 2>&1 | tail -1; env "$@" timeout 120 python tools/frame_time.py 2>&1 | tail -1; env "$@" timeout 120 python tools/frame_time.py 472 607 2>&1 | tail -1; }
for t in $ROMIS_AB_TOGGLES; do q $t; done
for v in romis_b200/build/lib_*.so; do q ROMIS_GPU_LIB=$PWD/$v; done
