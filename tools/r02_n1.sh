# usage (GPU box): bash tools/r02_n1.sh <tag> -- initial -> temporal row-group link A/B, parity subset, multi-device R-MIS tests, R-MIS / R-OMIS lines
R=$1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_rmis.py tests/test_gpu_romis.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/tests_${R}_mis.log; tail -4 gpurun_out/tests_${R}_mis.log
ROMIS_AB_TOGGLES="ROMIS_FINE_INITIAL=0 ROMIS_FINE_INITIAL=1" bash tools/r02_ab.sh $R tests
bash tools/r02_run.sh $R rmis romis
