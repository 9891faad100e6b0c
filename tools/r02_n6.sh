# usage (GPU box): bash tools/r02_n6.sh <tag> -- R-OMIS parity, then the R-OMIS line with the register-resident 6x6 solve and with the generic one
R=$1
timeout 300 python -m pytest tests/test_gpu_romis.py tests/test_gpu_dropin.py tests/test_gpu_rmis.py -m gpu -x -q 2>&1 | tail -2
bash tools/r02_run.sh ${R} romis
ROMIS_SOLVE_GENERIC=1 bash tools/r02_run.sh ${R}g romis
