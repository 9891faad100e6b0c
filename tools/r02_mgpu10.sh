# usage (4-GPU box): bash tools/r02_mgpu10.sh <tag> -- the 4-GPU band parity test (interior ranks: two neighbours each) and the C2 line on 4 GPUs
R=$1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "four or two_gpu_bands_equal_the_single" 2>&1 | tail -3
bash tools/r02_mgpu7.sh $R 4 c2
