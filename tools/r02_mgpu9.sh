# usage (2-GPU box): bash tools/r02_mgpu9.sh <tag> -- band parity on 2 GPUs (torchrun tests + multi-device contexts), then the C2 line on 2 GPUs
R=$1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
bash tools/r02_mgpu7.sh $R 2 c2
