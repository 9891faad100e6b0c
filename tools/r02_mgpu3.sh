# usage (N-GPU box): bash tools/r02_mgpu3.sh <tag> <N> -- multi-GPU parity (row-group counters forced on and default), then the bench line at N GPUs
R=$1; N=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
ROMIS_FINE=1 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_light_edits.py -m gpu -x -q 2>&1 | tail -3
ROMIS_FINE=1 timeout 300 $TR tools/check_bands_multi_gpu.py --width 1920 --height 1080 --frames 4 2>&1 | grep -v "^W\|warn" | tail -4
timeout 300 $TR tools/check_bands_multi_gpu.py --width 1280 --height 720 --frames 4 --edit-lights 2>&1 | grep -v "^W\|warn" | tail -2
b() { tag=$1; shift; env "$@" timeout 300 $TR bench.py --gpus $N --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${R}_${tag}_n$N.json 2> gpurun_out/bench_${R}_${tag}_n$N.err; echo "== $tag N=$N"; python tools/show_bench.py gpurun_out/bench_${R}_${tag}_n$N.json; }
b auto X=1
b fine ROMIS_FINE=1
