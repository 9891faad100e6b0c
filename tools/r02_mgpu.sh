# usage (on an N-GPU box): bash tools/r02_mgpu.sh <tag> <N> [check] [tests] [bench] [c3] [diag]
R=$1; N=$2; shift; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for what in "$@"; do
case $what in
check) timeout 300 $TR tools/check_bands_multi_gpu.py --width 1920 --height 1080 --frames 3 2>&1 | grep -v "^W\|warn" | tail -4
       timeout 300 $TR tools/check_bands_multi_gpu.py --width 1280 --height 720 --frames 4 --edit-lights 2>&1 | grep -v "^W\|warn" | tail -4 ;;
tests) timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_light_edits.py -m gpu -x -q 2>&1 | tail -8 ;;
bench) timeout 600 $TR bench.py --gpus $N --steps 60 --warmup 5 > gpurun_out/bench_${R}_n$N.json 2> gpurun_out/bench_${R}_n$N.err; python tools/show_bench.py gpurun_out/bench_${R}_n$N.json; tail -3 gpurun_out/bench_${R}_n$N.err ;;
bench1) timeout 600 python bench.py --gpus 1 --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${R}_n1.json 2> gpurun_out/bench_${R}_n1.err; python tools/show_bench.py gpurun_out/bench_${R}_n1.json ;;
c3) timeout 900 $TR bench.py --gpus $N --steps 30 --warmup 5 --config c3 > gpurun_out/bench_${R}_c3_n$N.json 2> gpurun_out/bench_${R}_c3_n$N.err; python tools/show_bench.py gpurun_out/bench_${R}_c3_n$N.json ;;
diag) HALO=peer timeout 300 $TR tools/diag_bands.py 2>&1 | grep "^rank" | sort ;;
esac
done
