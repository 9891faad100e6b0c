# usage (N-GPU box): bash tools/r02_mgpu6.sh <tag> <N> -- multi-GPU tests, then C4 (orbiting camera, bands cut for the path's mean profile) on 1 and N GPUs, C2 on N
R=$1; N=${2:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
show() { python tools/show_bench.py $1; python -c "
import json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('   edges', d['config'].get('band_edges'), 'one image', d['e2e'].get('one_host_image'))" || tail -5 ${1%.json}.err; }
timeout 300 python bench.py --config c4 --steps 64 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${R}_c4_n1.json 2> gpurun_out/bench_${R}_c4_n1.err; show gpurun_out/bench_${R}_c4_n1.json
timeout 300 $TR bench.py --gpus $N --config c4 --steps 64 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${R}_c4_n$N.json 2> gpurun_out/bench_${R}_c4_n$N.err; show gpurun_out/bench_${R}_c4_n$N.json
timeout 300 $TR bench.py --gpus $N --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${R}_c2_n$N.json 2> gpurun_out/bench_${R}_c2_n$N.err; show gpurun_out/bench_${R}_c2_n$N.json
