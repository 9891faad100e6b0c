# usage (8-GPU box): bash tools/r02_mgpu8.sh <tag> -- C2 on 8 and 4 GPUs, C4 and C3 on 8 (band calibration by measured frame time)
R=$1
show() { python tools/show_bench.py $1; python -c "
import json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('   edges', d['config'].get('band_edges'), d['config'].get('band_calibration'), 'one image', d['e2e'].get('one_host_image'))" || tail -5 ${1%.json}.err; }
run() { N=$1; cfg=$2; steps=$3
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + N)) bench.py --gpus $N --config $cfg --steps $steps --warmup 5 --no-cpu-baseline > gpurun_out/bench_${R}_${cfg}_n$N.json 2> gpurun_out/bench_${R}_${cfg}_n$N.err; show gpurun_out/bench_${R}_${cfg}_n$N.json; }
run 8 c2 100
run 4 c2 100
run 8 c4 64
run 8 c3 40
