# usage (N-GPU box): bash tools/r02_mgpu7.sh <tag> <N> [configs...] -- bench lines on N GPUs (band calibration by measured frame time), twice for c2
R=$1; N=${2:-2}; shift; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
show() { python tools/show_bench.py $1; python -c "
import json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('   edges', d['config'].get('band_edges'), d['config'].get('band_calibration'), 'one image', d['e2e'].get('one_host_image'))" || tail -5 ${1%.json}.err; }
for cfg in ${@:-c2 c2}; do
  i=$((i+1))
  timeout 300 $TR bench.py --gpus $N --config $cfg --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${R}_${cfg}_n${N}_$i.json 2> gpurun_out/bench_${R}_${cfg}_n${N}_$i.err; show gpurun_out/bench_${R}_${cfg}_n${N}_$i.json
done
