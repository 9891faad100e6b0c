# usage (2-GPU box): bash tools/r02_mgpu5.sh <tag> -- R-MIS / R-OMIS band tests, the bench lines of rmis / romis / c2 on 1 and 2 GPUs
R=$1; N=${2:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests/test_gpu_rmis.py tests/test_gpu_romis.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -3
for cfg in rmis romis; do
  timeout 300 $TR bench.py --gpus $N --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_${cfg}_n$N.json 2> gpurun_out/bench_${R}_${cfg}_n$N.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${R}_${cfg}_n$N.json').read().strip().splitlines()[-1]); print('$cfg N=$N', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['e2e'].get('one_host_image'), d['config'].get('band_edges'), d['roofline']['stages_ms_per_frame'])" || tail -5 gpurun_out/bench_${R}_${cfg}_n$N.err
done
timeout 300 $TR bench.py --gpus $N --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${R}_c2_n$N.json 2> gpurun_out/bench_${R}_c2_n$N.err; python tools/show_bench.py gpurun_out/bench_${R}_c2_n$N.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_${R}_c2_n$N.json').read().strip().splitlines()[-1]); print('one image', d['e2e'].get('one_host_image'), d['config'].get('band_edges'))" || tail -5 gpurun_out/bench_${R}_c2_n$N.err
