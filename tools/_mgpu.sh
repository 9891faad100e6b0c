N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR tools/check_bands_multi_gpu.py --width 1920 --height 1080 --frames 3 2>&1 | grep -v "^W\|warn" | tail -4
$TR bench.py --gpus $N --steps 60 --warmup 5 > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_n$N.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c2_n$N.json').read().strip().splitlines()[-1]); print('c2 N=$N', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['sharding'][:40], d['config']['band_edges'])"
if [ "$N" = "8" ]; then
$TR bench.py --gpus $N --steps 30 --warmup 5 --config c4k > gpurun_out/bench_c4k_n$N.json 2> gpurun_out/bench_c4k_n$N.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c4k_n$N.json').read().strip().splitlines()[-1]); print('c4k N=$N', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
HALO=peer $TR tools/diag_bands.py 2>&1 | grep "^rank" | sort
fi
