# usage (2-GPU box): bash tools/r02_mgpu11.sh <tag> -- does the NVML clock sampler (a thread of rank 0) disturb short frames?  C2 on 2 GPUs per polling interval
R=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
for ms in 2 20 200 2; do
  ROMIS_CLOCK_POLL_MS=$ms timeout 200 $TR bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('poll $ms ms:', round(d['ms_per_step'],4), 'calibrated', round(d['config']['band_calibration']['frame_ms'],4), d['config']['band_edges'], 'e2e', round(d['e2e']['value'],1), 'samples', d['clocks']['samples'])"
done
