# usage (8-GPU box): bash tools/r02_mgpu4.sh <tag> -- round-2 multi-GPU tables: C2 at 8 / 4 GPUs, C3 at 1 / 4 / 8, band parity of C3, C5 sweep on 8
R=$1
TRN() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
b() { cfg=$1; N=$2; if [ $N = 1 ]; then timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 --config $cfg --no-cpu-baseline > gpurun_out/bench_${R}_${cfg}_n$N.json 2> gpurun_out/bench_${R}_${cfg}_n$N.err;
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 --config $cfg --no-cpu-baseline > gpurun_out/bench_${R}_${cfg}_n$N.json 2> gpurun_out/bench_${R}_${cfg}_n$N.err; fi
  echo "== $cfg N=$N"; python tools/show_bench.py gpurun_out/bench_${R}_${cfg}_n$N.json; }
b c2 8
b c2 4
echo "== diag c2 N=8"; HALO=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/diag_bands.py 2>&1 | grep "^rank" | sort
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/check_bands_multi_gpu.py --config c3 --width 1920 --height 1080 --frames 2 2>&1 | grep -v "^W\|warn" | tail -2
b c3 8
b c3 4
b c3 1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/sweep_c5.py > gpurun_out/c5_sweep_${R}_n8.md 2> gpurun_out/c5_sweep_${R}_n8.err; tail -8 gpurun_out/c5_sweep_${R}_n8.md; tail -2 gpurun_out/c5_sweep_${R}_n8.err
