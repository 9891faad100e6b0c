# usage (8-GPU box): bash tools/r02_mgpu2.sh <tag> -- old (block-per-tile) vs new (persistent) kernels as row bands on 8 / 2 GPUs, C3 on 8
R=$1
tr() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
b() { N=$1; tag=$2; shift; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${R}_${tag}_n$N.json 2> gpurun_out/bench_${R}_${tag}_n$N.err; echo "== $tag N=$N"; python tools/show_bench.py gpurun_out/bench_${R}_${tag}_n$N.json; }
OLD=ROMIS_GPU_LIB=$PWD/romis_b200/build/lib_old.so
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/check_bands_multi_gpu.py --width 1920 --height 1080 --frames 3 2>&1 | grep -v "^W\|warn" | tail -3
b 8 new X=1
b 8 old $OLD
echo "== diag new"; HALO=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/diag_bands.py 2>&1 | grep "^rank" | sort
echo "== diag old"; env $OLD HALO=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/diag_bands.py 2>&1 | grep "^rank" | sort
b 2 new X=1
b 2 old $OLD
echo "== c3 N=8 (old kernels)"; env $OLD timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 --config c3 --no-cpu-baseline > gpurun_out/bench_${R}_c3_n8.json 2> gpurun_out/bench_${R}_c3_n8.err; python tools/show_bench.py gpurun_out/bench_${R}_c3_n8.json
