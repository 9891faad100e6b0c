# usage (GPU box): bash tools/r02_n5.sh <tag> -- the default bench leg (L2 flushed per step) of the product library against a library built from an earlier commit, alternating
R=$1
b() { env "$@" timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), {k:round(v['ms_per_launch'],4) for k,v in d['roofline']['passes'].items()})"; }
b X=0
b ROMIS_GPU_LIB=$PWD/romis_b200/build/lib_old.so
b X=0
b ROMIS_GPU_LIB=$PWD/romis_b200/build/lib_old.so
