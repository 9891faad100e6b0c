for v in rm2 rm4; do for c in romis rmis; do ROMIS_GPU_LIB=romis_b200/build/lib_$v.so python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v $c', round(d['ms_per_step'],2), d['roofline']['stages_ms_per_frame']['gather_ms'])"; done; done
