for by in 4 8; do for c in romis rmis; do ROMIS_BLOCK_Y=$by python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BY=$by $c', round(d['ms_per_step'],3), d['roofline']['stages_ms_per_frame'])"; done; done
