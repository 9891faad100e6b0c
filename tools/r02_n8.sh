# usage (GPU box): bash tools/r02_n8.sh -- row-group counters on / off for bands of 270 and 540 rows (the 4- and 2-GPU band heights at 1080p)
for band in "405 675" "270 810"; do for t in ROMIS_FINE=0 ROMIS_FINE=1; do echo "== $t band $band"; env $t timeout 60 python tools/frame_time.py $band 2>&1 | tail -1; done; done
