# usage (GPU box): ROMIS_AB_TOGGLES="A=1 B=2" bash tools/r02_n3.sh -- frame times (full frame and the 135-row band 472..607) under each toggle, no tests
q() { echo "== $*"; env "$@" timeout 120 python tools/frame_time.py 2>&1 | tail -1; env "$@" timeout 120 python tools/frame_time.py 472 607 2>&1 | tail -1; }
for t in $ROMIS_AB_TOGGLES; do q $t; done
