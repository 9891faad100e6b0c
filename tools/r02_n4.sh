# usage (GPU box): bash tools/r02_n4.sh -- one-directional fences of the row-group links: frame times (full, 135-row band) per library and toggle
q() { echo "== $*"; env "$@" timeout 120 python tools/frame_time.py 2>&1 | tail -1; env "$@" timeout 120 python tools/frame_time.py 472 607 2>&1 | tail -1; }
B=$PWD/romis_b200/build
q X=0
q ROMIS_FINE=0
q ROMIS_FINE=1
q ROMIS_FINE_INITIAL=1
q ROMIS_GPU_LIB=$B/lib_sc.so
q ROMIS_GPU_LIB=$B/lib_sc.so ROMIS_FINE=1
q ROMIS_GPU_LIB=$B/lib_strong.so ROMIS_FINE_INITIAL=1
