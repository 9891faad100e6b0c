for v in tw4 tw16; do echo $v; ROMIS_GPU_LIB=romis_b200/build/lib_$v.so python tools/quick_bench.py 2>&1 | tail -1; done
echo default; python tools/quick_bench.py 2>&1 | tail -1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_rmis.py tests/test_gpu_romis.py -m gpu -x -q 2>&1 | tail -3
