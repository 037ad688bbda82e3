#!/bin/bash
# Round 2, GPU call 3: whole-trunk fast path (gcd_run_ops), asynchronous table ring, 16 gather warps A/B.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/r2c3_$name.log 2>&1; echo "$name rc=$?"; tail -${TAIL:-4} gpurun_out/r2c3_$name.log; }
run tests 1500 python -m pytest tests -m gpu -q --timeout 600 -rfE -x
TAIL=12 run bench_default 600 python bench.py --steps 20 --warmup 5 --debug-steps
GCD_TC_WARPS=16 run bench_w16 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
GCDLSS_FUSED_BLOCKS=0 run bench_perop 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
TAIL=70 run layers 300 python tools/diag_tc.py
TAIL=70 GCDLSS_LIB_PATH=$PWD/generalized-class-discovery-for-lidar-semantic-segmentation_b200/gcdlss_b200/libgcdlss_sm100a_profile.so run layers_profile 300 python tools/diag_tc.py
GCD_TC_WARPS=16 run tests_w16 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_zz_tilesort.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -rfE
run bench_stage2 900 python bench.py --steps 10 --warmup 3 --workload stage2 --no-cpu-baseline
run bench_dense 900 python bench.py --steps 10 --warmup 3 --workload dense --no-cpu-baseline
