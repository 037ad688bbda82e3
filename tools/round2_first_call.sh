#!/bin/bash
# First GPU call of round 2: everything written after round 1's GPU budget ran out, each step under its own timeout.
# (tools/cabi_check already ran the new C entry points on a B200 in the last seconds of round 1: profiles/r1_cabi_check_*.txt;
#  what has not run yet is the Python layer on top of them and the full parity suite under the flags.)
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Results land in gpurun_out/r2_*.  Opt-in variants under test:
#   GCDLSS_KMAP=runs    kernel-map search over x-runs (csrc/runtable.cuh)          -> default if parity + faster
#   GCD_PAIRS_FUSED=1   pair lists straight from the table, 2 passes instead of 8 (csrc/scan.cu) -> default if parity + faster
#   GCD_GATHER_FLAT=1   devoxelise gather: float4 elements dealt flat, 4 chains per thread (csrc/gather_rows.cuh)
#   GCDLSS_TILE_SORT=1  forward/dgrad conv on tile-sorted 3x3x3 tables (csrc/tilesort.cuh): 2.2-2.6x fewer stages -> default if parity + faster
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/r2_$name.log 2>&1; echo "$name rc=$?"; tail -3 gpurun_out/r2_$name.log; }
( bash tools/build_cabi_check.sh && tools/cabi_check ) > gpurun_out/r2_cabi_check.log 2>&1; echo "cabi_check rc=$?"
run tests          900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread
run smoke          300 python -c "import __graft_entry__ as g; g.smoke()"
run maps           300 python tools/bench_maps.py
run bench_default  900 python bench.py --steps 20 --warmup 5      # tile sort decided by the self-check (its verdict is in config.tile_sort)
GCDLSS_TILE_SORT=0 run bench_scan_order 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run selfcheck      300 python tools/selfcheck_tilesort.py 0
GCDLSS_KMAP=runs   run tests_runs    600 python -m pytest tests/test_gpu_coords.py tests/test_gpu_minkunet.py -m gpu -q --timeout 300 --timeout-method thread
GCDLSS_TILE_SORT=0 GCDLSS_KMAP=runs   run bench_runs    600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
GCD_PAIRS_FUSED=1  run tests_pairs   600 python -m pytest tests/test_gpu_coords.py tests/test_gpu_conv.py -m gpu -q --timeout 300 --timeout-method thread
GCD_GATHER_FLAT=1  run tests_gather  600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_mmdet_path.py tests/test_gpu_stage2.py -m gpu -q --timeout 300 --timeout-method thread
GCD_PAIRS_FUSED=1 GCD_GATHER_FLAT=1 run maps_optin 300 python tools/bench_maps.py
GCDLSS_TILE_SORT=1 GCDLSS_TILE_SORT_MIN_ROWS=1 run tests_tilesort 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_fused_block.py tests/test_gpu_minkunet.py tests/test_gpu_stage2.py -m gpu -q --timeout 300 --timeout-method thread
GCDLSS_TILE_SORT=1 run bench_tilesort 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run layers_default 300 python tools/diag_tc.py
( cd tools/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../generalized-class-discovery-for-lidar-semantic-segmentation_b200/csrc -o smem_port smem_port.cu ) > /dev/null 2>&1
run smem_port      120 tools/ubench/smem_port
# fresh launch list of the default path (fixed warm-up so --launch-skip lands inside the timed region)
GCDLSS_TILE_SORT=0 GCDLSS_BENCH_FIXED_WARMUP=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 1800 --csv \
  --log-file gpurun_out/r2_launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu.log 2>&1; echo "ncu rc=$?"
