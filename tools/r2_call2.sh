#!/bin/bash
# Round 2, GPU call 2: new library defaults (tile-sorted tables + per-tile masks, run-table search, fused pair lists,
# flat gather), A/B against scan order, role counters of the forward kernel (PROFILE build), shared-memory-port ubench.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/r2c2_$name.log 2>&1; echo "$name rc=$?"; tail -${TAIL:-4} gpurun_out/r2c2_$name.log; }
run tests 1500 python -m pytest tests -m gpu -q --timeout 600 -rfE
run bench_default 600 python bench.py --steps 20 --warmup 5
GCDLSS_TILE_SORT=0 run bench_scan 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run bench_nusc 600 python bench.py --steps 20 --warmup 5 --workload nuscenes_b16 --no-cpu-baseline
TAIL=60 GCDLSS_LIB_PATH=$PWD/generalized-class-discovery-for-lidar-semantic-segmentation_b200/gcdlss_b200/libgcdlss_sm100a_profile.so run layers_profile 300 python tools/diag_tc.py
TAIL=40 run layers 300 python tools/diag_tc.py
TAIL=60 run smem_port 120 tools/ubench/smem_port
TAIL=30 run maps 300 python tools/bench_maps.py
