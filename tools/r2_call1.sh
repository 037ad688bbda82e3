#!/bin/bash
# Round 2, GPU call 1: the whole GPU suite under the library defaults and again with every opt-in path switched on,
# smoke(), and A/B bench lines (scan-order vs tile-sorted tables, point-wise vs run-table kernel-map search).
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/r2c1_$name.log 2>&1; echo "$name rc=$?"; tail -4 gpurun_out/r2c1_$name.log; }
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader; nproc
run tests_default 1500 python -m pytest tests -m gpu -q --timeout 600 -rfE --durations=8
GCDLSS_TILE_SORT=1 GCDLSS_KMAP=runs run tests_optin 1500 python -m pytest tests -m gpu -q --timeout 600 -rfE
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
GCDLSS_TILE_SORT=0 run bench_scan 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
GCDLSS_TILE_SORT=1 run bench_sort 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
GCDLSS_TILE_SORT=1 GCDLSS_KMAP=runs run bench_sort_runs 600 python bench.py --steps 20 --warmup 5
GCDLSS_TILE_SORT=1 GCDLSS_KMAP=runs run bench_nusc 600 python bench.py --steps 20 --warmup 5 --workload nuscenes_b16 --no-cpu-baseline
GCDLSS_TILE_SORT=0 run layers_scan 300 python tools/diag_tc.py
run maps 300 python tools/bench_maps.py
run reference 600 python bench.py --impl reference --steps 3 --warmup 1
