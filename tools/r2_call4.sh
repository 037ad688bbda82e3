#!/bin/bash
# Round 2, GPU call 4: per-launch time list of the bench step (ncu, time only), tile-sort threshold A/B, epilogue changes.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/r2c4_$name.log 2>&1; echo "$name rc=$?"; tail -${TAIL:-3} gpurun_out/r2c4_$name.log; }
run tests 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_zz_tilesort.py tests/test_gpu_fullsize.py tests/test_gpu_fused_block.py -m gpu -q --timeout 600 -rfE -x
run bench_default 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
GCDLSS_TILE_SORT_MIN_ROWS=2048 run bench_sort2k 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
GCDLSS_TILE_SORT_MIN_ROWS=65536 run bench_sort64k 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
TAIL=40 run layers 300 python tools/diag_tc.py
export GCDLSS_BENCH_FIXED_WARMUP=1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c4_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 7000 -c 2400 --csv --log-file gpurun_out/r2c4_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c4_ncu.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r2c4_ncu.log
