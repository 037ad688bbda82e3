#!/bin/bash
# Round 2, GPU call 6: re-run of the adjusted tests, HBM-side kernel table, full-set ncu capture of the quoted kernels.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/r2c6_$name.log 2>&1; echo "$name rc=$?"; tail -${TAIL:-3} gpurun_out/r2c6_$name.log; }
TAIL=6 run tests 900 python -m pytest tests/test_gpu_fused_block.py tests/test_gpu_coords.py tests/test_gpu_mmdet_path.py tests/test_gpu_dataprep.py -m gpu -q --timeout 600 -rfE
TAIL=40 run maps 300 python tools/bench_maps.py
run bench_default 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
python tools/ncu_targets.py > gpurun_out/r2c6_targets_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_full python tools/ncu_targets.py > gpurun_out/r2c6_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2c6_ncu.log; ls -la gpurun_out/r2_full.ncu-rep
