#!/bin/bash
# Round 2, GPU call 14: why test_basic_block fails with the mask re-derived from x; pair-list forms; slow steps vs host run-ahead.
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_gpu_fused_block.py" -m gpu -q --timeout 300 -x -k "basic_block" > gpurun_out/r2c14_bb.log 2>&1; echo "basic_block rc=$?"; grep -E "^E|passed|failed" gpurun_out/r2c14_bb.log | head -12
GCD_BN_MASK_FROM_X=0 timeout 600 python -m pytest tests/test_gpu_fused_block.py -m gpu -q --timeout 300 -x > gpurun_out/r2c14_bb0.log 2>&1; echo "fused_block (mask from y) rc=$?"; tail -1 gpurun_out/r2c14_bb0.log
timeout 600 python -m pytest tests/test_gpu_coords.py -m gpu -q --timeout 300 -x > gpurun_out/r2c14_coords.log 2>&1; echo "coords rc=$?"; tail -3 gpurun_out/r2c14_coords.log
GCD_PAIRS_FUSED=1 timeout 300 python tools/bench_maps.py 2>&1 | grep -i "pairs"; GCD_PAIRS_FUSED=2 timeout 300 python tools/bench_maps.py 2>&1 | grep -i "pairs"
short() { grep '^{' gpurun_out/r2c14_$1.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$1', 'value', round(d['value'], 1), 'mean', round(d['ms_per_step'], 3), d['step_ms'], 'e2e', round(d['e2e']['value'], 1), d['e2e']['step_ms'])"; grep -E '^(step|host)_ms resident' gpurun_out/r2c14_$1.log; }
b() { name=$1; shift; env GCDLSS_BENCH_DUMP_STEPS=1 GCD_BN_MASK_FROM_X=0 "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $EXTRA > gpurun_out/r2c14_$name.log 2>&1; echo "$name rc=$?"; short $name; }
b free1 GCDLSS_BENCH_MAX_AHEAD=0
b ahead2_1 GCDLSS_BENCH_MAX_AHEAD=2
b free2 GCDLSS_BENCH_MAX_AHEAD=0
b ahead2_2 GCDLSS_BENCH_MAX_AHEAD=2
b free3 GCDLSS_BENCH_MAX_AHEAD=0
b ahead1 GCDLSS_BENCH_MAX_AHEAD=1
b ahead3 GCDLSS_BENCH_MAX_AHEAD=3
