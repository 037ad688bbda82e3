#!/bin/bash
# Round 2, GPU call 13: cost of clock / throttle queries next to a launch stream; host time per step; pipelined dynamic schedule.
mkdir -p gpurun_out
timeout 300 python tools/nvml_probe.py > gpurun_out/r2c13_nvml_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2c13_nvml_probe.log
GCDLSS_BENCH_NO_CLOCKS=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --debug-steps > gpurun_out/r2c13_debug.log 2>&1; echo "debug rc=$?"; grep "debug step" gpurun_out/r2c13_debug.log | tail -8
short() { grep '^{' gpurun_out/r2c13_$1.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$1', 'value', round(d['value'], 1), 'mean', round(d['ms_per_step'], 3), d['step_ms'], 'e2e', round(d['e2e']['value'], 1), d['e2e']['step_ms'])"; }
b() { name=$1; shift; env GCDLSS_BENCH_NO_CLOCKS=1 "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $EXTRA > gpurun_out/r2c13_$name.log 2>&1; echo "$name rc=$?"; short $name; }
timeout 900 python -m pytest tests/test_gpu_zzzz_sched.py tests/test_gpu_fused_block.py tests/test_gpu_minkunet.py tests/test_gpu_conv.py -m gpu -q --timeout 300 -x 2>&1 | tail -3
b static A=1
b dyn3 GCD_DYN_TILES=1
b dyn3_a8 GCD_DYN_TILES=1 GCD_DYN_AHEAD=8
b masky GCD_BN_MASK_FROM_X=0
b static2 A=1
