#!/bin/bash
# Round 2, GPU call 11: where do the slow steps come from (clock sampler A/B with per-step dumps), dynamic schedule v2.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/r2c11_$name.log 2>&1; echo "$name rc=$?"; tail -${TAIL:-3} gpurun_out/r2c11_$name.log; }
TAIL=4 run tests 900 python -m pytest tests/test_gpu_zzzz_sched.py tests/test_gpu_fused_block.py tests/test_gpu_minkunet.py -m gpu -q --timeout 600 -rfE -x
short() { grep '^{' gpurun_out/r2c11_$1.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$1', 'value', round(d['value'], 1), 'mean', round(d['ms_per_step'], 3), d['step_ms'], 'e2e', round(d['e2e']['value'], 1), d['e2e']['step_ms'], d['clocks'])"; grep '^step_ms' gpurun_out/r2c11_$1.log; }
b() { name=$1; shift; env GCDLSS_BENCH_DUMP_STEPS=1 "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $EXTRA > gpurun_out/r2c11_$name.log 2>&1; echo "$name rc=$?"; short $name; }
b smi GCDLSS_BENCH_CLOCKS=smi
b nvml GCDLSS_BENCH_CLOCKS=nvml
b none GCDLSS_BENCH_NO_CLOCKS=1
b smi2 GCDLSS_BENCH_CLOCKS=smi
b nvml2 GCDLSS_BENCH_CLOCKS=nvml
b dyn2 GCD_DYN_TILES=1
b dyn2_asc GCD_DYN_TILES=2
