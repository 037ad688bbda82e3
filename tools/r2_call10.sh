#!/bin/bash
# Round 2, GPU call 10: batch-norm reverse walk + the full suite, cabi_check under compute-sanitizer, dynamic-schedule variants.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/r2c10_$name.log 2>&1; echo "$name rc=$?"; tail -${TAIL:-3} gpurun_out/r2c10_$name.log; }
TAIL=8 run tests 1500 python -m pytest tests -m gpu -q --timeout 600 -rfE -x
TAIL=25 run cabi 120 tools/cabi_check
TAIL=6 run memcheck 420 compute-sanitizer --tool memcheck --error-exitcode 7 tools/cabi_check
TAIL=6 run racecheck 420 compute-sanitizer --tool racecheck --error-exitcode 7 tools/cabi_check
short() { tail -1 gpurun_out/r2c10_$1.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$1', 'value', round(d['value'], 1), 'p50', round(d['step_ms']['p50'], 3), 'p10', round(d['step_ms']['p10'], 3), 'e2e', round(d['e2e']['value'], 1), 'e2e_p50', round(d['e2e']['step_ms']['p50'], 3))"; }
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $EXTRA > gpurun_out/r2c10_$name.log 2>&1; echo "$name rc=$?"; short $name; }
b default A=1
b dyn_desc GCD_DYN_TILES=1
b dyn_asc GCD_DYN_TILES=2
b dyn_asc_a8 GCD_DYN_TILES=2 GCD_DYN_AHEAD=8
b dyn_desc_a8 GCD_DYN_TILES=1 GCD_DYN_AHEAD=8
b default2 A=1
EXTRA="--workload nuscenes_b16" b nusc A=1
EXTRA="--workload stage2 --steps 10 --warmup 3" b stage2 A=1
