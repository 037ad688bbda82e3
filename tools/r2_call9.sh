#!/bin/bash
# Round 2, GPU call 9: dynamic tile schedule + execution context (weight gradients on a second stream): tests, then bench A/B.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/r2c9_$name.log 2>&1; echo "$name rc=$?"; tail -${TAIL:-3} gpurun_out/r2c9_$name.log; }
TAIL=15 run tests_new 600 python -m pytest tests/test_gpu_zzzz_sched.py -m gpu -q --timeout 300 -rfE -x
TAIL=8 run tests 1500 python -m pytest tests -m gpu -q --timeout 600 -rfE -x --deselect tests/test_gpu_zzzz_sched.py
short() { tail -1 gpurun_out/r2c9_$1.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$1', 'value', round(d['value'], 1), 'p50', round(d['step_ms']['p50'], 3), 'p10', round(d['step_ms']['p10'], 3), 'e2e', round(d['e2e']['value'], 1), 'e2e_p50', round(d['e2e']['step_ms']['p50'], 3))"; }
for cfg in "0 0" "0 1" "1 0" "1 1" "2 1" "1 1"; do
  set -- $cfg
  GCD_WGRAD_SIDE=$1 GCD_DYN_TILES=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c9_bench_s$1_d$2.log 2>&1; echo "bench side=$1 dyn=$2 rc=$?"; short bench_s$1_d$2
done
