#!/bin/bash
# Round 2, GPU call 20: full suite at HEAD (static kernel restored as its own instantiation, explicit roundings in the BN backward), bench.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/r2c20_$name.log 2>&1; echo "$name rc=$?"; tail -${TAIL:-3} gpurun_out/r2c20_$name.log; }
TAIL=6 run tests 1500 python -m pytest tests -m gpu -q --timeout 600 -rfE
grep -E "^E  |^FAILED" gpurun_out/r2c20_tests.log | head -20 | cut -c1-300
short() { grep '^{' gpurun_out/r2c20_$1.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$1', 'value', round(d['value'], 1), 'mean', round(d['ms_per_step'], 3), d['step_ms'], 'e2e', round(d['e2e']['value'], 1), d['e2e']['step_ms'], 'frac', round(d['roofline']['frac'], 4), {k: round(v['kernel_ms_per_step'], 3) for k, v in d['roofline']['by_kernel'].items()})"; }
b() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $EXTRA > gpurun_out/r2c20_$name.log 2>&1; echo "$name rc=$?"; short $name; }
b default A=1
b default2 A=1
