#!/bin/bash
# Round 2, GPU call 19: full suite at HEAD, smoke, HBM-side table, the bench on every workload, launch list of one step.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/r2c19_$name.log 2>&1; echo "$name rc=$?"; tail -${TAIL:-3} gpurun_out/r2c19_$name.log; }
TAIL=12 run tests_conv 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_zz_tilesort.py tests/test_gpu_coords.py -m gpu -q --timeout 300 -rfE -x
if ! grep -q " passed" gpurun_out/r2c19_tests_conv.log || grep -q "failed" gpurun_out/r2c19_tests_conv.log; then echo "conv / coords tests failed: stopping"; grep -E "^E " gpurun_out/r2c19_tests_conv.log | head -20; exit 1; fi
TAIL=6 run tests 1500 python -m pytest tests -m gpu -q --timeout 600 -rfE -x
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
TAIL=25 run maps 300 python tools/bench_maps.py
TAIL=30 run cabi 120 tools/cabi_check
short() { grep '^{' gpurun_out/r2c19_$1.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$1', 'value', round(d['value'], 1), 'mean', round(d['ms_per_step'], 3), d['step_ms'], 'e2e', round(d['e2e']['value'], 1), d['e2e']['step_ms'], 'frac', round(d['roofline']['frac'], 4), {k: round(v['kernel_ms_per_step'], 3) for k, v in d['roofline']['by_kernel'].items()})"; grep -E '^host_ms e2e|^step_ms e2e' gpurun_out/r2c19_$1.log | cut -c1-300; }
b() { name=$1; shift; env GCDLSS_BENCH_DUMP_STEPS=1 "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $EXTRA > gpurun_out/r2c19_$name.log 2>&1; echo "$name rc=$?"; short $name; }
b default A=1
b default2 A=1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c19_full.log 2>&1; echo "full (with cpu baseline) rc=$?"; short full
EXTRA="--workload nuscenes_b16" b nusc A=1
EXTRA="--workload stage2 --steps 10 --warmup 3" b stage2 A=1
EXTRA="--workload dense" b dense A=1
export GCDLSS_BENCH_FIXED_WARMUP=1 GCDLSS_BENCH_PROFILE_STEP=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r2c19_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c19_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2c19_ncu.log; wc -l gpurun_out/r2c19_launches.csv
