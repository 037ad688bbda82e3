#!/bin/bash
# Round 2, GPU call 21: end-of-round ncu full-set refresh of the kernels the bench line quotes; slow-step frequency by sampler.
mkdir -p gpurun_out
NCU_TARGETS=final python tools/ncu_targets.py > gpurun_out/r2c21_plain.log 2>&1 &&
NCU_TARGETS=final timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'conv_|bn_' -f -o /tmp/r2_final python tools/ncu_targets.py > gpurun_out/r2c21_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2c21_ncu.log
ncu -i /tmp/r2_final.ncu-rep --page raw --csv > gpurun_out/r2_final_raw.csv 2> gpurun_out/r2c21_export.log; ls -la /tmp/r2_final.ncu-rep gpurun_out/r2_final_raw.csv
one() { env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$*', 'value', round(d['value'], 1), 'p50', round(d['step_ms']['p50'], 2), 'max', round(d['step_ms']['max'], 1), '| e2e', round(d['e2e']['value'], 1), 'max', round(d['e2e']['step_ms']['max'], 1), d['clocks'].get('samples'))"; }
for i in 1 2 3 4; do one GCDLSS_BENCH_CLOCKS=nvml; done
for i in 1 2 3 4; do one GCDLSS_BENCH_CLOCKS=smi; done
for i in 1 2; do one GCDLSS_BENCH_NO_CLOCKS=1; done
