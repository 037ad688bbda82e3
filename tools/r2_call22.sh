#!/bin/bash
# Round 2, GPU call 22: slow steps of the e2e loop: allocator (cudaMalloc inside the timed region?) and expandable segments A/B.
mkdir -p gpurun_out
one() { env GCDLSS_BENCH_DUMP_STEPS=1 "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > /tmp/o.log 2> /tmp/e.log; grep '^{' /tmp/o.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$*', 'value', round(d['value'], 1), 'max', round(d['step_ms']['max'], 1), '| e2e', round(d['e2e']['value'], 1), 'max', round(d['e2e']['step_ms']['max'], 1))"; grep -E "allocator segments|^host_ms e2e" /tmp/e.log | cut -c1-260; }
for i in 1 2 3 4; do one A=1; done
for i in 1 2 3 4; do one PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True; done
