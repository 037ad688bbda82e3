#!/bin/bash
# Round 2, GPU call 18 (8 GPUs): weak-scaling bench with the gradient sink, and through autograd hooks for comparison.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
short() { grep '^{' gpurun_out/r2c18_$1.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$1', 'n', d['n_gpus'], 'value', round(d['value'], 1), 'mean', round(d['ms_per_step'], 3), d['step_ms'], 'comm_exposed', d['comm_exposed_ms_per_step'], 'e2e', round(d['e2e']['value'], 1), d['e2e']['step_ms'])"; }
port=29570
b() { name=$1; shift; port=$((port + 1)); env "$@" timeout 500 $TR --master-port $port bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c18_$name.log 2>&1; echo "$name rc=$?"; short $name; }
b default A=1
b hooks GCDLSS_DDP_DIRECT=0
