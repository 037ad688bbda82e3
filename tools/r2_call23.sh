#!/bin/bash
# Round 2, GPU call 23 (4 GPUs): the missing point of the weak-scaling table.
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c23_n4.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/r2c23_n4.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('n', d['n_gpus'], 'value', round(d['value'], 1), 'mean', round(d['ms_per_step'], 3), d['step_ms'], 'comm_exposed', d['comm_exposed_ms_per_step'], 'e2e', round(d['e2e']['value'], 1), d['e2e']['step_ms'])"
