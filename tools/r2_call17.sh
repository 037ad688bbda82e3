#!/bin/bash
# Round 2, GPU call 17 (2 GPUs): gradient sink after the hook fix: check + bench A/B.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 tools/ddp_check.py > gpurun_out/r2c17_ddp_check.log 2>&1; echo "ddp_check rc=$?"; grep "^rank" gpurun_out/r2c17_ddp_check.log | grep -v "parameters whose" | cut -c1-300
short() { grep '^{' gpurun_out/r2c17_$1.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$1', 'n', d['n_gpus'], 'value', round(d['value'], 1), 'mean', round(d['ms_per_step'], 3), d['step_ms'], 'comm_exposed', d['comm_exposed_ms_per_step'], 'e2e', round(d['e2e']['value'], 1))"; }
port=29560
b() { name=$1; shift; port=$((port + 1)); env "$@" timeout 600 $TR --master-port $port bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c17_$name.log 2>&1; echo "$name rc=$?"; short $name; }
b default A=1
b nogate GCDLSS_DDP_GATE=0
b hooks GCDLSS_DDP_DIRECT=0
b default2 A=1
