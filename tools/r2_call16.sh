#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tools/ddp_check.py > gpurun_out/r2c16_default.log 2>&1; echo "rc=$?"; grep -v "^  File\|^    raise\|Traceback" gpurun_out/r2c16_default.log | cut -c1-400 | head -60
