#!/bin/bash
# Round 2, GPU call 15 (2 GPUs): where the gradient sink goes wrong.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
c() { name=$1; port=$2; shift 2; env "$@" timeout 600 $TR --master-port $port tools/ddp_check.py > gpurun_out/r2c15_$name.log 2>&1; echo "$name rc=$?"; grep "^rank" gpurun_out/r2c15_$name.log | cut -c1-900; }
c default 29531 A=1
c nogate 29532 GCDLSS_DDP_GATE=0
c noside 29533 GCD_WGRAD_SIDE=0
c nopdl 29534 GCD_PDL=0
