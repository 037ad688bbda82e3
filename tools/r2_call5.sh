#!/bin/bash
# Round 2, GPU call 5: full GPU suite with the fused batch norms / heads / backbone / dataprep, bench A/B of the BN fusion.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/r2c5_$name.log 2>&1; echo "$name rc=$?"; tail -${TAIL:-3} gpurun_out/r2c5_$name.log; }
TAIL=8 run tests 1500 python -m pytest tests -m gpu -q --timeout 600 -rfE
run bench_default 600 python bench.py --steps 20 --warmup 5
GCD_BN_FUSED=0 run bench_bn2pass 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run bench_default2 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run bench_nusc 600 python bench.py --steps 20 --warmup 5 --workload nuscenes_b16 --no-cpu-baseline
run bench_stage2 900 python bench.py --steps 10 --warmup 3 --workload stage2 --no-cpu-baseline
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
