#!/bin/bash
# Round 2, GPU call 8: programmatic dependent launch (GCD_PDL) A/B on the bench + the GPU suite with it on.
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/r2c8_$name.log 2>&1; echo "$name rc=$?"; tail -${TAIL:-3} gpurun_out/r2c8_$name.log; }
TAIL=8 run tests 1500 python -m pytest tests -m gpu -q --timeout 600 -rfE -x
GCD_PDL=0 run bench_pdl0 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
GCD_PDL=1 run bench_pdl1 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
GCD_PDL=0 run bench_pdl0b 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
GCD_PDL=1 run bench_pdl1b 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
GCD_PDL=1 run bench_stage2 900 python bench.py --steps 10 --warmup 3 --workload stage2 --no-cpu-baseline
