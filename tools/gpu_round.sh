#!/bin/bash
# One GPU round trip: diagnostics, tests, smoke, bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 300 python tools/diag_grad.py > gpurun_out/diag.log 2>&1; echo "diag rc=$?"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread -s > gpurun_out/tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench.log
