#!/bin/bash
# Round 2, GPU call 24: the full GPU suite at HEAD, nothing else.
mkdir -p gpurun_out
timeout 270 python -m pytest tests -m gpu -q --timeout 200 -x > gpurun_out/r2c24_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c24_tests.log
