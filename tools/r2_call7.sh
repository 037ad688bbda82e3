#!/bin/bash
# Round 2, GPU call 7: full-set ncu captures, exported to CSV on the box (the reports themselves are too large to bring back).
mkdir -p gpurun_out
rm -f gpurun_out/r2_full.ncu-rep
for part in conv side; do
  NCU_TARGETS=$part python tools/ncu_targets.py > gpurun_out/r2c7_plain_$part.log 2>&1 &&
  NCU_TARGETS=$part timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/r2_$part python tools/ncu_targets.py > gpurun_out/r2c7_ncu_$part.log 2>&1
  echo "ncu $part rc=$?"; tail -2 gpurun_out/r2c7_ncu_$part.log
  ncu -i /tmp/r2_$part.ncu-rep --page raw --csv > gpurun_out/r2_full_${part}_raw.csv 2> gpurun_out/r2c7_export_$part.log
  ls -la /tmp/r2_$part.ncu-rep gpurun_out/r2_full_${part}_raw.csv
done
# source-level view of the forward kernel's first captured launch (stall reasons per line), conv report only
ncu -i /tmp/r2_conv.ncu-rep --page source --csv --kernel-name regex:conv_fwd_tc_kernel --launch-count 1 > gpurun_out/r2_conv_fwd_source.csv 2>> gpurun_out/r2c7_export_conv.log
ls -la gpurun_out/r2_conv_fwd_source.csv
du -sh gpurun_out
