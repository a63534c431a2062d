#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --launch-skip 6 --launch-count 1 -o gpurun_out/prof_mb3 ./tools/microbench3.bin > gpurun_out/ncu_mb3.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_mb3.ncu-rep --page raw --csv > gpurun_out/mb3_raw.csv 2>/dev/null
