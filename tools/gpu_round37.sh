#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lincomb_mma -s 11 -c 1 -o /tmp/prof_lcmma_r37 python tools/prof_hmult.py 4 > gpurun_out/ncu_lcmma37.log 2>&1
ncu -i /tmp/prof_lcmma_r37.ncu-rep --page source --csv --print-source sass > gpurun_out/lcmma37_src.csv 2>/dev/null
ncu -i /tmp/prof_lcmma_r37.ncu-rep --page raw --csv > gpurun_out/lcmma37_raw.csv 2>/dev/null
ls -la gpurun_out/lcmma37*
