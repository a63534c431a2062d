#!/bin/bash
mkdir -p gpurun_out
python tools/prof_hmult.py 4 > gpurun_out/prof_hmult_plain24.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lincomb_mma -s 11 -c 11 -o gpurun_out/prof_lcmma_r24 python tools/prof_hmult.py 4 > gpurun_out/ncu_lcmma24.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_lcmma_r24.ncu-rep --page raw --csv > gpurun_out/lcmma24_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/lcmma24_raw.csv
