#!/bin/bash
mkdir -p gpurun_out
python tools/prof_ntt.py 64 1 > gpurun_out/prof_plain19.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bal_ -c 4 -o gpurun_out/prof_bal_r19 python tools/prof_ntt.py 64 1 > gpurun_out/ncu_full19.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_bal_r19.ncu-rep --page raw --csv > gpurun_out/bal19_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/bal19_raw.csv
