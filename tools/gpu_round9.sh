#!/bin/bash
mkdir -p gpurun_out
python tools/prof_hmult.py 1 > gpurun_out/prof_hmult_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_hmult.csv python tools/prof_hmult.py 1 > gpurun_out/ncu_hmult.log 2>&1
echo "ncu rc=$?"
