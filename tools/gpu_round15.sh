#!/bin/bash
mkdir -p gpurun_out
python tools/prof_hmult.py 8 > gpurun_out/prof_hmult_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lincomb -s 11 -c 11 -o gpurun_out/prof_lincomb_r15 python tools/prof_hmult.py 8 > gpurun_out/ncu_lincomb15.log 2>&1
echo "ncu rc=$?"
