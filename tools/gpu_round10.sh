#!/bin/bash
mkdir -p gpurun_out
python tools/prof_hmult.py 1 > gpurun_out/prof_hmult_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lincomb -s 21 -c 12 -o gpurun_out/prof_lincomb_r01 python tools/prof_hmult.py 1 > gpurun_out/ncu_lincomb.log 2>&1
echo "ncu rc=$?"
