#!/bin/bash
mkdir -p gpurun_out
PROF=1 python tools/prof_lincomb.py 4 > gpurun_out/tc4_plain.log 2>&1 && PROF=1 ncu --set full --clock-control none --import-source on -k regex:lincomb_tc -s 0 -c 4 -o gpurun_out/tc4_prof -f python tools/prof_lincomb.py 4 > gpurun_out/tc4_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/tc4_ncu.log
