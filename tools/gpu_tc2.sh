#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/tc2_pytest.log 2>&1; echo "sharded pytest rc=$?"; tail -3 gpurun_out/tc2_pytest.log
python tools/prof_lincomb.py 4 > gpurun_out/tc2_times.json 2> gpurun_out/tc2.err; echo "times rc=$?"; cat gpurun_out/tc2_times.json
PROF=1 python tools/prof_lincomb.py 4 > gpurun_out/tc2_plain.log 2>&1 && PROF=1 ncu --set full --clock-control none --import-source on -k regex:lincomb_tc -s 4 -c 4 -o gpurun_out/tc2_prof -f python tools/prof_lincomb.py 4 > gpurun_out/tc2_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/tc2_ncu.log
