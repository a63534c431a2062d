#!/bin/bash
# round-2 profiles (run under gpurun, one GPU): launch lists of bench.py and of one HMult, full ncu capture of the tcgen05 conversion kernel
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-hmult --no-cpu-baseline > gpurun_out/p2_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02b_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-hmult --no-cpu-baseline > gpurun_out/p2_bench_ncu.log 2>&1
echo "bench launches rc=$?"
FHE_B200_HMULT_STREAMS=1 python tools/prof_hmult.py 4 > gpurun_out/p2_hmult_plain.log 2>&1 && \
FHE_B200_HMULT_STREAMS=1 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "hmult/" --csv --log-file gpurun_out/r02b_hmult_launches.csv python tools/prof_hmult.py 4 > gpurun_out/p2_hmult_ncu.log 2>&1
echo "hmult launches rc=$?"
FHE_B200_HMULT_STREAMS=1 ncu --set full --clock-control none --import-source on -k regex:lincomb_tc --nvtx --nvtx-include "hmult/" -c 14 -o gpurun_out/r02b_lincomb_tc_hmult -f python tools/prof_hmult.py 4 > gpurun_out/p2_tc_ncu.log 2>&1
echo "tc full rc=$?"
ls -la gpurun_out/r02b_*
