#!/bin/bash
# profiles for the judged state (run under gpurun): launch lists of bench.py and of one HMult, full ncu captures of the NTT passes
# and of the two base-conversion kernels; exported to csv (the .ncu-rep files stay on the box: gpurun_out/ is copied back below 64 MiB).
# Summaries: tools/ncu_summary.py, tools/launch_summary.py, tools/lincomb_profile_summary.py -> profiles/r01_*.md
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-hmult --no-cpu-baseline > gpurun_out/b28.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01b_bench.csv python bench.py --steps 2 --warmup 3 --no-hmult --no-cpu-baseline > gpurun_out/ncu_b28.log 2>&1
echo "bench launches rc=$?"
python tools/prof_hmult.py 4 > gpurun_out/h28.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "hmult/" --csv --log-file gpurun_out/launches_r01b_hmult.csv python tools/prof_hmult.py 4 > gpurun_out/ncu_h28.log 2>&1
echo "hmult launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:bal_ -c 4 -o /tmp/prof_bal_r28 python tools/prof_ntt.py 64 1 > gpurun_out/ncu_full28.log 2>&1
ncu -i /tmp/prof_bal_r28.ncu-rep --page raw --csv > gpurun_out/bal28_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:lincomb_mma -s 11 -c 11 -o /tmp/prof_lcmma_r28 python tools/prof_hmult.py 4 > gpurun_out/ncu_lcmma28.log 2>&1
ncu -i /tmp/prof_lcmma_r28.ncu-rep --page raw --csv > gpurun_out/lcmma28_raw.csv 2>/dev/null
FHE_B200_LINCOMB_MMA=0 ncu --set full --clock-control none -k regex:lincomb_kernel -s 11 -c 11 -o /tmp/prof_lcimad_r28 python tools/prof_hmult.py 4 > gpurun_out/ncu_lcimad28.log 2>&1
ncu -i /tmp/prof_lcimad_r28.ncu-rep --page raw --csv > gpurun_out/lcimad28_raw.csv 2>/dev/null
du -sh gpurun_out
