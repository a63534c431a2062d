#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python bench.py --steps 30 --no-hmult --no-cpu-baseline > gpurun_out/bench22_$tag.json 2> gpurun_out/bench22_$tag.err; python -c "
import json;d=json.load(open('gpurun_out/bench22_$tag.json'));print('$tag value',round(d['value']),'int frac',round(d['int_pipe']['frac'],3),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'])"; }
run m1 FHE_B200_BAL_M=1
run m2 FHE_B200_BAL_M=2
run m4 FHE_B200_BAL_M=4
run m2g2 FHE_B200_BAL_M=2 FHE_B200_BAL_GROUPS=2
run m2g1 FHE_B200_BAL_M=2 FHE_B200_BAL_GROUPS=1
python tools/prof_ntt.py 64 1 > gpurun_out/prof_plain22.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bal_ -c 4 -o gpurun_out/prof_bal_r22 python tools/prof_ntt.py 64 1 > gpurun_out/ncu_full22.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_bal_r22.ncu-rep --page raw --csv > gpurun_out/bal22_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/bal22_raw.csv
