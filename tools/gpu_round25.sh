#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_bfv.py -m gpu -q -x > gpurun_out/pytest_gpu25.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu25.log
for mma in 1; do for b in 1 4 8; do FHE_B200_LINCOMB_MMA=$mma timeout 300 python bench_hmult.py --batch $b --steps 5 2>gpurun_out/hmult25_m${mma}_b$b.err > gpurun_out/hmult25_m${mma}_b$b.json; python -c "
import json;d=json.load(open('gpurun_out/hmult25_m${mma}_b$b.json'));print('mma=$mma hmult b$b',round(d['value'],1),round(d['ms_per_op'],3),d['decrypts_to_product'],'e2e',round(d['e2e']['value'],1),{k:v['ms'] for k,v in d['kernel_ms_per_call'].items() if isinstance(v,dict)})"; done; done
python tools/prof_hmult.py 4 > gpurun_out/prof_hmult_plain25.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lincomb_mma -s 11 -c 11 -o gpurun_out/prof_lcmma_r25 python tools/prof_hmult.py 4 > gpurun_out/ncu_lcmma25.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_lcmma_r25.ncu-rep --page raw --csv > gpurun_out/lcmma25_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/lcmma25_raw.csv | grep -v "^  launch__\|l1tex__data_pipe\|dram__bytes\|lts__t_sector"
