#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu3.log
for mb in 32 128 512; do FHE_B200_NTT_CHUNK_MB=$mb python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_chunk$mb.json 2>gpurun_out/bench3.err; python -c "
import json;d=json.load(open('gpurun_out/bench3_chunk$mb.json'));print('chunk',$mb,'value',round(d['value']),'e2e',round(d['e2e']['value']),'step_frac',round(d['roofline']['step_frac'],4),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'])"; done
for b in 1 4; do python bench_hmult.py --batch $b --steps 5 > gpurun_out/hmult_b$b.json 2> gpurun_out/hmult_b$b.err; echo "hmult b=$b rc=$?"; cut -c1-600 gpurun_out/hmult_b$b.json; tail -2 gpurun_out/hmult_b$b.err; done
python tools/prof_ntt.py 64 2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --cache-control none -k regex:ntt_ --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 128 -c 8 --csv --log-file gpurun_out/traffic_r3.csv python tools/prof_ntt.py 64 2 > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
