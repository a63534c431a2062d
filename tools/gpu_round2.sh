#!/bin/bash
mkdir -p gpurun_out
./tools/microbench2.bin > gpurun_out/microbench2.jsonl 2>&1; echo "microbench2 rc=$?"
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu2.log
for b in 1 4 8; do python bench_hmult.py --batch $b --steps 5 > gpurun_out/hmult_b$b.json 2> gpurun_out/hmult_b$b.err; echo "hmult b=$b rc=$?"; cut -c1-700 gpurun_out/hmult_b$b.json; tail -3 gpurun_out/hmult_b$b.err; done
python tools/prof_ntt.py 64 2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -s 8 -c 16 --csv --log-file gpurun_out/traffic_r1.csv python tools/prof_ntt.py 64 2 > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
