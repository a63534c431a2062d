#!/bin/bash
# first GPU visit: microbenchmarks, parity tests, smoke, bench, ncu launch list + one full capture of the tile pass
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
./tools/microbench.bin > gpurun_out/microbench.json 2>&1; echo "microbench rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json | cut -c1-1500
for mb in 8 16 64 128; do FHE_B200_NTT_CHUNK_MB=$mb python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chunk$mb.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_chunk$mb.json'));print('chunk',$mb,'value',d['value'],'step_frac',d['roofline']['step_frac'],d['roofline']['per_kernel_ms'])"; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref.json
python tools/prof_ntt.py 64 2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/prof_ntt.py 64 2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python tools/prof_ntt.py 64 1 > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 4 -c 4 -o gpurun_out/prof_ntt_r1 python tools/prof_ntt.py 64 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
