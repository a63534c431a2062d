#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu4.log
for mb in 128 1024; do FHE_B200_NTT_CHUNK_MB=$mb python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench4_chunk$mb.json 2>gpurun_out/bench4.err; python -c "
import json;d=json.load(open('gpurun_out/bench4_chunk$mb.json'));print('chunk',$mb,'value',round(d['value']),'e2e',round(d['e2e']['value']),'step_frac',round(d['roofline']['step_frac'],4),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'])"; done
python tools/prof_ntt.py 64 1 > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ntt_tile_fwd -c 1 -o gpurun_out/prof_tile_r4 python tools/prof_ntt.py 64 1 > gpurun_out/ncu_full4.log 2>&1
echo "ncu full rc=$?"
