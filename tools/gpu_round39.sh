#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu39.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu39.log
for b in 1 16 64; do timeout 300 python bench_hmult.py --preset c2 --batch $b --steps 10 2>gpurun_out/hmult39_c2_b$b.err > gpurun_out/hmult39_c2_b$b.json; python -c "
import json;d=json.load(open('gpurun_out/hmult39_c2_b$b.json'));print('c2 b$b hmult',round(d['value'],1),'us/op',round(1e3*d['ms_per_op'],1),d['decrypts_to_product'],'enc',round(d['encrypt']['value']),'dec',round(d['decrypt']['value']),'e2e',round(d['e2e']['value'],1))"; done
timeout 300 python bench_hmult.py --batch 32 --steps 3 2>/dev/null | python -c "
import json,sys;d=json.load(sys.stdin);print('c4 b32 hmult',round(d['value'],1),'enc',round(d['encrypt']['value']),'dec',round(d['decrypt']['value']))"
