#!/bin/bash
mkdir -p gpurun_out
for b in 16 32; do timeout 300 python bench_hmult.py --batch $b --steps 3 2>gpurun_out/hmult36_b$b.err > gpurun_out/hmult36_b$b.json; python -c "
import json;d=json.load(open('gpurun_out/hmult36_b$b.json'));print('hmult b$b',round(d['value'],1),round(d['ms_per_op'],3),d['decrypts_to_product'],'e2e',round(d['e2e']['value'],1),{k:v['ms'] for k,v in d['kernel_ms_per_call'].items() if isinstance(v,dict)})"; done
python -c "import __graft_entry__ as g; g.smoke()"
