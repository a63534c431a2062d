#!/bin/bash
# A/B on one box: resident CTAs per SM of the tensor-core base conversion (2 = before, unset = occupancy API)
mkdir -p gpurun_out
for rep in 1 2; do for c in 2 0; do
  FHE_B200_LINCOMB_CTAS=$c timeout 300 python bench_hmult.py --batch 8 --steps 5 2>/dev/null > gpurun_out/ab_$c.json
  python -c "
import json;d=json.load(open('gpurun_out/ab_$c.json'));print('ctas=$c',round(d['value'],1),round(d['ms_per_op'],4),{k:v['ms'] for k,v in d['kernel_ms_per_call'].items() if isinstance(v,dict)})"
done; done
