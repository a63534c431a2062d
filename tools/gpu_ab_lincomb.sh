#!/bin/bash
# A/B on one box: resident CTAs per SM of the tensor-core base conversion (FHE_B200_LINCOMB_CTAS=1|2 caps them; 0 = occupancy API)
mkdir -p gpurun_out
for b in 8 16; do for c in 0 2 1; do
  FHE_B200_LINCOMB_CTAS=$c timeout 300 python bench_hmult.py --batch $b --steps 5 2>/dev/null > gpurun_out/ab_$c.json
  python -c "
import json;d=json.load(open('gpurun_out/ab_$c.json'));print('batch $b ctas=$c',round(d['value'],1),round(d['ms_per_op'],4))"
done; done
