#!/bin/bash
mkdir -p gpurun_out
for MB in 64 32 16 8; do
FHE_B200_HOST_CHUNK_MB=$MB python bench.py --steps 5 --warmup 3 --no-hmult --no-cpu-baseline > gpurun_out/e2e_chunk_$MB.json 2>/dev/null
python -c "
import json;d=json.loads(open('gpurun_out/e2e_chunk_$MB.json').read().strip().splitlines()[-1]);print($MB, round(d['e2e']['value']), round(d['e2e']['pcie_GBs_per_direction_per_gpu'],1), d['roundtrip_bit_exact'])"
done
