#!/bin/bash
# A/B on one box: end-to-end numbers with the pinned host buffers allocated on the GPU's NUMA node (default) or wherever the
# process happens to run (FHE_B200_BENCH_NUMA=0)
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -4
for rep in 1 2; do for m in 0 1; do
  FHE_B200_BENCH_NUMA=$m python bench.py --no-cpu-baseline > gpurun_out/numa_$m.json 2>/dev/null
  python -c "
import json;d=json.loads(open('gpurun_out/numa_$m.json').read().strip().splitlines()[-1]);print('numa=$m value',round(d['value']),'e2e',round(d['e2e']['value']),'hmult e2e',round(d['hmult']['e2e']['value'],1), 'hmult', round(d['hmult']['value'],1))"
done; done
