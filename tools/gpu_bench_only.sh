#!/bin/bash
# headline bench only (experiments on the NTT passes): value, per-kernel times
mkdir -p gpurun_out
python bench.py --steps 30 --no-hmult --no-cpu-baseline > gpurun_out/bench_only.json 2> gpurun_out/bench_only.err; python -c "
import json;d=json.load(open('gpurun_out/bench_only.json'));print('value',round(d['value']),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'], 'loop', round(d['roofline']['register_only_butterfly_loop']['Gbutterflies_s']))"
