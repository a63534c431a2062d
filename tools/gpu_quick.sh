#!/bin/bash
# quick correctness + headline check after a change (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_quick.log
python bench.py --steps 30 --no-hmult --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print('value',round(d['value']),'e2e',round(d['e2e']['value']),'int frac',round(d['int_pipe']['frac'],3),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'])"
