#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu29.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu29.log
python bench.py --steps 30 --no-hmult --no-cpu-baseline > gpurun_out/bench29.json 2> gpurun_out/bench29.err; python -c "
import json;d=json.load(open('gpurun_out/bench29.json'));print('value',round(d['value']),'e2e',round(d['e2e']['value']),'int frac',round(d['int_pipe']['frac'],3),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'], d['clocks'])"
./tools/gpu_round28.sh
