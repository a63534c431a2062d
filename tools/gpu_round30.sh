#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ntt.py -m gpu -q -x > gpurun_out/pytest_gpu30.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu30.log
run() { tag=$1; shift; env "$@" python bench.py --steps 30 --no-hmult --no-cpu-baseline > gpurun_out/bench30_$tag.json 2> gpurun_out/bench30_$tag.err; python -c "
import json;d=json.load(open('gpurun_out/bench30_$tag.json'));print('$tag value',round(d['value']),'int frac',round(d['int_pipe']['frac'],3),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'])"; }
run default X=1
run g2 FHE_B200_BAL_GROUPS=2
run g4 FHE_B200_BAL_GROUPS=4
run g6 FHE_B200_BAL_GROUPS=6
run default2 X=1
