#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ntt.py -m gpu -q -x > gpurun_out/pytest_gpu21.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu21.log
run() { tag=$1; shift; env "$@" python bench.py --steps 30 --no-hmult --no-cpu-baseline > gpurun_out/bench21_$tag.json 2> gpurun_out/bench21_$tag.err; python -c "
import json;d=json.load(open('gpurun_out/bench21_$tag.json'));print('$tag value',round(d['value']),'int frac',round(d['int_pipe']['frac'],3),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'])"; }
run default X=1
run m4 FHE_B200_BAL_M=4
run m16 FHE_B200_BAL_M=16
run m2 FHE_B200_BAL_M=2
run g2 FHE_B200_BAL_GROUPS=2
run g8 FHE_B200_BAL_GROUPS=8
run g16 FHE_B200_BAL_GROUPS=16
