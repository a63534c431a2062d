#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bfv.py tests/test_gpu_sharded.py tests/test_golden.py -x -q -m gpu > gpurun_out/quick_tc_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/quick_tc_pytest.log
for B in 1 8; do timeout 300 python bench_hmult.py --batch $B --steps 10 > gpurun_out/quick_tc_b$B.json 2>> gpurun_out/quick_tc.err; echo "b$B rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/quick_tc_b$B.json').read().strip().splitlines()[-1]);print($B, round(d['value'],1), d['kernel_ms_per_call']['lincomb'], d['decrypts_to_product'])"; done
timeout 300 python bench_hmult.py --limb-sharded --batch 1 --steps 20 | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('sharded world1', round(d['value'],1), d['matches_single_gpu_bit_exact'])"
