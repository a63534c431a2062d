#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ntt.py -m gpu -q -x > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest ntt rc=$?"; tail -4 gpurun_out/pytest_gpu11.log
timeout 600 python -m pytest tests/test_gpu_bfv.py tests/test_gpu_compat.py -m gpu -q -x > gpurun_out/pytest_gpu11b.log 2>&1; echo "pytest bfv rc=$?"; tail -3 gpurun_out/pytest_gpu11b.log
timeout 300 python bench.py --steps 30 --no-hmult --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err; python -c "
import json;d=json.load(open('gpurun_out/bench11.json'));print('FUSED value',round(d['value']),'e2e',round(d['e2e']['value']),'int frac',round(d['int_pipe']['frac'],3),'hbm',round(d['roofline']['step_frac'],3),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'])"; tail -2 gpurun_out/bench11.err
FHE_B200_NTT_FUSED=0 timeout 300 python bench.py --steps 30 --no-hmult --no-cpu-baseline > gpurun_out/bench11b.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench11b.json'));print('2PASS value',round(d['value']),d['roofline']['per_kernel_ms'])"
timeout 300 python bench_hmult.py --batch 4 --steps 5 > gpurun_out/hmult11.json 2>gpurun_out/hmult11.err; python -c "
import json;d=json.load(open('gpurun_out/hmult11.json'));print('hmult',round(d['value'],1),round(d['ms_per_op'],3),d['decrypts_to_product'],'e2e',round(d['e2e']['value'],1))"
