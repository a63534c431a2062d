#!/bin/bash
# NTT + BFV parity tests, headline bench and HMult bench (run under gpurun)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_bfv.py -m gpu -q -x > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_full.log
python bench.py --steps 30 --no-hmult --no-cpu-baseline > gpurun_out/bench_chk.json 2> gpurun_out/bench_chk.err; python -c "
import json;d=json.load(open('gpurun_out/bench_chk.json'));print('value',round(d['value']),'e2e',round(d['e2e']['value']),'int frac',round(d['int_pipe']['frac'],3),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'])"
for b in 1 4 8; do timeout 300 python bench_hmult.py --batch $b --steps 5 2>gpurun_out/hmult_full_b$b.err > gpurun_out/hmult_full_b$b.json; python -c "
import json;d=json.load(open('gpurun_out/hmult_full_b$b.json'));print('hmult b$b',round(d['value'],1),round(d['ms_per_op'],3),d['decrypts_to_product'],'e2e',round(d['e2e']['value'],1),{k:v['ms'] for k,v in d['kernel_ms_per_call'].items() if isinstance(v,dict)})"; done
