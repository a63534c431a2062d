#!/bin/bash
# BFV parity tests + HMult bench at batch 1, 4, 8 (run under gpurun)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_bfv.py -m gpu -q -x > gpurun_out/pytest_gpu_hmult.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_hmult.log
for b in 1 4 8; do timeout 300 python bench_hmult.py --batch $b --steps 5 2>gpurun_out/hmult_chk_b$b.err > gpurun_out/hmult_chk_b$b.json; python -c "
import json;d=json.load(open('gpurun_out/hmult_chk_b$b.json'));print('hmult b$b',round(d['value'],1),round(d['ms_per_op'],3),d['decrypts_to_product'],'e2e',round(d['e2e']['value'],1),{k:v['ms'] for k,v in d['kernel_ms_per_call'].items() if isinstance(v,dict)}, {k:round(d[k]['ms_per_op'],4) for k in ('square','multiply_no_relin','relinearize','encrypt','decrypt') if k in d}, d.get('noise_budget_bits'))"; done
./tests/cpp/test_fhe_compat.bin 2>&1 | tail -3
