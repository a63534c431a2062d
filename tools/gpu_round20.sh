#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu20.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu20.log
python bench.py --steps 30 --no-hmult --no-cpu-baseline > gpurun_out/bench20.json 2> gpurun_out/bench20.err; python -c "
import json;d=json.load(open('gpurun_out/bench20.json'));print('value',round(d['value']),'e2e',round(d['e2e']['value']),'int frac',round(d['int_pipe']['frac'],3),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'])"
FHE_B200_NTT_BAL=0 python bench.py --steps 30 --no-hmult --no-cpu-baseline > gpurun_out/bench20_old.json 2> gpurun_out/bench20_old.err; python -c "
import json;d=json.load(open('gpurun_out/bench20_old.json'));print('OLD value',round(d['value']),'e2e',round(d['e2e']['value']),'int frac',round(d['int_pipe']['frac'],3),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'])"
for b in 1 4; do timeout 300 python bench_hmult.py --batch $b --steps 5 2>gpurun_out/hmult20_b$b.err > gpurun_out/hmult20_b$b.json; python -c "
import json;d=json.load(open('gpurun_out/hmult20_b$b.json'));print('hmult b$b',round(d['value'],1),round(d['ms_per_op'],3),d['decrypts_to_product'],d['kernel_ms_per_call'])"; done
