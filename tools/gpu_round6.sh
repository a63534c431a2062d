#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu6.log
python bench.py --steps 30 --no-hmult --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err; python -c "
import json;d=json.load(open('gpurun_out/bench6.json'));print('NEAR value',round(d['value']),'int frac',round(d['int_pipe']['frac'],3),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'],d['clocks'])"
FHE_B200_NO_NEAR60=1 python bench.py --steps 30 --no-hmult --no-cpu-baseline > gpurun_out/bench6b.json 2> gpurun_out/bench6b.err; python -c "
import json;d=json.load(open('gpurun_out/bench6b.json'));print('generic value',round(d['value']),'int frac',round(d['int_pipe']['frac'],3),d['roofline']['per_kernel_ms'])"
python bench_hmult.py --batch 4 --steps 5 > gpurun_out/hmult6.json 2>gpurun_out/hmult6.err; python -c "
import json;d=json.load(open('gpurun_out/hmult6.json'));print('hmult',round(d['value'],1),round(d['ms_per_op'],3),d['decrypts_to_product'])"
python bench.py --steps 2 --warmup 3 --no-hmult --no-cpu-baseline > gpurun_out/bench_short.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_bench.csv python bench.py --steps 2 --warmup 3 --no-hmult --no-cpu-baseline > gpurun_out/ncu_launches6.log 2>&1
echo "ncu launches rc=$?"
