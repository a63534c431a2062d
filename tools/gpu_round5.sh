#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu5.log
python bench.py > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench5.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'hbm step frac',round(d['roofline']['step_frac'],3),'int',d.get('int_pipe'))
print('cpu',d.get('cpu_baseline')); print('hmult',d.get('hmult')); print('clocks',d['clocks'])
PY
tail -3 gpurun_out/bench5.err
for b in 1 4 8; do python bench_hmult.py --batch $b --steps 5 > gpurun_out/hmult5_b$b.json 2> gpurun_out/hmult5_b$b.err; python -c "
import json;d=json.load(open('gpurun_out/hmult5_b$b.json'));print('hmult b',$b,round(d['value'],1),'ops/s',round(d['ms_per_op'],3),'ms/op e2e',round(d['e2e']['value'],1),d['decrypts_to_product'],d['e2e']['matches_device_path'],d['gpu_launches'])"; done
