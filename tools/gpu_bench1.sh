#!/bin/bash
mkdir -p gpurun_out
python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/b1_ref.json 2> gpurun_out/b1_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/b1_ref.json; tail -3 gpurun_out/b1_ref.err
T0=$(date +%s); python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/b1.json 2> gpurun_out/b1.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 )) s"; tail -5 gpurun_out/b1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','sustained','clocks','roofline','hbm','cpu_baseline','e2e'):
    print(k, d.get(k))
h=d.get('hmult',{})
print('hmult', {k:v for k,v in h.items() if k in ('value','batch','ms_per_op','decrypts_to_product','error','kernel_ms_per_call')})
PY
