#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bfv.py -x -q -m gpu -k "base_conversion or scale_and_round" > gpurun_out/tc1_pytest.log 2>&1; echo "conv pytest rc=$?"; tail -15 gpurun_out/tc1_pytest.log
timeout 900 python -m pytest tests/test_gpu_bfv.py tests/test_golden.py tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/tc1_pytest2.log 2>&1; echo "bfv pytest rc=$?"; tail -8 gpurun_out/tc1_pytest2.log
for B in 1 8; do
FHE_B200_LINCOMB_TC=0 timeout 300 python bench_hmult.py --batch $B --steps 10 > gpurun_out/tc1_mma_b$B.json 2>> gpurun_out/tc1.err; echo "mma b$B rc=$?"
timeout 300 python bench_hmult.py --batch $B --steps 10 > gpurun_out/tc1_tc_b$B.json 2>> gpurun_out/tc1.err; echo "tc b$B rc=$?"
done
python - <<'PY'
import json
for n in ('mma_b1','tc_b1','mma_b8','tc_b8'):
    try:
        d=json.loads(open(f'gpurun_out/tc1_{n}.json').read().strip().splitlines()[-1])
        k=d['kernel_ms_per_call']
        print(n, round(d['value'],1), 'ops/s', round(d['ms_per_op'],4), d['gpu_launches'], {a:b['ms'] for a,b in k.items() if isinstance(b,dict)}, k['whole_call_ms'], 'ok', d['decrypts_to_product'], 'enc', round(d['encrypt']['value']), 'dec', round(d['decrypt']['value']))
    except Exception as e: print(n, 'ERR', e)
PY
tail -5 gpurun_out/tc1.err
