#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/fused1_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/fused1_pytest.log
for B in 1 8; do
FHE_B200_FUSED_TILE=0 timeout 300 python bench_hmult.py --batch $B --steps 10 > gpurun_out/fused1_sep_b$B.json 2>> gpurun_out/fused1.err; echo "sep b$B rc=$?"
timeout 300 python bench_hmult.py --batch $B --steps 10 > gpurun_out/fused1_fus_b$B.json 2>> gpurun_out/fused1.err; echo "fused b$B rc=$?"
done
python - <<'PY'
import json
for n in ('sep_b1','fus_b1','sep_b8','fus_b8'):
    try:
        d=json.loads(open(f'gpurun_out/fused1_{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],1), 'ops/s', d['ms_per_op'], d['gpu_launches'], d['kernel_ms_per_call'], 'sq', round(d['square']['value'],1), 'relin', round(d['relinearize']['value'],1), 'ok', d['decrypts_to_product'])
    except Exception as e: print(n, 'ERR', e)
PY
tail -5 gpurun_out/fused1.err
