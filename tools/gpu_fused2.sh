#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bfv.py tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/fused2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/fused2_pytest.log
for B in 1 8; do
FHE_B200_FUSED_TILE=0 timeout 300 python bench_hmult.py --batch $B --steps 10 > gpurun_out/fused2_sep_b$B.json 2>> gpurun_out/fused2.err; echo "sep b$B rc=$?"
timeout 300 python bench_hmult.py --batch $B --steps 10 > gpurun_out/fused2_fus_b$B.json 2>> gpurun_out/fused2.err; echo "fused b$B rc=$?"
FHE_B200_HMULT_STREAMS=1 timeout 300 python bench_hmult.py --batch $B --steps 10 > gpurun_out/fused2_fus1s_b$B.json 2>> gpurun_out/fused2.err; echo "fused 1 stream b$B rc=$?"
done
python - <<'PY'
import json
for n in ('sep_b1','fus_b1','fus1s_b1','sep_b8','fus_b8','fus1s_b8'):
    try:
        d=json.loads(open(f'gpurun_out/fused2_{n}.json').read().strip().splitlines()[-1])
        k=d['kernel_ms_per_call']
        print(n, round(d['value'],1), 'ops/s', round(d['ms_per_op'],4), d['gpu_launches'], {a:b['ms'] for a,b in k.items() if isinstance(b,dict)}, k['whole_call_ms'], 'sq', round(d['square']['value'],1), 'relin', round(d['relinearize']['value'],1), 'ok', d['decrypts_to_product'])
    except Exception as e: print(n, 'ERR', e)
PY
tail -5 gpurun_out/fused2.err
