#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bfv.py tests/test_golden.py tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/tc3_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/tc3_pytest.log
python tools/prof_lincomb.py 4 > gpurun_out/tc3_times.json 2> gpurun_out/tc3.err; echo "times rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/tc3_times.json').read().rsplit('prof_lincomb ok',1)[0])
for k,v in d.items(): print(k, v['tc'], v['mma'], 'same' if v['sha_tc']==v['sha_mma'] else 'DIFF')
"
for B in 1 8; do
timeout 300 python bench_hmult.py --batch $B --steps 10 > gpurun_out/tc3_tc_b$B.json 2>> gpurun_out/tc3.err; echo "tc b$B rc=$?"
done
python - <<'PY'
import json
for n in ('tc_b1','tc_b8'):
    try:
        d=json.loads(open(f'gpurun_out/tc3_{n}.json').read().strip().splitlines()[-1])
        k=d['kernel_ms_per_call']
        print(n, round(d['value'],1), 'ops/s', round(d['ms_per_op'],4), d['gpu_launches'], {a:b['ms'] for a,b in k.items() if isinstance(b,dict)}, k['whole_call_ms'], 'ok', d['decrypts_to_product'], 'enc', round(d['encrypt']['value']), 'dec', round(d['decrypt']['value']))
    except Exception as e: print(n, 'ERR', e)
PY
tail -5 gpurun_out/tc3.err
