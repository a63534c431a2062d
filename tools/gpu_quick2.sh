#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/quick2_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/quick2_pytest.log
python tools/prof_lincomb.py 4 2> gpurun_out/quick2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().rsplit('prof_lincomb ok',1)[0])
for k,v in d.items(): print(k, v['tc'], v['mma'], 'same' if v['sha_tc']==v['sha_mma'] else 'DIFF')"
for B in 1 8; do timeout 300 python bench_hmult.py --batch $B --steps 10 > gpurun_out/quick2_b$B.json 2>> gpurun_out/quick2.err; python -c "
import json;d=json.loads(open('gpurun_out/quick2_b$B.json').read().strip().splitlines()[-1]);print($B, round(d['value'],1), d['kernel_ms_per_call']['lincomb'], d['decrypts_to_product'], 'enc', round(d['encrypt']['value']), 'dec', round(d['decrypt']['value']))"; done
