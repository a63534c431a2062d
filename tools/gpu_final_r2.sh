#!/bin/bash
# round 2, final state on one GPU: whole GPU test tier, smoke, reference arm, headline bench (what the driver runs at round end, minus the sweep)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; echo "ref rc=$?"
T0=$(date +%s); python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2f_ref.json').read().strip().splitlines()[-1])
print('value', d['value'], 'sustained', d['sustained']['value'], 'e2e', d['e2e']['value'], 'ref', r['value'], 'e2e/ref', d['e2e']['value']/r['value'])
print('roofline', {k:d['roofline'][k] for k in ('bound','achieved','peak','frac','kernel')}, d['roofline']['whole_step'], d['roofline']['register_only_butterfly_loop'])
print('per kernel', d['roofline']['per_kernel_ms'])
h=d.get('hmult',{}); print('hmult', {k:h.get(k) for k in ('value','ms_per_op','decrypts_to_product','gpu_launches','error')}, {k:h[k]['value'] for k in ('encrypt','decrypt','square','relinearize','rotate','e2e') if k in h})
PY
