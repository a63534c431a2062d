#!/bin/bash
# what the driver runs at round end, in order: GPU tests, smoke(), the reference arm, the headline bench (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/final_ref.json
T0=$(date +%s); python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 )) s"
python -c "
import json;d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print({k:(v if not isinstance(v,dict) else '...') for k,v in d.items()})
print('roofline',d['roofline']); print('int_pipe',d['int_pipe']); print('cpu',d.get('cpu_baseline')); print('e2e',d['e2e']); print('hmult',{k:v for k,v in d.get('hmult',{}).items() if k in ('value','batch','ms_per_op','decrypts_to_product','int_pipe_floor','e2e','encrypt','decrypt')})"
