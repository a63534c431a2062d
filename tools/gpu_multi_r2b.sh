#!/bin/bash
# round 2, final state on N GPUs of one box: multi-process parity of the limb-sharded multiply, headline bench (with hmult + hmult_limb_sharded)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_sharded.py "tests/test_gpu_ntt.py::test_two_devices_in_one_process" -x -q -m gpu > gpurun_out/r2f_multi_pytest_$N.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_multi_pytest_$N.log
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2f_bench_$N.json 2> gpurun_out/r2f_bench_$N.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 )) s"
grep -v "^\*\|OMP_NUM" gpurun_out/r2f_bench_$N.err | tail -3
python - "$N" <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f'gpurun_out/r2f_bench_{n}.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
h=d.get('hmult',{}); print('hmult', {k:h.get(k) for k in ('value','ms_per_op','decrypts_to_product','parallelism','error')})
for k,v in d.get('hmult_limb_sharded',{}).items(): print('limb-sharded', k, {x:v.get(x) for x in ('value','ms_per_op','speedup_vs_single_gpu_same_batch','matches_single_gpu_bit_exact','error')}, v.get('nvlink'))
PY
