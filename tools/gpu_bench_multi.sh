#!/bin/bash
# the driver's SCALE step for one N: bench.py under torchrun on all GPUs of the box
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_bench_$N.json 2> gpurun_out/final_bench_$N.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 )) s"
python - "$N" <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f'gpurun_out/final_bench_{n}.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
h=d.get('hmult',{}); print('hmult', {k:h.get(k) for k in ('value','ms_per_op','decrypts_to_product','error')})
for k,v in d.get('hmult_limb_sharded',{}).items(): print('limb-sharded', k, {x:v.get(x) for x in ('value','ms_per_op','speedup_vs_single_gpu_same_batch','matches_single_gpu_bit_exact','launches_per_op_per_rank','error')}, round(v.get('nvlink',{}).get('achieved_GBs_per_rank_out',0)))
PY
