#!/bin/bash
# final multi-GPU record: parity, bench line (hmult + hmult_limb_sharded with graph replay), round-1 NCCL all-gather path for comparison
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/final_multi_pytest_$N.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_multi_pytest_$N.log
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
for B in 1 8; do
FHE_B200_SHARD_GRAPH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench_hmult.py --limb-sharded --batch $B --steps 20 > gpurun_out/final_shard_${N}_b${B}_nograph.json 2> /dev/null
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench_hmult.py --limb-sharded --batch $B --steps 20 > gpurun_out/final_shard_${N}_b${B}_graph.json 2> /dev/null
python -c "
import json
for t in ('nograph','graph'):
    d=json.loads(open('gpurun_out/final_shard_${N}_b${B}_'+t+'.json').read().strip().splitlines()[-1]);print('B=$B',t,round(d['value'],1), round(d['ms_per_op'],4), d['matches_single_gpu_bit_exact'], round(d['speedup_vs_single_gpu_same_batch'],2))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 tools/run_sharded_hmult.py --steps 10 > gpurun_out/final_nccl_allgather_$N.json 2> gpurun_out/final_nccl_allgather_$N.err; echo "r01 NCCL all-gather path rc=$?"; tail -1 gpurun_out/final_nccl_allgather_$N.json | cut -c1-400
