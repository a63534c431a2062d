#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/shard3_pytest_$N.log 2>&1; echo "sharded pytest rc=$?"; tail -5 gpurun_out/shard3_pytest_$N.log
for B in 1 4; do
for G in 1 0; do
FHE_B200_SHARD_GRAPH=$G timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench_hmult.py --limb-sharded --batch $B --steps 20 > gpurun_out/shard3_${N}_b${B}_g$G.json 2> gpurun_out/shard3_${N}_b${B}_g$G.err; echo "b$B graph=$G rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/shard3_${N}_b${B}_g$G.json').read().strip().splitlines()[-1]);print(round(d['value'],1), round(d['ms_per_op'],4), d['matches_single_gpu_bit_exact'], d['launches_per_op_per_rank'], round(d['speedup_vs_single_gpu_same_batch'],2))"
done; done
