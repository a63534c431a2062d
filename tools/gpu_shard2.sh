#!/bin/bash
# limb-sharded multiply across real GPUs: parity (processes, CUDA IPC), then timing at batch 1 / 4
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/shard_topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/shardN_pytest_$N.log 2>&1; echo "sharded pytest rc=$?"; tail -15 gpurun_out/shardN_pytest_$N.log
for B in 1 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench_hmult.py --limb-sharded --batch $B --steps 20 > gpurun_out/shardN_${N}_b$B.json 2> gpurun_out/shardN_${N}_b$B.err; echo "b$B rc=$?"; tail -1 gpurun_out/shardN_${N}_b$B.json; tail -3 gpurun_out/shardN_${N}_b$B.err
done
