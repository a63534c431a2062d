#!/bin/bash
# 8-GPU (or however many are visible) scaling checks: headline bench, batch-sharded HMult, limb-sharded HMult
./tools/gpu_round31.sh
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/run_sharded_hmult.py --steps 10 > gpurun_out/sharded34_$N.json 2> gpurun_out/sharded34_$N.err; echo "sharded rc=$?"; tail -1 gpurun_out/sharded34_$N.json
