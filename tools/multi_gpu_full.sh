#!/bin/bash
# as multi_gpu_check.sh plus the limb-sharded multiply (NCCL all-gathers)
./tools/multi_gpu_check.sh
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/run_sharded_hmult.py --steps 10 > gpurun_out/sharded_mg_$N.json 2> gpurun_out/sharded_mg_$N.err; echo "sharded rc=$?"; tail -1 gpurun_out/sharded_mg_$N.json
