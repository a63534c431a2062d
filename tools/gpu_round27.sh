#!/bin/bash
# 2-GPU checks: headline bench (batch-sharded, weak scaling) and the limb-sharded multiply
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench27_2gpu.json 2> gpurun_out/bench27_2gpu.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench27_2gpu.json').read().strip().splitlines()[-1]);print('2gpu value',round(d['value']),'e2e',round(d['e2e']['value']),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/run_sharded_hmult.py --steps 10 > gpurun_out/sharded27_2.json 2> gpurun_out/sharded27_2.err; echo "sharded rc=$?"; tail -1 gpurun_out/sharded27_2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench27_ref2.json 2> gpurun_out/bench27_ref2.err; echo "ref rc=$?"; tail -c 400 gpurun_out/bench27_ref2.json
