#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"; wc -c gpurun_out/bench_2gpu.json; cut -c1-400 gpurun_out/bench_2gpu.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_2gpu.err | tail -15 | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-330
