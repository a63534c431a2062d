#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bfv.py -m gpu -q -x -k "sharded" > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu7.log
python tools/run_sharded_hmult.py --preset c4 --steps 5 > gpurun_out/sharded_1.json 2> gpurun_out/sharded_1.err; cat gpurun_out/sharded_1.json; tail -2 gpurun_out/sharded_1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/run_sharded_hmult.py --preset c4 --steps 5 > gpurun_out/sharded_2.json 2> gpurun_out/sharded_2.err; cat gpurun_out/sharded_2.json; tail -3 gpurun_out/sharded_2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; python -c "
import json;d=json.load(open('gpurun_out/bench_2gpu.json'));print('2gpu value',round(d['value']),'e2e',round(d['e2e']['value']),d['n_gpus'],d['roundtrip_bit_exact'])"; tail -2 gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref2.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref2.json
