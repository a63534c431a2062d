#!/bin/bash
# first GPU check of the limb-sharded multiply (one GPU: world = 1 and virtual ranks), then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/shard1_pytest.log 2>&1; echo "sharded pytest rc=$?"; tail -15 gpurun_out/shard1_pytest.log
timeout 300 python bench_hmult.py --limb-sharded --batch 1 --steps 20 > gpurun_out/shard1_b1.json 2> gpurun_out/shard1_b1.err; echo "b1 rc=$?"; cat gpurun_out/shard1_b1.json; tail -3 gpurun_out/shard1_b1.err
timeout 300 python bench_hmult.py --limb-sharded --batch 4 --steps 10 > gpurun_out/shard1_b4.json 2> gpurun_out/shard1_b4.err; echo "b4 rc=$?"; cat gpurun_out/shard1_b4.json; tail -3 gpurun_out/shard1_b4.err
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/shard1_full.log 2>&1; echo "full pytest rc=$?"; tail -5 gpurun_out/shard1_full.log
