#!/bin/bash
# N-GPU checks (N = number of visible GPUs; gpurun --gpus N): headline bench and batch-sharded HMult
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_mg_${N}gpu.json 2> gpurun_out/bench_mg_${N}gpu.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_mg_${N}gpu.json').read().strip().splitlines()[-1]);print('${N}gpu value',round(d['value']),'e2e',round(d['e2e']['value']),d['roofline']['per_kernel_ms'],d['roundtrip_bit_exact'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench_hmult.py --batch 8 --steps 5 > gpurun_out/hmult_mg_${N}gpu.json 2> gpurun_out/hmult_mg_${N}gpu.err; echo "hmult rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/hmult_mg_${N}gpu.json').read().strip().splitlines()[-1]);print('${N}gpu hmult',round(d['value'],1),d['n_gpus'],round(d['ms_per_op'],3),d['decrypts_to_product'],'e2e',round(d['e2e']['value'],1))"
tail -3 gpurun_out/hmult_mg_${N}gpu.err
