#!/bin/bash
mkdir -p gpurun_out
./tools/microbench3.bin > gpurun_out/microbench3.jsonl 2>&1; cat gpurun_out/microbench3.jsonl
for b in 1 4; do timeout 300 python bench_hmult.py --batch $b --steps 5 2>gpurun_out/hmult16_b$b.err > gpurun_out/hmult16_b$b.json; cat gpurun_out/hmult16_b$b.json; done
