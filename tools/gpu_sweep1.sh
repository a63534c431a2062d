#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/sweep_c5.py --out gpurun_out/sweep_c5_1.jsonl > gpurun_out/sweep_c5_1.log 2>&1; echo "sweep rc=$?"; tail -3 gpurun_out/sweep_c5_1.log | cut -c1-300
