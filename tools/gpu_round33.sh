#!/bin/bash
mkdir -p gpurun_out
python tools/sweep_ntt.py > gpurun_out/sweep_ntt_r01.jsonl 2> gpurun_out/sweep_ntt.err; echo "sweep rc=$?"; cat gpurun_out/sweep_ntt_r01.jsonl | cut -c1-200
