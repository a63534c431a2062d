#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ntt.py -m gpu -q -x > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest ntt rc=$?"; tail -3 gpurun_out/pytest_gpu13.log
python tools/prof_ntt.py 64 1 > gpurun_out/prof_plain13.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ntt_tile_fwd -c 1 -o gpurun_out/prof_tile_r13 python tools/prof_ntt.py 64 1 > gpurun_out/ncu_full13.log 2>&1
echo "ncu rc=$?"
