#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_compat.py -m gpu -q -x > gpurun_out/pytest_gpu38.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu38.log
python tools/bench_c1.py | tee gpurun_out/bench_c1.jsonl
