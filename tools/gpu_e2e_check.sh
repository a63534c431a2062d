#!/bin/bash
# host-buffer HMult path (three-stream chunk pipeline): parity tests + end-to-end rates at config 4 and config 2 (run under gpurun)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_bfv.py tests/test_gpu_compat.py -m gpu -q -x > gpurun_out/pytest_gpu_e2e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_e2e.log
for b in 8 16; do timeout 300 python bench_hmult.py --batch $b --steps 5 2>/dev/null | python -c "
import json,sys;d=json.load(sys.stdin);print('c4 b$b hmult',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['e2e']['matches_device_path'])"; done
for b in 64 256; do timeout 300 python bench_hmult.py --preset c2 --batch $b --steps 5 2>/dev/null | python -c "
import json,sys;d=json.load(sys.stdin);print('c2 b$b hmult',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['e2e']['matches_device_path'])"; done
