#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bfv.py tests/test_gpu_compat.py -m gpu -q -x > gpurun_out/pytest_gpu14.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu14.log
for tpc in 0 1 2; do for b in 1 4 8; do FHE_B200_LINCOMB_TPC=$tpc timeout 300 python bench_hmult.py --batch $b --steps 5 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print('tpc',$tpc,'hmult b',d['batch'],round(d['value'],1),round(d['ms_per_op'],3),d['decrypts_to_product'],'e2e',round(d['e2e']['value'],1))"; done; done
