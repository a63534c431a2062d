#!/bin/bash
# A/B on one box: BFV multiply / encrypt / decrypt on one stream (FHE_B200_HMULT_STREAMS=1) vs the two-half split (default)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bfv.py tests/test_golden.py -m gpu -q -x > gpurun_out/pytest_streams.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_streams.log
for b in ${BATCHES:-2 4 8 16}; do for s in 1 2; do
  FHE_B200_HMULT_STREAMS=$s timeout 300 python bench_hmult.py --batch $b --steps 5 2>/dev/null > gpurun_out/abs_$s.json
  python -c "
import json;d=json.load(open('gpurun_out/abs_$s.json'));print('batch $b streams=$s',round(d['value'],1),round(d['ms_per_op'],4),d['decrypts_to_product'],'e2e',round(d['e2e']['value'],1),d['e2e']['matches_device_path'],'square',round(d['square']['value'],1),'encrypt',round(d['encrypt']['value']),'decrypt',round(d['decrypt']['value']),'relin',round(d['relinearize']['value']),'rotate',round(d['rotate']['value']))"
done; done
