#!/bin/bash
# one compute-sanitizer tool per gpurun call (B200_PROFILING.md): TOOL=memcheck|racecheck|synccheck|initcheck
TOOL=${TOOL:-memcheck}
mkdir -p gpurun_out
SEL='tests/test_gpu_bfv.py::test_bfv_multiply_relin_vs_oracle[mid] tests/test_gpu_bfv.py::test_bfv_multiply_relin_vs_oracle[mid-fused] tests/test_gpu_ntt.py::test_balanced_passes_ragged_batches_and_limb_ranges[13] tests/test_gpu_sharded.py::test_virtual_ranks_equal_single_gpu[2-0]'
python -m pytest $SEL -x -q -m gpu > gpurun_out/san_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20 python -m pytest $SEL -x -q -m gpu > gpurun_out/sanitizer_$TOOL.log 2>&1
echo "$TOOL rc=$?"; tail -12 gpurun_out/sanitizer_$TOOL.log
