#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" timeout 300 python bench.py --steps 20 --no-hmult --no-cpu-baseline 2>/dev/null | python -c "
import sys,json;d=json.load(sys.stdin);print('$*','value',round(d['value']),d['roofline']['per_kernel_ms']['tile_fwd'],d['roofline']['per_kernel_ms']['tile_inv'],d['roundtrip_bit_exact'])"; }
run FHE_B200_FUSED_PG=4 FHE_B200_FUSED_LBK=8 FHE_B200_FUSED_LEAD=2
run FHE_B200_FUSED_PG=4 FHE_B200_FUSED_LBK=8 FHE_B200_FUSED_LEAD=1
run FHE_B200_FUSED_PG=8 FHE_B200_FUSED_LBK=8 FHE_B200_FUSED_LEAD=1
run FHE_B200_FUSED_PG=4 FHE_B200_FUSED_LBK=16 FHE_B200_FUSED_LEAD=1
run FHE_B200_FUSED_PG=2 FHE_B200_FUSED_LBK=16 FHE_B200_FUSED_LEAD=2
run FHE_B200_FUSED_PG=8 FHE_B200_FUSED_LBK=32 FHE_B200_FUSED_LEAD=1
run FHE_B200_FUSED_PG=4 FHE_B200_FUSED_LBK=32 FHE_B200_FUSED_LEAD=3
