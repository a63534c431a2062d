#!/bin/bash
# Static FMA-heavy-pipe accounting of the balanced NTT passes (no GPU needed): compiles tools/sass/balexp.cu for sm_100a and counts
# instruction classes per kernel (IMAD.WIDE / IMAD.HI = 4 pipe-cycles, IMAD / IMAD.X / IMAD.MOV / IMAD.IADD = 2 per warp instruction per
# SM sub-partition -- the model ncu's sm__pipe_fmaheavy_cycles_active confirms).  Pass A is reported for the (peeled) single-item copy of its loop, pass B for its main (convergent) loop.
#   tools/sass/run.sh [extra nvcc flags]
cd "$(dirname "$0")"
OUT=${TMPDIR:-/tmp}/balexp.cubin
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DFHE_BAL_EXPERIMENT "$@" -cubin -o $OUT balexp.cu -Xptxas -v 2> $OUT.log || { tail -20 $OUT.log; exit 1; }
grep -E "Used [0-9]+ registers" $OUT.log | sed 's/ptxas info    : //'
for k in bal_a_kernelILi8ELi16ELb1ELb0 bal_a_kernelILi8ELi16ELb1ELb1; do python count.py $OUT $k 64 $(python arange.py $OUT $k); done
for k in bal_b_kernelILi8ELi16ELb1ELb0 bal_b_kernelILi8ELi16ELb1ELb1; do python count.py $OUT $k 64 $(python brange.py $OUT $k); done
