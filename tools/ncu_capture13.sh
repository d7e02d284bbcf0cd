#!/bin/bash
# round 2: GEMM captures after the lean MMA-issue path (xw fwd/dX block 6, dW block 6 / block 3)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2b}
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python tools/prof_step.py 1 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
# xw_gemm launches in one eager step: forward blocks 2..7 = #0..5, dense1 (split-K) #6, projections #7..10, backward head #11.., dX blocks 7..2 at the end
cap xw_fwd_b6    'xw_gemm_tc_v2_kernel' 4 1
cap xw_dx_b6     'xw_gemm_tc_v2_kernel' 17 1
# xty_gemm_tc_kernel launches (both instantiations): 12 recurrent-layer weight gradients, dense1, then blocks 7, 6, 5, 4, 3, 2
cap xty_dw_b6    'xty_gemm_tc_kernel' 14 1
cap xty_dw_b3    'xty_gemm_tc_kernel' 17 1
ls $OUT/*${TAG}*.ncu-rep
