#!/bin/bash
# round 2: CTA-pair (cta_group::2) variant of the xw GEMM, block 6 forward
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2k}
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python tools/prof_step.py 1 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
cap xw_pair_fwd_b6    'xw_gemm_tc_v2_kernel' 4 1
