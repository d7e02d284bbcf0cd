#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1b}
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python tools/prof_step.py 1 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
cap xw2_fwd_b3 'xw_gemm_tc_v2_kernel' 1 1
cap xw2_fwd_b7 'xw_gemm_tc_v2_kernel' 5 1
cap xty_b3     'xty_gemm_tc_kernel' 17 1
cap gru_fwd    'gru_fwd_cluster_kernel' 0 1
