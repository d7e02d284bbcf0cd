#!/bin/bash
# round-1: full-set captures of the HBM-bound element-wise kernels at block-6 shapes (64x66x9x512) and block 3, eager serial mode
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1h}
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python tools/prof_step.py 1 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
# launch order inside one step: forward blocks 2..7 use dwconv3x3_cb_kernel (block 1 is the C=1 kernel); backward runs blocks 7..2
cap dwfwd       'dwconv3x3_cb_kernel' 2 2      # forward blocks 4, 5
cap dwbwd       'dwconv3x3_cb_kernel' 6 2      # backward-data blocks 7, 6
cap actbwd      'act_pool_bwd_kernel' 0 4      # block 7 reduce+apply, block 6 reduce+apply
cap relu6bwd    'relu6_bwd_kernel' 0 2         # block 7 reduce + apply
cap actfwd      'act_pool_fwd_kernel' 3 2      # blocks 4, 5
cap dwbwdw      'dwconv3x3_bwd_weight_vec4' 0 1
cap xw2_dx_b6   'xw_gemm_tc_v2_kernel' 18 1
ls $OUT/*.ncu-rep
