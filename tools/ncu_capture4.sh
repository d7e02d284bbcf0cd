#!/bin/bash
# source-level captures of the non-GEMM hot kernels (one launch each) from one training step
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1e}
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python tools/prof_step.py 1 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
cap actpool_red_b3 'act_pool_bwd_kernel<0, 2, 2>' 0 1
cap actpool_app_b3 'act_pool_bwd_kernel<1, 2, 2>' 0 1
cap ctc_loss       'ctc_loss_grad_kernel' 0 1
cap gru_fwd        'gru_fwd_cluster_kernel' 0 1
cap dwbwdw         'dwconv3x3_bwd_weight_vec4' 5 1
cap actfwd_b3      'act_pool_fwd_kernel' 2 1
cap stn_bwd        'stn_trunk_bwd_kernel' 0 1
