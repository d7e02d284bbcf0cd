#!/bin/bash
# round-1 final captures: full-set (source-level) profiles of the latency-bound kernels + the K=512 pointwise GEMM, eager serial mode
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1f}
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python tools/prof_step.py 1 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
cap gru_fwd     'gru_fwd_cluster_kernel' 0 1
cap gru_bwd     'gru_bwd_cluster_kernel' 0 1
cap ctc_loss    'ctc_loss_grad_kernel' 0 1
cap xw2_fwd_b6  'xw_gemm_tc_v2_kernel' 4 1
cap actpool_app_b3 'act_pool_bwd_kernel<1, 2, 2>' 0 1
cap dwfwd_b4    'dwconv3x3_cb_kernel<0, 1>' 2 1
ls -la $OUT | head -30
