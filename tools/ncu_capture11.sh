#!/bin/bash
# round-1 final (2): full-set captures of the fused-reduction depthwise backward-data kernel and the LSTM recurrence kernels
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1o}
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python tools/prof_step.py 1 $5 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
cap dwrows_bwd_red 'dwconv3x3_rows_kernel' 6 1 gru      # launches 0..5 forward, 6 = block 7 backward-data (fused reduction of block 6)
cap lstm_fwd_mma   'lstm_fwd_mma_kernel' 0 1 lstm
cap lstm_bwd_mma   'lstm_bwd_mma_kernel' 0 1 lstm
ls $OUT/*${TAG}*.ncu-rep
