#!/bin/bash
# round-1: full-set capture of the LSTM tensor-core cluster kernels (layer 1 launch)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1l}
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python tools/prof_step.py 1 lstm > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
cap lstm_fwd_mma 'lstm_fwd_mma_kernel' 0 1
[ "${2:-}" = "bwd" ] && cap lstm_bwd_mma 'lstm_bwd_mma_kernel' 0 1
ls $OUT/*${TAG}*.ncu-rep
