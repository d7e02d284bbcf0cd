#!/bin/bash
# round-1: full-set captures of the register-resident mma.sync GRU kernels (layer 1 launch of each)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1i}
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python tools/prof_step.py 1 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
cap gru_fwd_mma 'gru_fwd_mma_kernel' 0 1
cap gru_bwd_mma 'gru_bwd_mma_kernel' 0 1
ls $OUT/*${TAG}*.ncu-rep
