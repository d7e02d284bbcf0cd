#!/bin/bash
# round-1: full-set captures of the row-marching depthwise kernels (forward blocks 2 and 6, backward-data block 7, backward-weight block 7)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1k}
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python tools/prof_step.py 1 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
cap dwrows_fwd   'dwconv3x3_rows_kernel' 0 5       # forward blocks 2..6
cap dwrows_bwd   'dwconv3x3_rows_kernel' 6 2       # backward-data blocks 7, 6
cap dwrows_bwdw  'dwconv3x3_rows_bwd_weight_kernel' 0 2
ls $OUT/*${TAG}*.ncu-rep
