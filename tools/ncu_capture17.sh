#!/bin/bash
# round 2: full-set captures of the fused depthwise backward kernel (dwconv_bwd_fused.cu), blocks 7..2 of one eager step
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2q}
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python $5 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
cap dwfused 'dwconv3x3_bwd_fused_kernel' 0 6 "tools/prof_step.py 1"
ls $OUT/*${TAG}*.ncu-rep
