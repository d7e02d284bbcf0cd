#!/bin/bash
# round 2 (after the depthwise fusions): launch list of the bench command, DRAM bytes of one eager step, full-set captures of the new kernels
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2s}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/launches_${TAG}.log 2>&1; echo "launch list rc=$?"
bash tools/ncu_step_dram.sh ${TAG}
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python $5 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
# fused backward kernel: blocks 7..2 of one eager step = launches #0..5 (7, 5, 3, 2 with the fused BN2 reduction); forward: blocks 2, 3, 5, 7 = #0..3
cap dwbwd_fused 'dwconv3x3_bwd_fused_kernel' 0 6 "tools/prof_step.py 1"
cap dwfwd_fused 'dwconv3x3_fwd_fused_kernel' 0 4 "tools/prof_step.py 1"
ls $OUT/*${TAG}*
