#!/bin/bash
# launch list of ONE training step (all kernels, per-launch duration) + source-level captures of the two dominant GEMM kernels
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1d}
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_step_${TAG}.csv python tools/prof_step.py 1 > $OUT/launches_step.log 2>&1
echo "launch list rc=$?"
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python tools/prof_step.py 1 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
cap xw2_fwd_b3 'xw_gemm_tc_v2_kernel' 1 1
cap xty_b3     'xty_gemm_tc_kernel' 17 1
cap gru_bwd    'gru_bwd_cluster_kernel' 0 1
