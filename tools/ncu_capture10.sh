#!/bin/bash
# round-1 final: launch list of the bench command + full-set captures of the dominant kernels of the committed code
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1n}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/launches_${TAG}.log 2>&1; echo "launch list rc=$?"
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python tools/prof_step.py 1 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
# xw_gemm launches in one eager step: forward blocks 2..7 = #0..5, dense1 (split-K) #6, projections #7..10, backward head #11.., dX blocks 7..2 at the end
cap xw_fwd_b6    'xw_gemm_tc_v2_kernel' 4 1
cap xw_dx_b6     'xw_gemm_tc_v2_kernel' 17 1
cap gru_fwd_mma  'gru_fwd_mma_kernel' 0 1
cap gru_bwd_mma  'gru_bwd_mma_kernel' 0 1
cap dwrows_bwd   'dwconv3x3_rows_kernel<true' 0 2
cap actbwd_apply 'act_pool_bwd_kernel<true' 0 1
ls $OUT/*${TAG}*
