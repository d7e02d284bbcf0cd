#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch list of the bench command + full captures of one representative launch
# of each hot kernel (block-3 shapes: M = 64*132*36 pixels).  Outputs under gpurun_out/.
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-r1}
cap() {   # name regex skip count
  timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python tools/prof_step.py 1 > $OUT/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
cap xw_fwd_b3   'xw_gemm_tc_kernel' 1 1
cap xty_b3      'xty_gemm_tc_kernel' 17 1
cap actpool_b3  'act_pool_bwd_kernel' 8 2
cap relu6bn_b3  'relu6_bwd_kernel' 8 2
cap dwbwdw_b3   'dwconv3x3_bwd_weight_vec4' 4 1
cap dwfwd_b3    'dwconv3x3_vec4' 1 1
cap dwbwdd_b3   'dwconv3x3_vec4' 10 1
cap gru_fwd     'gru_fwd_cluster_kernel' 0 1
cap gru_bwd     'gru_bwd_cluster_kernel' 0 1
cap actfwd_b3   'act_pool_fwd_kernel' 2 1
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
echo "launch list rc=$?"
ls -la $OUT | head -40
