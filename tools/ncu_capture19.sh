#!/bin/bash
# end of round 2 (committed code): launch list of the bench command, DRAM bytes of one eager step, full-set captures of the kernels that
# changed in the last session (BatchNorm finalize tails in the statistics-producing kernels, beam decode) + the dominant GEMM launches
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2x}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/launches_${TAG}.log 2>&1; echo "launch list rc=$?"
( export CRNN_GRAPH=0 CRNN_OVERLAP=0
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 700 --csv \
      --log-file $OUT/step_dram_${TAG}.csv python tools/prof_step.py 3 > $OUT/step_dram_${TAG}.log 2>&1; echo "step dram rc=$?" )
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python $5 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
# xw_gemm launches in one eager step: forward blocks 2..7 = #0..5, dense1 (split-K) #6, projections #7..10, backward head #11.., dX blocks 7..2 at the end
cap xw_fwd_b6    'xw_gemm_tc_v2_kernel' 4 1 "tools/prof_step.py 1"
cap xw_dx_b6     'xw_gemm_tc_v2_kernel' 17 1 "tools/prof_step.py 1"
cap xty_dw_b6    'xty_gemm_tc_kernel' 14 1 "tools/prof_step.py 1"
cap dwfwd_fused  'dwconv3x3_fwd_fused_kernel' 0 2 "tools/prof_step.py 1"
cap beam_cfg3    'ctc_beam_kernel' 1 1 "tools/prof_beam.py"
ls $OUT/*${TAG}*.ncu-rep
