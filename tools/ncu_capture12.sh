#!/bin/bash
# round 2, first session: full-set captures of the kernels VERDICT r1 found without a record (dW GEMM, beam, final CTC kernel, STN, Adam)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2a}
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_${TAG}_$1 python $5 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
# xty_gemm_tc_kernel<256> launches of one eager step: 12 recurrent-layer weight gradients, then blocks 7, 6, 5, 4, 3
cap xty_dw_b6   'xty_gemm_tc_kernel<256>' 13 1 "tools/prof_step.py 1"
cap xty_dw_b3   'xty_gemm_tc_kernel<256>' 16 1 "tools/prof_step.py 1"
cap ctc_loss    'ctc_loss_grad_kernel' 0 1 "tools/prof_step.py 1"
cap stn         'stn_' 0 6 "tools/prof_step.py 1"
cap adam        'adam_kernel|sumsq_kernel' 0 2 "tools/prof_step.py 1"
cap beam_cfg3   'ctc_beam_kernel' 1 1 "tools/prof_beam.py"
ls $OUT/*${TAG}*.ncu-rep
