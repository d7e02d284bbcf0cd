set -u
OUT=gpurun_out; mkdir -p $OUT
export CRNN_GRAPH=0 CRNN_OVERLAP=0
cap() { timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/prof_r2y_$1 python $5 > $OUT/ncu_$1.log 2>&1; echo "$1 rc=$?"; }
cap xty_tma_b6 'xty_gemm' 14 1 "tools/prof_step.py 1"
CRNN_XTY_V1=1 cap xty_v1_b6 'xty_gemm' 14 1 "tools/prof_step.py 1"
CRNN_XTY_2MMA=1 cap xty_2mma_b6 'xty_gemm' 14 1 "tools/prof_step.py 1"
ls -la $OUT/prof_r2y_*
