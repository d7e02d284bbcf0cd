#!/bin/bash
# DRAM bytes of ONE training step: one ncu pass (3 metrics, no replay sets) over two eager steps of tools/prof_step.py;
# tools/step_dram_summary.py turns the CSV into profiles/r02_step_dram.json (bench.py reads it: step_level.measured_dram_bytes_per_step)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2}
export CRNN_GRAPH=0 CRNN_OVERLAP=0
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file $OUT/step_dram_${TAG}.csv python tools/prof_step.py 3 > $OUT/step_dram_${TAG}.log 2>&1
echo "step dram rc=$?"
