#!/bin/bash
# round 2: beam kernel capture (configs[3]) after the sorted-candidate walk
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2f}
timeout -k 5 240 ncu --set full --clock-control none --import-source on -k "regex:ctc_beam_kernel" -s 1 -c 1 -f -o $OUT/prof_${TAG}_beam_cfg3 python tools/prof_beam.py > $OUT/ncu_beam_cfg3.log 2>&1; echo "beam rc=$?"
