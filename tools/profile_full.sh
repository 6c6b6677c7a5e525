#!/bin/bash
# `ncu --set full` of one warm step of a BASELINE config (run under gpurun, one config per call is enough):
#   tools/profile_full.sh <tag> <config> <launches per step>
tag=$1; c=$2; lps=$3
python bench.py --config $c --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_cfg${c}_plain.json 2> gpurun_out/${tag}_cfg${c}_plain.err &&
ncu --set full --clock-control none -k regex:"patch_|korn" -s $((3*lps)) -c $lps \
    -o gpurun_out/${tag}_cfg${c}_full python bench.py --config $c --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_cfg${c}_ncu.log 2>&1
