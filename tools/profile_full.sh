#!/bin/bash
# `ncu --set full` of one warm step of a BASELINE config (run under gpurun, one config per call is enough):
#   tools/profile_full.sh <tag> <config> <launches per step>
tag=$1; c=$2; lps=$3
python bench.py --config $c --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_cfg${c}_plain.json 2> gpurun_out/${tag}_cfg${c}_plain.err &&
ncu --set full --clock-control none -k regex:"patch_|korn" -s $((3*lps)) -c $lps \
    -o gpurun_out/${tag}_cfg${c}_full python bench.py --config $c --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_cfg${c}_ncu.log 2>&1 &&
ncu -i gpurun_out/${tag}_cfg${c}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_cfg${c}_raw.csv 2>/dev/null
# gpurun_out is limited to 64 MiB: keep the report itself only when it is small
[ $(stat -c %s gpurun_out/${tag}_cfg${c}_full.ncu-rep 2>/dev/null || echo 0) -gt 12000000 ] && rm -f gpurun_out/${tag}_cfg${c}_full.ncu-rep
true
