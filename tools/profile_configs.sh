#!/bin/bash
# Launch lists (per-launch device time, one warm step) of the four BASELINE configs; run under gpurun.
# usage: tools/profile_configs.sh <tag>
tag=${1:-r2}
for c in 1 3 2 4; do
  python bench.py --config $c --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_cfg${c}_plain.json 2> gpurun_out/${tag}_cfg${c}_plain.err || continue
  lps=$(python -c "import json;print(int(json.load(open('gpurun_out/${tag}_cfg${c}_plain.json'))['roofline']['launches_per_step']))")
  # skip setup kernels + 3 warm-up steps, record one step
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"patch_|korn" -s $((3*lps)) -c $lps --csv \
      --log-file gpurun_out/${tag}_cfg${c}_launches.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
done
