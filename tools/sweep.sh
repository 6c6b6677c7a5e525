#!/bin/bash
# device + e2e numbers of every configuration on the 1024x1024 crossed mesh (run from the repo root on a GPU box)
for cfg in "--path ev --k 2" "--path se --k 2" "--path ev --k 1" "--path se --k 1" "--path ev --k 3" "--path se --k 3" "--path se --k 2 --nrhs 2" "--stress --k 2" "--stress --k 2 --nrhs 3"; do
  python bench.py --no-cpu --steps 20 --warmup 3 $cfg 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$cfg |', round(d['ms_per_step'],3), 'ms |', '%.3g'%d['value'], 'p/s | e2e %.3g'%d['e2e']['value'], '| frac %.3f'%d['roofline']['frac'], '| launches/step', d['roofline']['launches_per_step'])"
done
