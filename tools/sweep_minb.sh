#!/bin/bash
# resident-CTA sweeps of the degree-3 lane-per-cell kernel (config 2 at 2048^2) and of the fused stress kernel (config 3)
out=gpurun_out/$1_minb.txt; : > $out
for mb in 2 3 4; do
  echo "== SE k=3 (config 2 layout, n=2048) EQLB_KW3_MINB=$mb" >> $out
  EQLB_KW3_MINB=$mb python bench.py --config 2 --n 2048 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> $out
done
for mb in 2 3 4; do
  echo "== EV k=3 n=1024 EQLB_KW3_MINB=$mb" >> $out
  EQLB_KW3_MINB=$mb python bench.py --path ev --k 3 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> $out
done
for mb in 3 4; do
  echo "== elasticity (config 3) EQLB_K2S_MINB=$mb" >> $out
  EQLB_K2S_MINB=$mb python bench.py --config 3 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> $out
done
cat $out
