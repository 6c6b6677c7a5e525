#!/bin/bash
# device-resident ms/step of the headline config for different spatial chunk sizes
for c in 0 2097152 1048576 524288 262144 131072; do
  if [ "$c" = "0" ]; then unset EQLB_CHUNK_CELLS; else export EQLB_CHUNK_CELLS=$c; fi
  python bench.py --steps 50 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk_cells', '$c', 'ms', round(d['ms_per_step'],4), 'launches/step', d['roofline']['launches_per_step'])"
done
