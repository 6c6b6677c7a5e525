#!/bin/bash
# end-to-end (host buffers) throughput of the headline config against the number of pipeline stages
for st in 8 12 16 24 32 48; do
  EQLB_PIPE_STAGES=$st python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stages', $st, 'e2e %.4e' % d['e2e']['value'], 'ms', round(2099201/d['e2e']['value']*1e3,3))"
done
