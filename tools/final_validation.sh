#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun): tests, smoke, bench lines of all configs, reference arm,
# launch lists.   usage: tools/final_validation.sh <tag>
tag=${1:-r2z}; o=gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > $o/${tag}_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/${tag}_smoke.txt 2>&1
python bench.py > $o/${tag}_cfg1.json 2> $o/${tag}_cfg1.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_reference.json 2> $o/${tag}_reference.err
for c in 2 3 4; do
  python bench.py --config $c --steps 10 --warmup 3 --no-cpu > $o/${tag}_cfg$c.json 2> $o/${tag}_cfg$c.err
done
# launch lists: the whole default command for config 1 (first 400 launches: setup, warm-up, timed steps, e2e legs),
# one warm step of the patch kernels for the large configs
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_cfg1_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
for c in 3 2 4; do
  lps=$(python -c "import json;print(int(json.loads(open('$o/${tag}_cfg$c.json').read().strip().splitlines()[-1])['roofline']['launches_per_step']))")
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"patch_|korn" -s $((3*lps)) -c $lps --csv \
      --log-file $o/${tag}_cfg${c}_launches.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
done
cat $o/${tag}_tests.txt $o/${tag}_smoke.txt
python - <<PY
import json
for c in (1,2,3,4):
    try:
        d=json.loads(open(f"$o/${tag}_cfg{c}.json").read().strip().splitlines()[-1])
        print(c, round(d["ms_per_step"],4), f'{d["value"]:.3e}', "frac", round(d["roofline"]["frac"],4), "e2e", f'{d["e2e"]["value"]:.3e}', "cold", d.get("e2e_cold") and round(d["e2e_cold"]["seconds"],4), "traffic", d["roofline"]["traffic"], "fp64", d["roofline"].get("fp64",{}).get("frac"))
    except Exception as e: print(c, "ERR", e)
try:
    d=json.loads(open("$o/${tag}_reference.json").read().strip().splitlines()[-1]); print("reference", d["value"], d["cpu_baseline"])
except Exception as e: print("ref ERR", e)
PY
