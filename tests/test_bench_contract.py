"""The reference arm of bench.py runs on CPU: check the JSON contract of its line (the b200 arm
needs a GPU and is exercised by the driver)."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-n", "16"],
        capture_output=True, text=True, timeout=600, cwd=ROOT,
    )
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "equilibrated patches/sec" and line["unit"] == "patches/s"
    assert line["dtype"] == "f64" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert "workload" in line["config"]
    cb = line["cpu_baseline"]
    # "reference": oracle/_ref (the reference's sources compiled unchanged) is present; "port": the oracle restatement
    from oracle import pyref as pr

    assert cb["kind"] == ("reference" if pr.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert line["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = {**os.environ, "RANK": "1", "WORLD_SIZE": "2"}
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
