#!/usr/bin/env python
"""Turn `ncu --set full` captures of ONE warm step of `bench.py --config C` into the record bench.py prints in
`roofline.traffic` / `roofline.fp64` (profiles/ncu_kernels.json, keyed by config, stamped with the hash of the
kernel sources so that bench.py withholds stale numbers).

    tools/ncu_to_json.py <config> <report.ncu-rep | raw.csv> <npatch>      (run where ncu is installed; no GPU needed)
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    cfg, rep, npatch = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
    if rep.endswith(".csv"):  # already exported on the GPU box (`ncu -i ... --page raw --csv`)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    raw_csv = os.path.join(ROOT, "profiles", f"r2_cfg{cfg}_ncu_raw.csv")
    with open(raw_csv, "w") as fh:  # the full metric table of every launch of the step (ncu --page raw)
        fh.write(out)
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]

    def col(r, name):
        return float(r[ix[name]].replace(",", "")) if name in ix and r[ix[name]] not in ("", "n/a") else 0.0

    def tunit(name):  # -> microseconds
        return {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[rows[1][ix[name]].strip().lower()]

    def flops(r):  # executed FP64 flop of the launch: per-cycle rates x elapsed cycles
        cyc = col(r, "smsp__cycles_elapsed.avg")
        f = lambda op: col(r, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") * cyc
        return f("dfma"), f("dmul"), f("dadd")

    per = []
    for r in data:
        name = r[ix["Kernel Name"]]
        per.append(dict(
            name=name, us=col(r, "gpu__time_duration.sum") * tunit("gpu__time_duration.sum"),
            dram=col(r, "dram__bytes_read.sum") * unit(rows[1][ix["dram__bytes_read.sum"]])
            + col(r, "dram__bytes_write.sum") * unit(rows[1][ix["dram__bytes_write.sum"]]),
            dfma=flops(r)[0], dmul=flops(r)[1], dadd=flops(r)[2],
            inst=col(r, "smsp__inst_executed.sum"),
            regs=col(r, "launch__registers_per_thread"),
            lsu=col(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            issue=col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            fp64=col(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            dram_pct=col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ))
    lps = len(per)
    tot_us = sum(p["us"] for p in per)
    dom = max(per, key=lambda p: p["us"])
    import re
    base = re.search(r"(\w+_kernel)", dom["name"]).group(1)
    flops = sum(2 * p["dfma"] + p["dmul"] + p["dadd"] for p in per)
    rec = {
        "kernel": base, "launches_per_step": lps, "traffic": sum(p["dram"] for p in per) / lps,
        "flop_per_patch": flops / npatch, "capture": f"{os.path.basename(rep)} (ncu --set full --clock-control none of one warm step, tools/profile_full.sh; all metrics: profiles/{os.path.basename(raw_csv)})", "kernel_source_hash": bench.kernel_source_hash(),
        "step_us_under_ncu": tot_us,
        "warp_inst_per_patch": sum(p["inst"] for p in per) / npatch,
        "launches": [{"kernel": re.search(r"(\w+_kernel<[^>]*>)", p["name"]).group(1), "us": round(p["us"], 1), "dram_MB": round(p["dram"] / 1e6, 1),
                      "gflop": round((2 * p["dfma"] + p["dmul"] + p["dadd"]) / 1e9, 2),
                      "registers": int(p["regs"]), "lsu_pipe_pct": round(p["lsu"], 1), "issue_pct": round(p["issue"], 1),
                      "fp64_pipe_pct": round(p["fp64"], 1), "dram_pct": round(p["dram_pct"], 1)} for p in per],
    }
    path = os.path.join(ROOT, "profiles", "ncu_kernels.json")
    db = json.load(open(path)) if os.path.exists(path) else {}
    db[f"config{cfg}"] = rec
    with open(path, "w") as fh:
        json.dump(db, fh, indent=1)
    print(json.dumps({k: v for k, v in rec.items() if k != "launches"}, indent=1))


def unit(u):
    u = u.lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)


if __name__ == "__main__":
    main()
