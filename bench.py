#!/usr/bin/env python
"""Benchmark of the equilibration hot path: equilibrated patches / second.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config 1|2|3|4] [--path ev|se] [--n 1024]

One "step" = one equilibrate_fluxes()-equivalent call (all patches of the mesh,
patch maps resident, inputs resident in HBM).  Workload at N=1 (default, `--config 1`):
BASELINE.json configs[1] (Poisson, FluxEqlbEV, flux degree 2, 1024x1024 crossed unit square,
pure Dirichlet, synthetic random coefficients; N > 1: weak scaling, one such block per GPU).
`--config 2..4` = BASELINE.json configs[2..4] at their stated sizes, STRONG-scaled over the
N GPUs (the global mesh is fixed, cut into N strips of rows):
  2  Poisson FluxEqlbSE, degree 3, 4096x4096, mixed flux BCs (sides 1,4 traction with random
     DG_2 data, sides 2,3 Dirichlet: `python/test/unit/testcase_general.py:251-272`)
  3  linear elasticity, 2 stress rows + weak symmetry, degree 2, 2048x2048, pure Dirichlet
  4  Biot: 2 stress rows + Darcy flux (3 RHS), weak symmetry, degree 2, 4096x4096
Prints ONE JSON line on rank 0.  At N > 1 the line carries `halo_check`: the halo-summed result
of a small partitioned problem of the same kind against the single-GPU result (max relative
error; the run fails above 1e-12).
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED = 20240611


def synthetic_inputs(ncell, ndg, nrhs, seed=SEED):
    """`G ~ +-2(U+0.1)`, `f ~ 2(U+0.1)` (testcase_general.py:118-131)."""
    rng = np.random.default_rng(seed)
    G, F = [], []
    for _ in range(nrhs):
        g = 2.0 * (rng.random(ncell * ndg * 2) + 0.1)
        g *= np.where(rng.random(ncell * ndg * 2) < 0.5, -1.0, 1.0)
        G.append(g)
        F.append(2.0 * (rng.random(ncell * ndg) + 0.1))
    return G, F


def alg_bytes_per_cell(k, ndg, nrt, nrhs, path):
    """Compulsory unique HBM traffic per cell: 48 B geometry +
    nrhs * (16 ndg [G] + 8 ndg [f] + 8 ndofs [sigma out])."""
    nout = nrt if path == "se" else (k * k - k) + 1.5 * k  # EV: facet dofs shared by 2 cells
    return 48 + nrhs * (16 * ndg + 8 * ndg + 8 * nout)


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.stamps, self.reasons, self.max_mhz = [], [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def _run_nvml(self):
        """NVML polling (about 1 ms per sample): enough samples even for a few-millisecond timed region."""
        import pynvml as nv

        nv.nvmlInit()
        try:
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {
                "hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap,
            }
            while not self._halt.is_set():
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.stamps.append(time.time())
                try:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for nm, b in bits.items():
                        if r & b:
                            self.reasons.add(nm)
                except Exception:
                    pass
                self._halt.wait(0.002)
        finally:
            nv.nvmlShutdown()

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.stamps.append(time.time())
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.02)

    def stop(self, t0=None, t1=None):
        """Median SM clock of the samples taken inside the timed region [t0, t1] (all samples if none fell inside)."""
        self._halt.set()
        self.join(timeout=5)
        inside = [v for v, t in zip(self.samples, self.stamps) if t0 is not None and t0 <= t <= t1]
        use = inside or self.samples
        med = float(np.median(use)) if use else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples_in_timed_region": len(inside), "samples": len(self.samples)}


CONFIGS = {
    # BASELINE.json configs[i]: path, k, nrhs, n, stress, traction sides
    1: dict(path="ev", k=2, nrhs=1, n=1024, stress=False, neumann=[]),
    2: dict(path="se", k=3, nrhs=1, n=4096, stress=False, neumann=[1, 4]),
    3: dict(path="se", k=2, nrhs=2, n=2048, stress=True, neumann=[]),
    4: dict(path="se", k=2, nrhs=3, n=4096, stress=True, neumann=[]),
}

KERNEL_SOURCES = ["patch_k2_kernel.cu", "patch_kw_kernel.cu", "patch_k1_kernel.cu", "se_kernel.cu"]  # the files that hold the patch kernels


def kernel_source_hash():
    """sha256 over the kernel sources: ncu numbers in profiles/ncu_kernels.json are only printed
    for the sources they were captured from."""
    import hashlib

    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "dolfinx_eqlb_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_record(key):
    """ncu evidence of the dominant kernel of a workload (dram bytes per launch, executed FP64 flop per
    patch) from profiles/ncu_kernels.json (written by tools/ncu_to_json.py from a `ncu --set full`
    capture of this very bench command); None if there is no capture or it is stale."""
    p = os.path.join(ROOT, "profiles", "ncu_kernels.json")
    if not os.path.exists(p):
        return None
    with open(p) as fh:
        db = json.load(fh)
    rec = db.get(key)
    if rec is None:
        return None
    if rec.get("kernel_source_hash") != kernel_source_hash():
        return {"stale": True, "capture": rec.get("capture"), "kernel_source_hash": rec.get("kernel_source_hash")}
    return rec


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def boundary_conditions(m, T, nrhs, neumann, seed=SEED + 1):
    """(list_bfct_prime, list_bcs): `neumann` sides carry a traction with random DG_{k-1} data
    (`2 (U + 0.1)`, testcase_general.py:263-272), the other sides are Dirichlet sides of the primal problem."""
    from dolfinx_eqlb_b200 import eqlb

    dsides = [s for s in (1, 2, 3, 4) if s not in neumann]
    bfct = [m.boundary_facets(dsides) for _ in range(nrhs)]
    bcs = []
    rng = np.random.default_rng(seed)
    for _ in range(nrhs):
        if neumann:
            fcts = m.boundary_facets(neumann)
            bcs.append([eqlb.fluxbc(fcts, 2.0 * (rng.random((fcts.shape[0], T.k)) + 0.1))] if fcts.size else [])
        else:
            bcs.append([])
    return bfct, bcs


def build_case(n, k, nrhs, neumann=(), fast=None):
    from dolfinx_eqlb_b200 import mesh as ms, tables as tb

    T = tb.make_tables(k)
    m = ms.crossed_unit_square(n, fast=(n > 1024) if fast is None else fast)
    G, F = synthetic_inputs(m.ncell, T.ndg, nrhs)
    bfct, bcs = boundary_conditions(m, T, nrhs, list(neumann))
    return m, T, G, F, bfct, bcs


def _oracle_worker(path, n, k, nrhs, warmup, steps, barrier, queue, stress=False, neumann=()):
    """One host process = one rank of the reference's MPI-parallel CPU path: it equilibrates
    its own n x n block (no communication: an upper bound of the reference's scaling)."""
    from oracle import pyoracle as po
    from dolfinx_eqlb_b200 import eqlb

    impl = cpu_impl()[1]  # the reference's own compiled sources (oracle/_ref) when present, else the oracle port
    m, T, G, F, bfct, bcs = build_case(n, k, nrhs, neumann, fast=False)
    bd = eqlb.boundarydata(bcs, m, T, bfct, stress)
    bc = po.BCData(bd.facet_type, bd.bflux, bd.local_fct_id, bd.node_on_stress_bnd)
    run = (lambda: impl.se_run(m, T, bc, G, F, stress=stress)) if path == "se" else (lambda: impl.ev_run(m, T, bc, G, F))
    for _ in range(warmup):
        run()
    barrier.wait()
    dts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        run()
        dts.append(time.perf_counter() - t0)
    queue.put((m.nnode, dts))


def cpu_impl():
    """("reference", pyref) when oracle/_ref/libeqlb_ref.so (the reference's own sources compiled unchanged
    against stand-in headers) is available, else ("port", pyoracle)."""
    from oracle import pyoracle as po, pyref as pr

    if os.environ.get("EQLB_CPU_IMPL", "") != "port" and pr.available():
        try:
            pr.lib()
            return "reference", pr
        except Exception:
            pass
    return "port", po


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def time_oracle(path, n, k, nrhs, steps, warmup=0, procs=None, stress=False, neumann=()):
    """CPU baseline: the oracle restatement (reference loop structure) on a bounded sample
    of the same workload, one process per host core like the reference under mpirun.
    Returns (patches/s over all processes, patches per process, seconds per step, processes)."""
    import multiprocessing as mp

    procs = procs or min(host_cores(), 128)
    ctx = mp.get_context("spawn")
    barrier, queue = ctx.Barrier(procs), ctx.Queue()
    ws = [ctx.Process(target=_oracle_worker, args=(path, n, k, nrhs, warmup, steps, barrier, queue, stress, tuple(neumann))) for _ in range(procs)]
    for w in ws:
        w.start()
    res = [queue.get() for _ in ws]
    for w in ws:
        w.join()
    npatch = res[0][0]
    # a step ends when the slowest process has finished it
    step_s = [max(r[1][i] for r in res) for i in range(steps)]
    sec = float(np.mean(step_s))
    return procs * npatch / sec, npatch, sec, procs


def reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on all host cores of the box:
    `oracle/_ref` (its sources compiled unchanged against stand-in DOLFINx/Basix/Eigen headers) when the
    library travelled with the snapshot, else the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = args.cpu_n
    t0 = time.perf_counter()
    value, npatch, sec, procs = time_oracle(args.path, n_sample, args.k, args.nrhs, args.steps, max(args.warmup, 0), stress=args.stress,
                                            neumann=args.neumann)
    total = time.perf_counter() - t0
    sample = (f"{procs} processes x crossed {n_sample}x{n_sample} ({npatch} patches each) of the {args.n}x{args.n} workload, "
              f"mean of {args.steps} steps, step = slowest process")
    line = {
        "impl": "reference", "metric": "equilibrated patches/sec", "value": value, "unit": "patches/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
        "higher_is_better": True, "scaling": scaling_mode(args), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": procs, "kind": cpu_impl()[0], "sample": sample},
        "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": total,
    }
    print(json.dumps(line))


def scaling_mode(args):
    """config 1: weak (one 1024^2 block per GPU); configs 2-4: strong (BASELINE sizes are global)."""
    return "weak" if args.config == 1 else "strong"


def workload_config(args):
    """Identical for every N and for both arms (the driver compares `config` across runs)."""
    name = {"ev": f"Poisson FluxEqlbEV, P{args.k} primal / RT{args.k} flux", "se": f"Poisson FluxEqlbSE, P{args.k} primal / RT{args.k} flux"}[args.path]
    if getattr(args, "stress", False):
        name = ("Linear elasticity stress equilibration with weak symmetry (FluxEqlbSE)" if args.nrhs == 2
                else "Poro-elasticity (Biot): stress rows + Darcy flux, weak symmetry (FluxEqlbSE)")
    bcs = "pure Dirichlet" if not args.neumann else f"mixed fluxbc (sides {args.neumann} traction with random DG_{args.k - 1} data, other sides Dirichlet)"
    size = f"{args.n}x{args.n} crossed unit square" + (" per GPU" if args.config == 1 else " (global)")
    return {
        "workload": f"BASELINE configs[{args.config}]: {name}, degree_flux={args.k}, {size}, {bcs}, nrhs={args.nrhs}",
        "path": args.path, "degree_flux": args.k, "n": args.n, "nrhs": args.nrhs, "bcs": bcs,
        "l2": "inputs larger than L2 (no flush needed)", "accumulation": "colour-ordered (deterministic)",
        "parallelism": "single GPU at N=1; N>1: vertex strips, owner-computes patches, halo sum over NVLink peer memory",
    }


def halo_check(args, rank, world, dist_mod):
    """Driver-visible multi-GPU correctness: a small problem of the benchmarked kind (same path, degree, RHS
    count, BC layout) is partitioned over the ranks, equilibrated, halo-summed - and compared with the
    single-GPU result of the whole mesh computed redundantly on every rank.  Returns the max relative error
    over all ranks and RHS."""
    import torch

    from dolfinx_eqlb_b200 import dist as dd, eqlb, mesh as ms, tables as tb

    k, nrhs = args.k, args.nrhs
    T = tb.make_tables(k)
    n = 8 * world
    gm = ms.crossed_unit_square(n)
    G, F = synthetic_inputs(gm.ncell, T.ndg, nrhs, seed=SEED + 7)
    gbf, gbc = boundary_conditions(gm, T, nrhs, args.neumann, seed=SEED + 9)
    cls = eqlb.FluxEqlbSE if args.path == "se" else eqlb.FluxEqlbEV
    kw = {"equilibrate_stress": True} if args.stress else {}
    geq = cls(k, gm, F, G, **kw)
    geq.set_boundary_conditions(gbf, gbc)
    geq.equilibrate_fluxes()
    part, _ = dd.crossed_rows(n, rank, world, fast=False)
    lm = part.mesh
    cg = part.cell_gid
    lG = [g.reshape(gm.ncell, -1)[cg].ravel() for g in G]
    lF = [f.reshape(gm.ncell, -1)[cg].ravel() for f in F]
    # local BCs: the global boundary data restricted to the local facets (same traction coefficients)
    gkey = gm.fct_node[:, 0].astype(np.int64) * gm.nnode + gm.fct_node[:, 1]
    lkey = part.fct_gid(gm.nnode)
    l2g_f = np.searchsorted(gkey, lkey)
    lbf, lbc = [], []
    for r in range(nrhs):
        gset = np.zeros(gm.nfct, dtype=bool)
        gset[gbf[r]] = True
        lbf.append(np.nonzero(gset[l2g_f])[0].astype(np.int32))
        bl = []
        for bc in gbc[r]:
            row = np.full(gm.nfct, -1, dtype=np.int64)
            row[bc.facets] = np.arange(bc.facets.shape[0])
            sel = np.nonzero(row[l2g_f] >= 0)[0].astype(np.int32)
            if sel.size:
                bl.append(eqlb.fluxbc(sel, bc.coeffs[row[l2g_f[sel]]]))
        lbc.append(bl)
    leq = cls(k, lm, lF, lG, node_owned=part.node_owned, **kw)
    leq.set_boundary_conditions(lbf, lbc)
    leq.equilibrate_fluxes()
    if args.path == "se":
        loc, gid = dd.se_dof_gids(part, T.nrt)
    else:
        loc, gid = dd.ev_dof_gids(part, k, gm.nnode)
    xs = [torch.from_numpy(np.ascontiguousarray(s)).cuda() for s in leq.list_flux]
    try:
        hx = dd.P2PHaloExchange(loc, gid, nrhs_max=nrhs) if os.environ.get("EQLB_HALO", "p2p") != "nccl" else dd.HaloExchange(loc, gid, device="cuda")
    except RuntimeError:
        hx = dd.HaloExchange(loc, gid, device="cuda")
    hx.apply(xs)
    torch.cuda.synchronize()
    # the host-buffer route of the end-to-end leg: halo sum on the handle's staged device copy, shared DOFs refreshed
    # in the host vectors (dist.HostHaloUpdate)
    dd.HostHaloUpdate(hx).finish(leq.problem, args.path == "ev", leq.list_flux)
    # compared: the DOFs of all local cells with at least one owned vertex (a rank's copy of the other halo
    # cells - e.g. the bottom triangles of the halo row - receives no contribution and is never read)
    live_c = part.node_owned[lm.cell_node].any(axis=1)
    live_f = np.zeros(lm.nfct, dtype=bool)
    live_f[lm.cell_fct[live_c].ravel()] = True
    err = 0.0
    for r in range(nrhs):
        ref = geq.list_flux[r]
        for got in (xs[r].cpu().numpy(), np.asarray(leq.list_flux[r])):
            if args.path == "se":
                want = ref.reshape(gm.ncell, T.nrt)[cg]
                d = np.abs(got.reshape(lm.ncell, T.nrt) - want)[live_c]
            else:
                ncd = k * k - k
                wf = ref[: gm.nfct * k].reshape(gm.nfct, k)[l2g_f]
                wc = ref[gm.nfct * k :].reshape(gm.ncell, ncd)[cg]
                want = np.concatenate([wf.ravel(), wc.ravel()])
                d = np.concatenate([np.abs(got[: lm.nfct * k].reshape(lm.nfct, k) - wf)[live_f].ravel(),
                                    np.abs(got[lm.nfct * k :].reshape(lm.ncell, ncd) - wc)[live_c].ravel()])
            err = max(err, float(d.max() / max(np.abs(want).max(), 1e-300)))
    t = torch.tensor([err], device="cuda", dtype=torch.float64)
    dist_mod.all_reduce(t, op=dist_mod.ReduceOp.MAX)
    if hasattr(hx, "close"):
        hx.close()
    del hx
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=int(os.environ.get("EQLB_BENCH_CONFIG", "1")), choices=[1, 2, 3, 4],
                    help="BASELINE.json configs[i]; --path/--k/--nrhs/--n/--stress override single fields")
    ap.add_argument("--path", default=None, choices=["se", "ev"])
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--nrhs", type=int, default=None)
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--neumann", type=str, default=None, help="comma separated traction sides, e.g. 1,4")
    ap.add_argument("--cpu-n", type=int, default=None, dest="cpu_n",
                    help="edge length of the CPU sample mesh per process (default: 256, shrunk so that the reference arm's steps+warmup fit ~90 s)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--stress", action="store_true", help="SE with weak symmetry: nrhs >= 2 rows of a stress tensor")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    cfg0 = CONFIGS[args.config]
    args.path = args.path or os.environ.get("EQLB_BENCH_PATH") or cfg0["path"]
    args.k = args.k or cfg0["k"]
    args.nrhs = args.nrhs or cfg0["nrhs"]
    args.n = args.n or cfg0["n"]
    args.stress = args.stress or cfg0["stress"]
    args.neumann = [int(s) for s in args.neumann.split(",") if s] if args.neumann is not None else list(cfg0["neumann"])
    if args.stress:
        args.path, args.nrhs = "se", max(args.nrhs, 2)

    if args.cpu_n is None:
        args.cpu_n = 256
        if args.impl == "reference" and int(os.environ.get("RANK", "0")) == 0:
            # calibrate on a 64 x 64 block (all processes busy), then size the sample so that the whole
            # run (steps + warm-up) takes about 90 s; the cost per step grows like n^2
            _, _, sec64, _ = time_oracle(args.path, 64, args.k, args.nrhs, 1, 0, stress=args.stress, neumann=args.neumann)
            budget = 90.0 / max(args.steps + args.warmup, 1)
            args.cpu_n = int(max(32, min(256, 64 * (budget / sec64) ** 0.5)) // 8 * 8)
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch

    from dolfinx_eqlb_b200 import cabi, eqlb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- problem setup (not timed): mesh, tables, device residency, patch maps ----
    k, nrhs = args.k, args.nrhs
    node_owned, part, nnode_global = None, None, None
    if world == 1:
        m, T, G, F, bfct, bcs = build_case(args.n, args.k, args.nrhs, args.neumann)
        npatch_total = m.nnode
    else:
        from dolfinx_eqlb_b200 import dist as dd, tables as tb

        T = tb.make_tables(k)
        if args.config == 1:
            # weak scaling: `world` stacked n x n blocks, strip-partitioned by vertex rows
            part, nnode_global = dd.crossed_strip(args.n, rank, world)
        else:
            # strong scaling: the n x n mesh of the config cut into `world` strips of rows
            part, nnode_global = dd.crossed_rows(args.n, rank, world)
        m, node_owned = part.mesh, part.node_owned
        G, F = synthetic_inputs(m.ncell, T.ndg, nrhs, seed=SEED + rank)
        bfct, bcs = boundary_conditions(m, T, nrhs, args.neumann, seed=SEED + 1 + rank)
        npatch_total = nnode_global
    if args.path == "se":
        eq = eqlb.FluxEqlbSE(k, m, F, G, node_owned=node_owned, host_pipeline=False, equilibrate_stress=args.stress,
                             interface_first=world > 1 and bool(os.environ.get("EQLB_OVERLAP")))
        nout = m.ncell * T.nrt
    else:
        eq = eqlb.FluxEqlbEV(k, m, F, G, node_owned=node_owned, host_pipeline=False, interface_first=world > 1 and bool(os.environ.get("EQLB_OVERLAP")))
        nout = eq.ndofs
    eq.set_boundary_conditions(bfct, bcs)
    prob = eq.problem
    lib = prob.lib
    stream = torch.cuda.current_stream()
    prob.set_stream(stream.cuda_stream)

    dG = [torch.from_numpy(g).cuda() for g in G]
    dF = [torch.from_numpy(f).cuda() for f in F]
    dS = [torch.zeros(nout, dtype=torch.float64, device="cuda") for _ in range(nrhs)]

    def dptrs(ts):
        arr = (cabi.c_double_p * len(ts))()
        for i, t in enumerate(ts):
            arr[i] = C.cast(t.data_ptr(), cabi.c_double_p)
        return arr

    pG, pF, pS = dptrs(dG), dptrs(dF), dptrs(dS)

    hx, halo_name = None, None
    if world > 1:
        if args.path == "se":
            loc, gid = dd.se_dof_gids(part, T.nrt)
        else:
            loc, gid = dd.ev_dof_gids(part, k, nnode_global)
        # halo sum: one kernel per rank over NVLink peer memory (EQLB_HALO=nccl: NCCL send/recv + index_add)
        if os.environ.get("EQLB_HALO", "p2p") == "nccl":
            hx = dd.HaloExchange(loc, gid, device="cuda")
        else:
            try:
                hx = dd.P2PHaloExchange(loc, gid, nrhs_max=nrhs)
            except RuntimeError as e:  # e.g. CUDA IPC not permitted on this box: all ranks raise together
                if rank == 0:
                    print(f"[bench] peer-memory halo unavailable ({e}); using NCCL send/recv", file=sys.stderr)
                hx = dd.HaloExchange(loc, gid, device="cuda")

    halo_name = type(hx).__name__ if hx is not None else None

    def run_device():
        if args.path == "se":
            rc = lib.eqlb_se_run(prob.h, pG, pF, pS, cabi.c_double_p(), 1)
        else:
            rc = lib.eqlb_ev_run(prob.h, pG, pF, pS, 1)
        if rc != 0:
            raise RuntimeError(lib.eqlb_last_error().decode())

    comm_stream = torch.cuda.Stream() if world > 1 else None
    ev_iface = torch.cuda.Event() if world > 1 else None
    # measured on B200: N=2 0.796 (overlap) vs 0.819 ms, N=8 0.865 (overlap) vs 0.825 ms - the NCCL kernels wait for
    # SMs behind the persistent patch kernel and the split costs 3 extra launches, so it is opt-in
    overlap = world > 1 and bool(os.environ.get("EQLB_OVERLAP"))

    def step_device():
        if hx is None:
            run_device()
        elif not overlap:
            run_device()
            hx.apply(dS)
        else:
            # interface patches first; their halo sum over NVLink (NCCL send/recv, the only exchange of
            # the path) runs on a second stream while the interior patches are computed
            prob.set_part(1)
            run_device()
            ev_iface.record(stream)
            prob.set_part(2)
            run_device()
            prob.set_part(0)
            comm_stream.wait_event(ev_iface)
            with torch.cuda.stream(comm_stream):
                hx.apply(dS)
            stream.wait_stream(comm_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    t_wall0 = time.time()
    l0 = prob.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    launches = prob.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, time.time())
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = npatch_total / (ms_step * 1e-3)

    # ---- end-to-end through the host API: pinned host buffers, H2D + D2H inside ----
    hG = [torch.from_numpy(g).pin_memory() for g in G]
    hF = [torch.from_numpy(f).pin_memory() for f in F]
    hS = [torch.zeros(nout, dtype=torch.float64).pin_memory() for _ in range(nrhs)]
    qG, qF, qS = dptrs(hG), dptrs(hF), dptrs(hS)
    # the user-facing default for host arrays: staged copy-in / kernels / copy-out
    cls = eqlb.FluxEqlbSE if args.path == "se" else eqlb.FluxEqlbEV
    kw = {"equilibrate_stress": True} if args.stress else {}
    heq = cls(k, m, F, G, node_owned=node_owned, host_pipeline=True, **kw)
    heq.set_boundary_conditions(bfct, bcs)
    hprob = heq.problem
    hprob.set_stream(stream.cuda_stream)

    def run_host(ps, memspace):
        if args.path == "se":
            rc = lib.eqlb_se_run(hprob.h, qG, qF, ps, cabi.c_double_p(), memspace)
        else:
            rc = lib.eqlb_ev_run(hprob.h, qG, qF, ps, memspace)
        if rc != 0:
            raise RuntimeError(lib.eqlb_last_error().decode())

    host_halo = dd.HostHaloUpdate(hx) if world > 1 else None

    def step_host():
        if world > 1:
            # staged call (copy-in / kernels / copy-out of finished ranges) as at N = 1, then the halo sum on the
            # device copy of the flux and a refresh of the few shared DOFs in the host vectors
            run_host(qS, 2)
            host_halo.finish(hprob, args.path == "ev", hS)
            return
        # EQLB_HOST_ZEROED: the flux starts from zero as in the reference's equilibrate_fluxes
        run_host(qS, 2)

    e2e_steps = max(3, min(args.steps, 5))
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = sum(g.numel() * 8 for g in hG) + sum(f.numel() * 8 for f in hF)
    d2h = sum(s.numel() * 8 for s in hS)

    # ---- fused call (SURVEY 8f ranks 2+3): the primal solution u_h and the data f_h (P_k vectors) go in, the
    # projection, the equilibration and the error estimator run on the device, the cell-wise indicators come back
    fused = None
    if world == 1 and not args.stress and k <= 3:
        from dolfinx_eqlb_b200 import mesh as ms_

        dm, npk_dofs = ms_.pk_dofmap(m, k)
        hprob.set_primal_space(dm, npk_dofs)
        rngf = np.random.default_rng(SEED + 3)
        hU = [torch.from_numpy(rngf.standard_normal(npk_dofs)).pin_memory() for _ in range(nrhs)]
        hFh = [torch.from_numpy(rngf.standard_normal(npk_dofs)).pin_memory() for _ in range(nrhs)]
        hE1 = [torch.zeros(m.ncell, dtype=torch.float64).pin_memory() for _ in range(nrhs)]
        hE2 = [torch.zeros(m.ncell, dtype=torch.float64).pin_memory() for _ in range(nrhs)]
        dU = [torch.empty(npk_dofs, dtype=torch.float64, device="cuda") for _ in range(nrhs)]
        dFh = [torch.empty(npk_dofs, dtype=torch.float64, device="cuda") for _ in range(nrhs)]
        dE1 = [torch.empty(m.ncell, dtype=torch.float64, device="cuda") for _ in range(nrhs)]
        dE2 = [torch.empty(m.ncell, dtype=torch.float64, device="cuda") for _ in range(nrhs)]
        pU, pFh, pE1, pE2 = dptrs(dU), dptrs(dFh), dptrs(dE1), dptrs(dE2)
        is_ev = args.path == "ev"

        def step_fused():
            for d, h_ in zip(dU + dFh, hU + hFh):
                d.copy_(h_, non_blocking=True)
            for d in dS:
                d.zero_()
            if is_ev:
                rc = lib.eqlb_ev_run_primal(hprob.h, pU, pFh, pS, 1)
            else:
                rc = lib.eqlb_se_run_primal(hprob.h, pU, pFh, pS, cabi.c_double_p(), 1)
            if rc == 0:
                rc = lib.eqlb_estimate_poisson(hprob.h, nrhs, pS, pU, pFh, pE1, pE2, int(is_ev), 1)
            if rc != 0:
                raise RuntimeError(lib.eqlb_last_error().decode())
            for d, h_ in zip(dE1 + dE2, hE1 + hE2):
                h_.copy_(d, non_blocking=True)
            torch.cuda.synchronize()

        step_fused()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_fused()
        dtf = (time.perf_counter() - t0) / e2e_steps
        fused = {"value": npatch_total / dtf, "unit": "patches/s", "seconds": dtf,
                 "h2d_bytes_per_step": sum(t.numel() * 8 for t in hU + hFh), "d2h_bytes_per_step": sum(t.numel() * 8 for t in hE1 + hE2),
                 "what": "u_h, f_h (P_k vectors, pinned host) -> device: projection + equilibration + Poisson error estimator -> "
                         "cell-wise eta_sig^2, eta_osc^2 back to the host (the flux stays on the device)"}

    # ---- cold call: what one equilibration of a NEW mesh costs (adaptive loops equilibrate every mesh once):
    # eqlb_create (mesh upload, Jacobians, colouring) + eqlb_set_bcs (device patch builder) + one host-buffer call
    cold = None
    if world == 1:
        best = None
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ceq = cls(k, m, F, G, host_pipeline=True, **kw)
            ceq.set_boundary_conditions(bfct, bcs)
            ceq.equilibrate_fluxes()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            parts = (ceq.problem.create_seconds, ceq.problem.set_bcs_seconds)
            del ceq
            if best is None or dt < best[0]:
                best = (dt, parts)
        cold = {"value": npatch_total / best[0], "unit": "patches/s", "seconds": best[0], "eqlb_create_ms": 1e3 * best[1][0],
                "eqlb_set_bcs_ms": 1e3 * best[1][1],
                "includes": "equilibrator construction (eqlb_create) + set_boundary_conditions (eqlb_set_bcs) + one equilibrate_fluxes() "
                            "with host buffers; best of 2"}

    hcheck = None
    if dist is not None:
        if hasattr(hx, "status"):
            hx.status()
        if hasattr(hx, "close"):
            hx.close()
        del hx
        torch.cuda.synchronize()
        dist.barrier()
        hcheck = halo_check(args, rank, world, dist)
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        if hcheck is not None and not (hcheck <= 1e-12):
            sys.exit(3)
        return
    # ---- roofline of the dominant kernel (the patch kernel is the whole step) ----
    peak, peak_src = peaks()
    bpc = alg_bytes_per_cell(k, T.ndg, T.nrt, nrhs, args.path)
    alg_bytes_step = bpc * m.ncell  # per rank
    achieved = alg_bytes_step / (ms_step * 1e-3) / 1e9
    ncu_key = f"config{args.config}" if (args.n == CONFIGS[args.config]["n"] and args.path == CONFIGS[args.config]["path"]
                                         and k == CONFIGS[args.config]["k"]) else None
    ncu = ncu_record(ncu_key) if (ncu_key and world == 1) else None
    lps = launches / args.steps
    roof = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": ncu["traffic"] if (ncu and not ncu.get("stale")) else None,
        "kernel": (ncu or {}).get("kernel") or ("patch_k2w_kernel" if k == 2 else ("patch_kw_kernel" if k == 3 else "patch_kernel")),
        "alg_bytes_per_patch": bpc * m.ncell / m.nnode, "alg_bytes_per_launch": alg_bytes_step / lps,
        "avg_launch_ms": ms_step / lps, "peak_source": peak_src, "launches_per_step": lps,
        "kernel_source_hash": kernel_source_hash(),
    }
    if ncu and ncu.get("stale"):
        roof["ncu"] = "stale capture (kernel sources changed since profiles/ncu_kernels.json was written): traffic/flops withheld"
    # FP64 peak: measured in this very run (DFMA chains on all SMs) with its own clocks record
    fp = C.c_double(0.0)
    psam = ClockSampler(local_rank)
    psam.start()
    tp0 = time.time()
    rc_peak = lib.eqlb_measure_fp64_peak(20000, 5, C.byref(fp))
    pclk = psam.stop(tp0, time.time())
    if rc_peak == 0 and ncu and not ncu.get("stale") and ncu.get("flop_per_patch"):
        tf = ncu["flop_per_patch"] * m.nnode / (ms_step * 1e-3) / 1e12
        roof["fp64"] = {"achieved": tf, "peak": fp.value, "unit": "TFLOP/s", "frac": tf / fp.value,
                        "flop_per_patch": ncu["flop_per_patch"], "source": ncu["capture"], "peak_clocks": pclk,
                        "peak_source": "eqlb_measure_fp64_peak (DFMA chains, this run)",
                        "note": "executed FP64 flop (ncu) / step time; the kernel is FP64/issue bound, see DESIGN.md section 4"}
    elif rc_peak == 0:
        roof["fp64"] = {"peak": fp.value, "unit": "TFLOP/s", "peak_clocks": pclk, "peak_source": "eqlb_measure_fp64_peak (DFMA chains, this run)"}
    cpu = None
    if not args.no_cpu and world == 1:  # the CPU baseline is timed at N = 1 only
        v, npatch_s, dt, procs = time_oracle(args.path, args.cpu_n, k, nrhs, 3, 1, stress=args.stress, neumann=args.neumann)
        cpu = {"value": v, "unit": "patches/s", "cores": procs, "kind": cpu_impl()[0],
               "sample": f"{procs} processes x crossed {args.cpu_n}x{args.cpu_n} ({npatch_s} patches each) of the same workload, mean of 3 steps"}
    cfg = workload_config(args)
    line = {
        "metric": "equilibrated patches/sec", "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling_mode(args), "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
        "e2e": {"value": npatch_total / e2e_s, "unit": "patches/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "e2e_fused": fused, "e2e_cold": cold, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "halo_check": hcheck, "halo": halo_name,
        # one-time cost per (mesh, BC set), outside the timed region: the reference redoes this work in every call
        # (second handle of the process = the staged host-call one: CUDA context and allocator are warm)
        "setup": {"eqlb_create_ms": 1e3 * hprob.create_seconds, "eqlb_set_bcs_ms": 1e3 * hprob.set_bcs_seconds,
                  "first_handle_ms": 1e3 * (prob.create_seconds + prob.set_bcs_seconds)},
    }
    print(json.dumps(line))
    if hcheck is not None and not (hcheck <= 1e-12):
        print(f"[bench] halo_check {hcheck:.3e} > 1e-12", file=sys.stderr)
        sys.exit(3)


if __name__ == "__main__":
    main()
