"""world_size-2/3 gloo tests (CPU) of the multi-GPU host logic: vertex partition, owned-
patch subsets and the halo sum.  The per-rank patch computation is done by the oracle
here (no GPU in this container); the partition/exchange code is the product code of
dolfinx_eqlb_b200/dist.py that bench.py runs over NCCL."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import PoissonCase, make_mesh
from dolfinx_eqlb_b200 import dist as dd, mesh as ms, tables as tb


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _facet_types(m):
    ft = np.zeros(m.nfct, dtype=np.int8)
    ft[m.bfct[m.bfct_side > 0]] = 1
    return ft


def _worker_generic(rank, world, port, path, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po

    m = make_mesh("crossed", 4, 3, perturb=0.2)
    T = tb.make_tables(k)
    rng = np.random.default_rng(5)
    G = rng.standard_normal(m.ncell * T.ndg * 2)
    F = rng.standard_normal(m.ncell * T.ndg)
    part = dd.extract_local(m, dd.strip_owner(m, world), rank)
    lm = part.mesh
    Gl = G.reshape(m.ncell, -1)[part.cell_gid].ravel()
    Fl = F.reshape(m.ncell, -1)[part.cell_gid].ravel()
    bc = po.BCData(_facet_types(lm)[None, :])
    if path == "se":
        x = po.se_run(lm, T, bc, [Gl], [Fl], node_owned=part.node_owned)[0]
        loc, gid = dd.se_dof_gids(part, T.nrt)
    else:
        x = po.ev_run(lm, T, bc, [Gl], [Fl], node_owned=part.node_owned)[0]
        loc, gid = dd.ev_dof_gids(part, T.k, m.nnode)
    xt = torch.from_numpy(x)
    hx = dd.HaloExchange(loc, gid)
    hx.apply([xt])
    # compare with the serial result on every local dof
    ref_bc = po.BCData(ms.facet_types(m, [1, 2, 3, 4], [])[None, :])
    if path == "se":
        ref = po.se_run(m, T, ref_bc, [G], [F])[0].reshape(m.ncell, T.nrt)[part.cell_gid].ravel()
        err = np.abs(xt.numpy() - ref).max() / np.abs(ref).max()
    else:
        ref = po.ev_run(m, T, ref_bc, [G], [F])[0]
        # map local facets/cells to global
        key_g = m.fct_node[:, 0].astype(np.int64) * m.nnode + m.fct_node[:, 1]
        fl = np.searchsorted(key_g, part.fct_gid(m.nnode))
        kk, ncd = T.k, T.k * T.k - T.k
        ref_l = np.concatenate([ref[: m.nfct * kk].reshape(m.nfct, kk)[fl].ravel(),
                                ref[m.nfct * kk :].reshape(m.ncell, ncd)[part.cell_gid].ravel()])
        err = np.abs(xt.numpy() - ref_l).max() / np.abs(ref_l).max()
    out[rank] = (err, len(hx.neigh), int(part.node_owned.sum()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("path,k", [("se", 2), ("ev", 2), ("se", 1)])
def test_partitioned_equals_serial(world, path, k):
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_generic, args=(world, port, path, k, out), nprocs=world, join=True)
        res = dict(out)
    assert len(res) == world
    assert sum(r[2] for r in res.values()) == make_mesh("crossed", 4, 3).nnode  # every vertex owned once
    for r in res.values():
        assert r[0] < 1e-12
        assert r[1] >= 1


def test_crossed_strip_parts_tile_the_stacked_mesh():
    n, world = 3, 3
    parts = [dd.crossed_strip(n, r, world) for r in range(world)]
    nng = parts[0][1]
    owned = np.concatenate([p.node_gid[p.node_owned.astype(bool)] for p, _ in parts])
    assert np.array_equal(np.sort(owned), np.arange(nng))  # each global vertex owned exactly once
    cells = np.unique(np.concatenate([p.cell_gid for p, _ in parts]))
    assert cells.shape[0] == 4 * n * n * world
    for p, _ in parts:
        # order preserving renumbering => orientation bits agree with global ids
        g = p.node_gid[p.mesh.cell_node]
        for f in range(3):
            a, b = ms.FACET_VERTS[f]
            assert ((g[:, a] > g[:, b]) == (p.mesh.cell_node[:, a] > p.mesh.cell_node[:, b])).all()
        # patches of owned vertices are complete: interior owned vertices have nf == nc
        nc = np.diff(p.mesh.node_cell_off)
        nf = np.diff(p.mesh.node_fct_off)
        x = p.mesh.x
        interior = (x[:, 0] > 0) & (x[:, 0] < 1) & (x[:, 1] > 0) & (x[:, 1] < world)
        sel = p.node_owned.astype(bool) & interior
        assert (nc[sel] == nf[sel]).all()


def _worker_rows(rank, world, port, cfg, out):
    """strong-scaling partition of bench.py's configs 2-4 (`dist.crossed_rows`: the global n x n mesh cut into strips
    of rows) with the boundary data restricted to the local facets - the CPU twin of `bench.halo_check`"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from dolfinx_eqlb_b200 import eqlb
    from oracle import pyoracle as po

    C = bench.CONFIGS[cfg]
    k, nrhs, stress = C["k"], C["nrhs"], C["stress"]
    T = tb.make_tables(k)
    n = 4 * world
    gm = ms.crossed_unit_square(n)
    G, F = bench.synthetic_inputs(gm.ncell, T.ndg, nrhs, seed=11)
    gbf, gbc = bench.boundary_conditions(gm, T, nrhs, list(C["neumann"]), seed=13)
    gbd = eqlb.boundarydata(gbc, gm, T, gbf, stress)
    ref = po.se_run(gm, T, po.BCData(gbd.facet_type, gbd.bflux, gbd.local_fct_id, gbd.node_on_stress_bnd), G, F, stress=stress)
    part, _ = dd.crossed_rows(n, rank, world, fast=False)
    lm, cg = part.mesh, part.cell_gid
    lG = [g.reshape(gm.ncell, -1)[cg].ravel() for g in G]
    lF = [f.reshape(gm.ncell, -1)[cg].ravel() for f in F]
    gkey = gm.fct_node[:, 0].astype(np.int64) * gm.nnode + gm.fct_node[:, 1]
    l2g_f = np.searchsorted(gkey, part.fct_gid(gm.nnode))
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
    lbd = eqlb.boundarydata(lbc, lm, T, lbf, stress)
    got = po.se_run(lm, T, po.BCData(lbd.facet_type, lbd.bflux, lbd.local_fct_id, lbd.node_on_stress_bnd), lG, lF, stress=stress,
                    node_owned=part.node_owned)
    loc, gid = dd.se_dof_gids(part, T.nrt)
    xs = [torch.from_numpy(np.ascontiguousarray(s)) for s in got]
    hx = dd.HaloExchange(loc, gid)
    hx.apply(xs)
    live_c = part.node_owned[lm.cell_node].any(axis=1)
    err = 0.0
    for r in range(nrhs):
        want = ref[r].reshape(gm.ncell, T.nrt)[cg]
        d = np.abs(xs[r].numpy().reshape(lm.ncell, T.nrt) - want)[live_c]
        err = max(err, float(d.max() / np.abs(want).max()))
    out[rank] = err
    dist.destroy_process_group()


@pytest.mark.parametrize("cfg", [2, 3, 4])
def test_strong_scaling_partition_equals_serial(cfg):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_rows, args=(world, port, cfg, out), nprocs=world, join=True)
        res = dict(out)
    assert len(res) == world
    for e in res.values():
        assert e < 1e-12
