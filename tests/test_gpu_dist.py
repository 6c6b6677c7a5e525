"""2-GPU NCCL test: strip-partitioned equilibration + halo sum equals the single-GPU
result (skipped on boxes with fewer than 2 GPUs)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path, k, out, split=False):
    import torch.distributed as dist

    from common import make_mesh
    from dolfinx_eqlb_b200 import dist as dd, eqlb, tables as tb

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m = make_mesh("crossed", 12, 3, perturb=0.2)
    T = tb.make_tables(k)
    rng = np.random.default_rng(5)
    G = rng.standard_normal(m.ncell * T.ndg * 2)
    F = rng.standard_normal(m.ncell * T.ndg)
    part = dd.extract_local(m, dd.strip_owner(m, world), rank)
    lm = part.mesh
    Gl = G.reshape(m.ncell, -1)[part.cell_gid].ravel()
    Fl = F.reshape(m.ncell, -1)[part.cell_gid].ravel()
    cls = eqlb.FluxEqlbSE if path == "se" else eqlb.FluxEqlbEV
    eq = cls(k, lm, [Fl], [Gl], node_owned=part.node_owned, host_pipeline=not split, interface_first=split)
    eq.set_boundary_conditions([lm.bfct[lm.bfct_side > 0].astype(np.int32)], [[]])
    if split:
        # interface patches and interior patches in two calls (eqlb_set_part): same total
        eq.problem.set_part(1)
        eq.equilibrate_fluxes()
        eq.problem.set_part(2)
        eq.equilibrate_fluxes()
        eq.problem.set_part(0)
    else:
        eq.equilibrate_fluxes()
    loc, gid = dd.se_dof_gids(part, T.nrt) if path == "se" else dd.ev_dof_gids(part, k, m.nnode)
    x = torch.from_numpy(eq.list_flux[0]).cuda()
    # halo sum over NVLink peer memory (one kernel per rank) == NCCL send/recv + index_add, bit
    # for bit, also when repeated (flag epochs, double buffering)
    y = [x.clone(), (0.5 * x).clone()]
    z = [t.clone() for t in y]
    nccl, p2p = dd.HaloExchange(loc, gid, device="cuda"), dd.P2PHaloExchange(loc, gid, nrhs_max=2)
    for _ in range(5):
        nccl.apply(y)
        p2p.apply(z)
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(y, z))
    dist.barrier()
    p2p.apply([x])
    torch.cuda.synchronize()
    # single-GPU reference of the whole mesh on this rank's device
    ref_eq = cls(k, m, [F], [G])
    ref_eq.set_boundary_conditions([m.boundary_facets([1, 2, 3, 4])], [[]])
    ref_eq.equilibrate_fluxes()
    ref = ref_eq.list_flux[0]
    if path == "se":
        ref_l = ref.reshape(m.ncell, T.nrt)[part.cell_gid].ravel()
    else:
        key_g = m.fct_node[:, 0].astype(np.int64) * m.nnode + m.fct_node[:, 1]
        fl = np.searchsorted(key_g, part.fct_gid(m.nnode))
        ncd = k * k - k
        ref_l = np.concatenate([ref[: m.nfct * k].reshape(m.nfct, k)[fl].ravel(), ref[m.nfct * k :].reshape(m.ncell, ncd)[part.cell_gid].ravel()])
    err = float(np.abs(x.cpu().numpy() - ref_l).max() / np.abs(ref_l).max())
    if not split:
        # host-buffer route (dist.HostHaloUpdate): halo sum on the handle's staged device copy of the flux, shared
        # DOFs refreshed in the host vector -> the same numbers as the device route, bit for bit
        dd.HostHaloUpdate(p2p).finish(eq.problem, path == "ev", eq.list_flux)
        if not np.array_equal(eq.list_flux[0], x.cpu().numpy()):
            err = max(err, 1.0)
    out[rank] = err
    dist.destroy_process_group()


@pytest.mark.parametrize("world,path,k,split", [(2, "se", 2, False), (2, "ev", 2, False), (2, "se", 3, False), (2, "ev", 2, True),
                                                 (2, "se", 1, True), (4, "ev", 2, False), (4, "se", 2, True)])
def test_multi_gpu_halo_sum(world, path, k, split):
    """world = 4: the middle strips have two neighbours (flag slots, per-neighbour barriers)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, path, k, out, split), nprocs=world, join=True)
        res = dict(out)
    assert len(res) == world
    for e in res.values():
        assert e < 1e-12


def test_set_part_single_gpu():
    """eqlb_set_part on one GPU with an artificial ownership mask: interface + interior
    patches in two calls give the one-call result; the interior part leaves the DOFs of
    cells with a foreign vertex untouched."""
    from common import PoissonCase, make_mesh
    from dolfinx_eqlb_b200 import eqlb

    m = make_mesh("crossed", 10, 2, perturb=0.2)
    case = PoissonCase(m, 2, [[]], seed=4, galerkin=False)
    owned = (m.x[:, 1] < 0.55).astype(np.uint8)
    bf = [m.boundary_facets([1, 2, 3, 4])]
    whole = eqlb.FluxEqlbSE(2, m, case.F, case.G, node_owned=owned, host_pipeline=False)
    whole.set_boundary_conditions(bf, [[]])
    whole.equilibrate_fluxes()
    eq = eqlb.FluxEqlbSE(2, m, case.F, case.G, node_owned=owned, host_pipeline=False, interface_first=True)
    eq.set_boundary_conditions(bf, [[]])
    eq.problem.set_part(2)
    eq.equilibrate_fluxes()
    nrt = case.T.nrt
    shared_cells = (owned[m.cell_node] == 0).any(axis=1)
    assert np.abs(eq.list_flux[0].reshape(m.ncell, nrt)[shared_cells]).max() == 0.0
    eq.problem.set_part(1)
    eq.equilibrate_fluxes()
    assert np.abs(eq.list_flux[0] - whole.list_flux[0]).max() < 1e-13 * np.abs(whole.list_flux[0]).max()
    # a handle without the flag refuses
    with pytest.raises(RuntimeError):
        whole.problem.set_part(1)
