"""Multi-GPU layer: one process per GPU, owner-computes patches, halo sum.

The reference has no working distributed path (`FluxEqlbSE.py:164` TODO; loops run over
`index_map(0)->size_local()` owned nodes and ghost contributions are never reduced,
SURVEY 5).  Here the mesh is partitioned by vertices; a rank holds all cells of its owned
vertices (owned cells + one layer of halo cells), equilibrates the patches of its owned
vertices only (`eqlb_mesh.node_owned`) and afterwards the partial sums on DOFs that also
live on other ranks are exchanged point-to-point and added in ascending rank order, which
makes the result independent of message arrival order: `P2PHaloExchange` (GPUs: one kernel
per rank over NVLink peer memory, `csrc/halo_p2p.cu`) or `HaloExchange` (`torch.distributed`
send/recv: NCCL on GPUs, gloo on CPU for the tests); both give the same bits.

Patches never talk to each other, so this is the only exchange step of the path; message
size is O(interface cells x ndofs x 8 B) (about 1 MB for a 1024-wide strip), i.e. latency
bound - it is issued once per equilibration call.
"""

from __future__ import annotations

import numpy as np

from .mesh import Mesh, build_topology, build_topology_fast


# --------------------------------------------------------------------------
# partitioning
# --------------------------------------------------------------------------


class LocalPart:
    """A rank's share of a mesh: local mesh (order-preserving renumbering, so facet
    orientations agree across ranks), ownership mask and global ids."""

    def __init__(self, mesh: Mesh, node_owned, node_gid, cell_gid):
        self.mesh = mesh
        self.node_owned = np.ascontiguousarray(node_owned, dtype=np.uint8)
        self.node_gid = np.asarray(node_gid, dtype=np.int64)
        self.cell_gid = np.asarray(cell_gid, dtype=np.int64)
        # cells that can also live on another rank: cells touching a non-owned vertex
        self.shared_cells = np.nonzero((self.node_owned[mesh.cell_node] == 0).any(axis=1))[0]

    def fct_gid(self, nnode_global):
        g = self.node_gid[self.mesh.fct_node]
        return g[:, 0] * np.int64(nnode_global) + g[:, 1]


def extract_local(mesh: Mesh, node_owner: np.ndarray, rank: int) -> LocalPart:
    """Cells touching a vertex owned by `rank`, renumbered order-preservingly."""
    owned_v = node_owner == rank
    cells = np.nonzero(owned_v[mesh.cell_node].any(axis=1))[0]
    verts = np.unique(mesh.cell_node[cells])  # sorted -> order preserving
    loc = np.full(mesh.nnode, -1, dtype=np.int64)
    loc[verts] = np.arange(verts.shape[0])
    lm = build_topology(mesh.x[verts, :2], loc[mesh.cell_node[cells]].astype(np.int32))
    # boundary ids of the global mesh (cut lines get 0)
    key_g = mesh.fct_node[mesh.bfct, 0].astype(np.int64) * mesh.nnode + mesh.fct_node[mesh.bfct, 1]
    side_of = dict(zip(key_g.tolist(), mesh.bfct_side.tolist()))
    gl = verts[lm.fct_node[lm.bfct]]
    key_l = gl[:, 0].astype(np.int64) * mesh.nnode + gl[:, 1]
    lm.bfct_side = np.array([side_of.get(int(kk), 0) for kk in key_l], dtype=np.int32)
    return LocalPart(lm, owned_v[verts], verts, cells)


def strip_owner(mesh: Mesh, world: int) -> np.ndarray:
    """Vertex owner by horizontal strips of the unit square."""
    return np.minimum((mesh.x[:, 1] * world).astype(np.int64), world - 1)


def crossed_strip(n: int, rank: int, world: int) -> tuple[LocalPart, int]:
    """Rank-local part of the weak-scaling benchmark mesh: `world` stacked n x n crossed
    blocks (n columns, world*n rows of squares).  Rank r owns grid-node rows
    [r n, (r+1) n) (the last rank also the top row) and the square centres of its rows; it
    holds one extra row of squares below as halo.  Returns (part, nnode_global)."""
    ny = world * n
    j0 = rank * n - (1 if rank > 0 else 0)
    j1 = (rank + 1) * n
    nr = j1 - j0
    ii, jj = np.meshgrid(np.arange(n + 1), np.arange(j0, j1 + 1), indexing="xy")
    xg = np.stack([ii.ravel() / n, jj.ravel() / n], axis=1)
    gid_g = (jj.ravel() * (n + 1) + ii.ravel()).astype(np.int64)
    ic, jc = np.meshgrid(np.arange(n), np.arange(j0, j1), indexing="xy")
    xc = np.stack([(ic.ravel() + 0.5) / n, (jc.ravel() + 0.5) / n], axis=1)
    ngrid_global = (n + 1) * (ny + 1)
    gid_c = (ngrid_global + jc.ravel() * n + ic.ravel()).astype(np.int64)
    x = np.concatenate([xg, xc])
    node_gid = np.concatenate([gid_g, gid_c])
    sq_i, sq_jl = ic.ravel(), jc.ravel() - j0
    v00 = sq_jl * (n + 1) + sq_i
    v10, v01 = v00 + 1, v00 + (n + 1)
    v11 = v01 + 1
    c = (n + 1) * (nr + 1) + sq_jl * n + sq_i
    tris = np.stack(
        [np.stack([v00, v10, c], 1), np.stack([v10, v11, c], 1), np.stack([v01, v11, c], 1), np.stack([v00, v01, c], 1)], axis=1
    ).reshape(-1, 3)
    m = build_topology(x, tris)
    mid = 0.5 * (m.x[m.fct_node[m.bfct, 0]] + m.x[m.fct_node[m.bfct, 1]])
    side = np.zeros(m.bfct.shape[0], dtype=np.int32)
    side[np.isclose(mid[:, 0], 0.0)] = 1
    side[np.isclose(mid[:, 1], 0.0)] = 2
    side[np.isclose(mid[:, 0], 1.0)] = 3
    side[np.isclose(mid[:, 1], float(world))] = 4
    m.bfct_side = side
    row_g = jj.ravel()
    own_g = (row_g >= rank * n) & ((row_g < (rank + 1) * n) | ((rank == world - 1) & (row_g == ny)))
    own_c = jc.ravel() >= rank * n
    cell_gid = (np.repeat(jc.ravel() * n + ic.ravel(), 4) * 4 + np.tile(np.arange(4), n * nr)).astype(np.int64)
    return LocalPart(m, np.concatenate([own_g, own_c]), node_gid, cell_gid), ngrid_global + n * ny


def crossed_rows(n: int, rank: int, world: int, fast: bool = True) -> tuple[LocalPart, int]:
    """Rank-local part of the STRONG-scaling meshes (BASELINE configs 3-5): the n x n crossed unit
    square cut into `world` horizontal strips of rows of squares.  Rank r owns the grid-node rows
    [r0, r1) of its strip (the last rank also the top row y = 1) and the centres of its squares; it
    holds one extra row of squares below as halo.  Node/cell numbering and geometry are those of
    `mesh.crossed_unit_square(n)` restricted to the strip (order preserving), boundary ids 1..4 of
    the global square.  Returns (part, nnode_global)."""
    rows = [(n * r) // world for r in range(world + 1)]
    j0 = rows[rank] - (1 if rank > 0 else 0)
    j1 = rows[rank + 1]
    nr = j1 - j0
    ii, jj = np.meshgrid(np.arange(n + 1), np.arange(j0, j1 + 1), indexing="xy")
    xg = np.stack([ii.ravel() / n, jj.ravel() / n], axis=1)
    gid_g = (jj.ravel() * (n + 1) + ii.ravel()).astype(np.int64)
    ic, jc = np.meshgrid(np.arange(n), np.arange(j0, j1), indexing="xy")
    xc = np.stack([(ic.ravel() + 0.5) / n, (jc.ravel() + 0.5) / n], axis=1)
    ngrid_global = (n + 1) * (n + 1)
    gid_c = (ngrid_global + jc.ravel() * n + ic.ravel()).astype(np.int64)
    x = np.concatenate([xg, xc])
    node_gid = np.concatenate([gid_g, gid_c])
    sq_i, sq_jl = ic.ravel(), jc.ravel() - j0
    v00 = sq_jl * (n + 1) + sq_i
    v10, v01 = v00 + 1, v00 + (n + 1)
    v11 = v01 + 1
    c = (n + 1) * (nr + 1) + sq_jl * n + sq_i
    tris = np.stack(
        [np.stack([v00, v10, c], 1), np.stack([v10, v11, c], 1), np.stack([v01, v11, c], 1), np.stack([v00, v01, c], 1)], axis=1
    ).reshape(-1, 3).astype(np.int32)
    m = (build_topology_fast if fast else build_topology)(x, tris)
    # boundary ids: the cut lines (rank > 0: bottom row of the halo squares) are not boundaries
    mid = 0.5 * (m.x[m.fct_node[m.bfct, 0]] + m.x[m.fct_node[m.bfct, 1]])
    side = np.zeros(m.bfct.shape[0], dtype=np.int32)
    side[np.isclose(mid[:, 0], 0.0)] = 1
    side[np.isclose(mid[:, 1], 0.0)] = 2
    side[np.isclose(mid[:, 0], 1.0)] = 3
    side[np.isclose(mid[:, 1], 1.0)] = 4
    m.bfct_side = side
    row_g = jj.ravel()
    own_g = (row_g >= rows[rank]) & ((row_g < rows[rank + 1]) | ((rank == world - 1) & (row_g == n)))
    own_c = jc.ravel() >= rows[rank]
    cell_gid = (np.repeat(jc.ravel() * n + ic.ravel(), 4) * 4 + np.tile(np.arange(4), n * nr)).astype(np.int64)
    return LocalPart(m, np.concatenate([own_g, own_c]), node_gid, cell_gid), ngrid_global + n * n


# --------------------------------------------------------------------------
# halo sum
# --------------------------------------------------------------------------


def se_dof_gids(part: LocalPart, nrt: int):
    """(local dof indices, global ids) of the DRT dofs that may live on other ranks."""
    c = part.shared_cells
    loc = (c[:, None] * nrt + np.arange(nrt)[None, :]).ravel()
    gid = (part.cell_gid[c][:, None] * nrt + np.arange(nrt)[None, :]).ravel()
    return loc.astype(np.int64), gid.astype(np.int64)


def ev_dof_gids(part: LocalPart, k: int, nnode_global: int):
    """Same for the conforming RT vector [fct*k+j][nfct*k + cell*(k*k-k) + i]."""
    m = part.mesh
    c = part.shared_cells
    f = np.unique(m.cell_fct[c])
    fg = part.fct_gid(nnode_global)
    ncd = k * k - k
    loc = [(f[:, None] * k + np.arange(k)[None, :]).ravel()]
    gid = [(fg[f][:, None] * k + np.arange(k)[None, :]).ravel()]
    if ncd:
        base = np.int64(nnode_global) ** 2 * k  # cell dofs after all possible facet keys
        loc.append((m.nfct * k + c[:, None] * ncd + np.arange(ncd)[None, :]).ravel())
        gid.append((base + part.cell_gid[c][:, None] * ncd + np.arange(ncd)[None, :]).ravel())
    return np.concatenate(loc).astype(np.int64), np.concatenate(gid).astype(np.int64)


class HaloExchange:
    """Sum of partial DOF values that live on several ranks.

    Setup (collective, once): ranks all-gather the global ids of their candidate dofs and
    intersect them pairwise; both partners order the shared set by global id.  `apply(x)`
    (collective, per call): pack -> batched isend/irecv -> add in ascending rank order."""

    def __init__(self, loc_idx, gid, device="cpu", group=None):
        import torch
        import torch.distributed as dist

        self.dist, self.torch, self.group = dist, torch, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        order = np.argsort(gid, kind="stable")
        gid_sorted, loc_sorted = gid[order], loc_idx[order]
        gathered = [None] * self.world
        dist.all_gather_object(gathered, gid_sorted, group=group)
        self.neigh = []
        for q in range(self.world):
            if q == self.rank:
                continue
            common, ia, _ = np.intersect1d(gid_sorted, gathered[q], assume_unique=True, return_indices=True)
            if common.size:
                idx = torch.as_tensor(loc_sorted[ia], dtype=torch.int64, device=device)
                self.neigh.append((q, idx))
        self.bytes_per_apply = sum(int(idx.numel()) * 8 for _, idx in self.neigh)

    def apply(self, xs):
        """xs: list of 1-D float64 tensors (one per RHS), updated in place."""
        torch, dist = self.torch, self.dist
        if not self.neigh:
            return
        ops, recvs = [], []
        for q, idx in self.neigh:
            send = torch.stack([x[idx] for x in xs]).contiguous()
            recv = torch.empty_like(send)
            ops.append(dist.P2POp(dist.isend, send, q, group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv, q, group=self.group))
            recvs.append(recv)
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for (q, idx), recv in zip(self.neigh, recvs):  # ascending rank order: deterministic
            for i, x in enumerate(xs):
                x.index_add_(0, idx, recv[i])


class P2PHaloExchange:
    """Same contract as `HaloExchange` on CUDA tensors, but the data path is one kernel per
    rank over NVLink peer memory (`csrc/halo_p2p.cu`): pack -> flag in the neighbour's memory
    -> add the neighbours' values, neighbours in ascending rank order (deterministic, the
    same bits as `HaloExchange`).  `torch.distributed` is only used once, to exchange the CUDA
    IPC handles of the communication buffers."""

    def __init__(self, loc_idx, gid, nrhs_max=1, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import cabi

        self.torch, self.C, self.cabi = torch, C, cabi
        self.lib = cabi.load_library()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        order = np.argsort(gid, kind="stable")
        gid_sorted, loc_sorted = gid[order], loc_idx[order]
        gathered = [None] * world
        dist.all_gather_object(gathered, gid_sorted, group=group)
        self.neigh = []
        for q in range(world):
            if q == rank:
                continue
            common, ia, _ = np.intersect1d(gid_sorted, gathered[q], assume_unique=True, return_indices=True)
            if common.size:
                self.neigh.append((q, np.ascontiguousarray(loc_sorted[ia], dtype=np.int64)))
        nn = len(self.neigh)
        counts = (C.c_int64 * max(nn, 1))(*[idx.size for _, idx in self.neigh])
        c_int64_p = C.POINTER(C.c_int64)
        idxp = (c_int64_p * max(nn, 1))(*[idx.ctypes.data_as(c_int64_p) for _, idx in self.neigh])
        handle = (C.c_ubyte * 64)()
        send_off = (C.c_int64 * max(nn, 1))()
        self.h = C.c_void_p()
        # every step that can fail locally (IPC export / mapping) is followed by a collective vote, so
        # that all ranks raise together and the caller can fall back to `HaloExchange` without a hang
        rc = self.lib.eqlb_halo_create(nn, counts, idxp, int(nrhs_max), C.byref(self.h), handle, send_off)
        err = self.lib.eqlb_last_error().decode() if rc != 0 else ""
        mine = {"ok": rc == 0, "err": err, "handle": bytes(handle), "neigh": [q for q, _ in self.neigh],
                "send_off": [int(v) for v in send_off[:nn]]}
        infos = [None] * world
        dist.all_gather_object(infos, mine, group=group)
        if not all(i["ok"] for i in infos):
            raise RuntimeError("P2PHaloExchange: " + "; ".join(i["err"] for i in infos if not i["ok"]))
        ok, err = True, ""
        for n, (q, _) in enumerate(self.neigh):
            peer = infos[q]
            slot = peer["neigh"].index(rank)
            ph = (C.c_ubyte * 64).from_buffer_copy(peer["handle"])
            if self.lib.eqlb_halo_connect(self.h, n, ph, peer["send_off"][slot], slot) != 0:
                ok, err = False, self.lib.eqlb_last_error().decode()
        votes = [None] * world
        dist.all_gather_object(votes, (ok, err), group=group)
        if not all(v[0] for v in votes):
            raise RuntimeError("P2PHaloExchange: " + "; ".join(v[1] for v in votes if not v[0]))
        self.bytes_per_apply = sum(int(idx.size) * 8 for _, idx in self.neigh)
        dist.barrier(group=group)

    def apply(self, xs):
        """xs: list of 1-D float64 CUDA tensors (one per RHS), updated in place on the current stream."""
        C, cabi = self.C, self.cabi
        arr = (cabi.c_double_p * len(xs))(*[C.cast(x.data_ptr(), cabi.c_double_p) for x in xs])
        rc = self.lib.eqlb_halo_apply(self.h, arr, len(xs), C.c_void_p(self.torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError(self.lib.eqlb_last_error().decode())

    def status(self):
        """Synchronise and raise if a device-side wait of an exchange timed out (`eqlb_halo_status`)."""
        rc = self.lib.eqlb_halo_status(self.h, self.C.c_void_p(self.torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError(self.lib.eqlb_last_error().decode())

    def close(self, group=None):
        """Collective: all ranks finish their exchanges (device synchronisation + barrier) before any rank frees
        its communication buffer - the neighbours read it over NVLink."""
        import torch.distributed as dist

        if getattr(self, "h", None):
            self.torch.cuda.synchronize()
            dist.barrier(group=group)
            self.lib.eqlb_halo_destroy(self.h)
            self.h = None
            dist.barrier(group=group)

    def __del__(self):
        # no collectives in a finaliser: `close()` is the orderly way; this only avoids a leak
        try:
            if getattr(self, "h", None):
                self.torch.cuda.synchronize()
                self.lib.eqlb_halo_destroy(self.h)
                self.h = None
        except Exception:
            pass


class _RawCuda:
    """a device pointer as an object torch can wrap without copying (`__cuda_array_interface__`)"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def staged_flux_tensors(problem, is_ev: bool):
    """CUDA tensors aliasing the device copies of the flux vectors written by the last host-buffer call of the
    handle (`eqlb_get_staged_flux`); valid until the next call."""
    import ctypes as C

    import torch

    out = []
    for r in range(problem.nrhs):
        ptr, n = C.c_void_p(), C.c_int64()
        if problem.lib.eqlb_get_staged_flux(problem.h, r, 1 if is_ev else 0, C.byref(ptr), C.byref(n)) != 0:
            raise RuntimeError(problem.lib.eqlb_last_error().decode())
        out.append(torch.as_tensor(_RawCuda(ptr.value, n.value), device="cuda"))
    return out


class HostHaloUpdate:
    """Host-buffer equilibration on a partitioned mesh without holding back the copy-out: the staged call writes the
    (not yet summed) result to the host vectors stage by stage while later stages are still computed; the halo sum
    runs on the device copy, and only the DOFs shared with other ranks - a few rows of cells next to the partition
    boundary - are fetched again.  `shared` = union of the exchange's index lists."""

    def __init__(self, hx):
        import torch

        idx = [np.asarray(i.cpu() if hasattr(i, "cpu") else i, dtype=np.int64) for _, i in hx.neigh]
        self.shared = np.unique(np.concatenate(idx)) if idx else np.zeros(0, np.int64)
        self.d_shared = torch.as_tensor(self.shared, device="cuda")
        self.hx = hx
        self._pin = None

    def finish(self, problem, is_ev: bool, host_flux):
        """after the host-buffer call: halo sum on the staged device copy, refresh of the shared DOFs in the host
        vectors (numpy arrays or CPU tensors)"""
        import torch

        dS = staged_flux_tensors(problem, is_ev)
        self.hx.apply(dS)
        if self.shared.size == 0:
            torch.cuda.current_stream().synchronize()
            return
        vals = torch.stack([d[self.d_shared] for d in dS])
        if self._pin is None or self._pin.shape != vals.shape:
            self._pin = torch.empty(vals.shape, dtype=torch.float64).pin_memory()
        self._pin.copy_(vals, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        for r, hflux in enumerate(host_flux):
            h_np = hflux.numpy() if hasattr(hflux, "numpy") else hflux
            h_np[self.shared] = self._pin[r].numpy()
