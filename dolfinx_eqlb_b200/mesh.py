"""Synthetic triangle meshes + the topology arrays the hot path reads.

The reference reads these arrays from DOLFINx 0.6.0 (`se/Patch.cpp:23-28`,
`se/reconstruction.hpp:83-93`, `se/solve_patch_semiexplt.hpp:223-224`,
`ev/reconstruction.hpp:79-107`).  DOLFINx is not available offline, so this
module produces arrays with the same *meaning*, with a fixed, documented
numbering:

* facet f of a cell is opposite local vertex f; `cell_fct[c, f]`
* facets are numbered by lexicographically sorted (min, max) vertex pair
* `fct_cell`, `node_cell`, `node_fct` adjacency lists are ascending
* `fct_perms[c*3+f] = 1` iff the two local vertices of facet f appear in
  descending global order in cell c (DOLFINx facet reflection bit)
* boundary ids follow `python/test/unit/utils.py:74-79`:
  1: x=0, 2: y=0, 3: x=1, 4: y=1

Everything is vectorised numpy so that the 1024^2 crossed mesh (4.2 M cells)
builds in seconds.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

FACET_VERTS = np.array([[1, 2], [0, 2], [0, 1]], dtype=np.int32)


@dataclass
class Mesh:
    x: np.ndarray  # [nnode][3] f64
    cell_node: np.ndarray  # [ncell][3] i32 (also geometry dofmap)
    cell_fct: np.ndarray  # [ncell][3] i32
    fct_node: np.ndarray  # [nfct][2] i32 (min, max)
    fct_cell_off: np.ndarray  # [nfct+1] i32
    fct_cell: np.ndarray  # CSR data i32
    node_cell_off: np.ndarray  # [nnode+1] i32
    node_cell: np.ndarray
    node_fct_off: np.ndarray  # [nnode+1] i32
    node_fct: np.ndarray
    fct_perms: np.ndarray  # [ncell*3] u8
    cell_perm_info: np.ndarray  # [ncell] u32
    bfct: np.ndarray  # boundary facet ids (ascending)
    bfct_side: np.ndarray  # boundary id 1..4 (0 if not on the unit-square frame)

    @property
    def nnode(self):
        return self.x.shape[0]

    @property
    def ncell(self):
        return self.cell_node.shape[0]

    @property
    def nfct(self):
        return self.fct_node.shape[0]

    def boundary_facets(self, sides):
        sides = np.atleast_1d(np.asarray(sides))
        return self.bfct[np.isin(self.bfct_side, sides)].astype(np.int32)


def _csr(keys: np.ndarray, vals: np.ndarray, n: int):
    """CSR adjacency key -> sorted vals."""
    order = np.lexsort((vals, keys))
    counts = np.bincount(keys, minlength=n)
    off = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(counts, out=off[1:])
    return off, vals[order].astype(np.int32)


def build_topology(x: np.ndarray, cell_node: np.ndarray) -> Mesh:
    cell_node = np.ascontiguousarray(cell_node, dtype=np.int32)
    ncell = cell_node.shape[0]
    nnode = x.shape[0]
    x3 = np.zeros((nnode, 3))
    x3[:, : x.shape[1]] = x

    # edges of every cell, local facet f = vertices FACET_VERTS[f]
    ea = cell_node[:, FACET_VERTS[:, 0]].astype(np.int64)  # [ncell][3]
    eb = cell_node[:, FACET_VERTS[:, 1]].astype(np.int64)
    lo, hi = np.minimum(ea, eb), np.maximum(ea, eb)
    key = lo * nnode + hi
    ukey, inv = np.unique(key.ravel(), return_inverse=True)
    nfct = ukey.shape[0]
    cell_fct = inv.reshape(ncell, 3).astype(np.int32)
    fct_node = np.stack([ukey // nnode, ukey % nnode], axis=1).astype(np.int32)

    cells_rep = np.repeat(np.arange(ncell, dtype=np.int64), 3)
    fct_cell_off, fct_cell = _csr(cell_fct.ravel().astype(np.int64), cells_rep, nfct)
    node_cell_off, node_cell = _csr(cell_node.ravel().astype(np.int64), cells_rep, nnode)
    fn = fct_node.ravel().astype(np.int64)
    node_fct_off, node_fct = _csr(fn, np.repeat(np.arange(nfct, dtype=np.int64), 2), nnode)

    perms = (ea > eb).astype(np.uint8)  # [ncell][3]
    info = (perms[:, 0].astype(np.uint32) | (perms[:, 1].astype(np.uint32) << 1) | (perms[:, 2].astype(np.uint32) << 2))

    nc_per_f = np.diff(fct_cell_off)
    bfct = np.nonzero(nc_per_f == 1)[0].astype(np.int32)
    mid = 0.5 * (x3[fct_node[bfct, 0]] + x3[fct_node[bfct, 1]])
    side = np.zeros(bfct.shape[0], dtype=np.int32)
    side[np.isclose(mid[:, 0], 0.0)] = 1
    side[np.isclose(mid[:, 1], 0.0)] = 2
    side[np.isclose(mid[:, 0], 1.0)] = 3
    side[np.isclose(mid[:, 1], 1.0)] = 4

    return Mesh(
        x=x3,
        cell_node=cell_node,
        cell_fct=cell_fct,
        fct_node=fct_node,
        fct_cell_off=fct_cell_off,
        fct_cell=fct_cell,
        node_cell_off=node_cell_off,
        node_cell=node_cell,
        node_fct_off=node_fct_off,
        node_fct=node_fct,
        fct_perms=np.ascontiguousarray(perms.ravel()),
        cell_perm_info=info,
        bfct=bfct,
        bfct_side=side,
    )


def build_topology_fast(x: np.ndarray, cell_node: np.ndarray, device=None) -> Mesh:
    """`build_topology` with the sorts done by torch (on the GPU when one is present): the
    numpy version needs minutes and > 20 GB for the 4096^2 crossed meshes of BASELINE configs
    3 and 5 (67 M cells, 100 M facets).  Same arrays, bit for bit (tests/test_mesh.py)."""
    import torch

    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    cell_node = np.ascontiguousarray(cell_node, dtype=np.int32)
    ncell, nnode = cell_node.shape[0], x.shape[0]
    x3 = np.zeros((nnode, 3))
    x3[:, : x.shape[1]] = x
    cn = torch.from_numpy(cell_node).to(device).long()
    fv = torch.as_tensor(FACET_VERTS, device=device).long()
    ea, eb = cn[:, fv[:, 0]], cn[:, fv[:, 1]]
    key = torch.minimum(ea, eb) * nnode + torch.maximum(ea, eb)
    perms = (ea > eb).to(torch.uint8)
    del ea, eb
    ukey, inv = torch.unique(key.reshape(-1), return_inverse=True)
    del key
    nfct = int(ukey.shape[0])
    cell_fct = inv.reshape(ncell, 3).to(torch.int32)
    del inv
    fct_node = torch.stack([ukey // nnode, ukey % nnode], dim=1).to(torch.int32)
    del ukey

    def csr(keys, n, width):
        # vals = repeat(arange, width) is ascending in input order: a stable sort by key is the lexsort
        order = torch.sort(keys, stable=True).indices
        counts = torch.bincount(keys, minlength=n)
        off = torch.zeros(n + 1, dtype=torch.int64, device=device)
        torch.cumsum(counts, 0, out=off[1:])
        return off.to(torch.int32).cpu().numpy(), (order // width).to(torch.int32).cpu().numpy()

    fct_cell_off, fct_cell = csr(cell_fct.reshape(-1).long(), nfct, 3)
    node_cell_off, node_cell = csr(cn.reshape(-1), nnode, 3)
    node_fct_off, node_fct = csr(fct_node.reshape(-1).long(), nnode, 2)
    info = (perms[:, 0].to(torch.int64) | (perms[:, 1].to(torch.int64) << 1) | (perms[:, 2].to(torch.int64) << 2))
    cell_fct_h = cell_fct.cpu().numpy()
    fct_node_h = fct_node.cpu().numpy()
    perms_h = perms.reshape(-1).cpu().numpy()
    info_h = info.cpu().numpy().astype(np.uint32)
    del cn, cell_fct, fct_node, perms, info
    nc_per_f = np.diff(fct_cell_off)
    bfct = np.nonzero(nc_per_f == 1)[0].astype(np.int32)
    mid = 0.5 * (x3[fct_node_h[bfct, 0]] + x3[fct_node_h[bfct, 1]])
    side = np.zeros(bfct.shape[0], dtype=np.int32)
    side[np.isclose(mid[:, 0], 0.0)] = 1
    side[np.isclose(mid[:, 1], 0.0)] = 2
    side[np.isclose(mid[:, 0], 1.0)] = 3
    side[np.isclose(mid[:, 1], 1.0)] = 4
    return Mesh(x=x3, cell_node=cell_node, cell_fct=cell_fct_h, fct_node=fct_node_h, fct_cell_off=fct_cell_off,
                fct_cell=fct_cell, node_cell_off=node_cell_off, node_cell=node_cell, node_fct_off=node_fct_off,
                node_fct=node_fct, fct_perms=np.ascontiguousarray(perms_h), cell_perm_info=info_h, bfct=bfct, bfct_side=side)


def _scramble(cell_node, rng):
    """Random permutation of the local vertex order of every cell: produces
    reversed facets and negative Jacobians (stand-in for the reference's gmsh
    fixture `python/test/unit/utils.py:98-137`)."""
    perm = np.argsort(rng.random(cell_node.shape), axis=1)
    return np.take_along_axis(cell_node, perm, axis=1)


def crossed_unit_square(n: int, scramble_seed: int | None = None, perturb: float = 0.0, seed: int = 7, fast: bool = False) -> Mesh:
    """n x n squares, each split into 4 triangles by both diagonals
    (`DiagonalType.crossed`, the reference's benchmark mesh `perftest.py:62-63`).

    Node ids: grid node (i, j) -> j (n+1) + i, then square centres
    (n+1)^2 + j n + i.  Cell 4 (j n + i) + t, t = bottom, right, top, left; local
    vertices ascending in global id."""
    i, j = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="xy")
    xg = np.stack([i.ravel() / n, j.ravel() / n], axis=1)
    ic, jc = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    xc = np.stack([(ic.ravel() + 0.5) / n, (jc.ravel() + 0.5) / n], axis=1)
    x = np.concatenate([xg, xc])
    if perturb > 0.0:
        rng = np.random.default_rng(seed)
        interior = np.ones(x.shape[0], dtype=bool)
        interior[: (n + 1) ** 2] = ((i.ravel() > 0) & (i.ravel() < n) & (j.ravel() > 0) & (j.ravel() < n))
        x[interior] += perturb / n * (rng.random((interior.sum(), 2)) - 0.5)

    sq_i, sq_j = ic.ravel(), jc.ravel()
    v00 = sq_j * (n + 1) + sq_i
    v10 = v00 + 1
    v01 = v00 + (n + 1)
    v11 = v01 + 1
    c = (n + 1) ** 2 + sq_j * n + sq_i
    tris = np.stack(
        [
            np.stack([v00, v10, c], axis=1),
            np.stack([v10, v11, c], axis=1),
            np.stack([v01, v11, c], axis=1),
            np.stack([v00, v01, c], axis=1),
        ],
        axis=1,
    ).reshape(-1, 3)
    if scramble_seed is not None:
        tris = _scramble(tris, np.random.default_rng(scramble_seed))
    return (build_topology_fast if fast else build_topology)(x, tris)


def random_diagonal_square(n: int, seed: int = 3, scramble_seed: int | None = None, perturb: float = 0.0) -> Mesh:
    """n x n squares split by one randomly chosen diagonal: vertex valences 2..8,
    i.e. patches with every cell count the reference accepts.  Corner squares get
    the diagonal that gives the corner node two cells (a 1-cell patch throws in
    the reference, `se/Patch.cpp:353-359`)."""
    rng = np.random.default_rng(seed)
    i, j = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="xy")
    x = np.stack([i.ravel() / n, j.ravel() / n], axis=1)
    if perturb > 0.0:
        interior = (i.ravel() > 0) & (i.ravel() < n) & (j.ravel() > 0) & (j.ravel() < n)
        x[interior] += perturb / n * (rng.random((interior.sum(), 2)) - 0.5)
    ic, jc = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    sq_i, sq_j = ic.ravel(), jc.ravel()
    v00 = sq_j * (n + 1) + sq_i
    v10, v01 = v00 + 1, v00 + (n + 1)
    v11 = v01 + 1
    # diag 0: v00-v11 ("right"), diag 1: v10-v01 ("left")
    diag = rng.integers(0, 2, size=n * n)
    diag[(sq_i == 0) & (sq_j == 0)] = 0
    diag[(sq_i == n - 1) & (sq_j == n - 1)] = 0
    diag[(sq_i == n - 1) & (sq_j == 0)] = 1
    diag[(sq_i == 0) & (sq_j == n - 1)] = 1
    t0 = np.where(diag[:, None] == 0, np.stack([v00, v10, v11], 1), np.stack([v00, v10, v01], 1))
    t1 = np.where(diag[:, None] == 0, np.stack([v00, v01, v11], 1), np.stack([v10, v01, v11], 1))
    tris = np.stack([t0, t1], axis=1).reshape(-1, 3)
    tris = np.sort(tris, axis=1)
    if scramble_seed is not None:
        tris = _scramble(tris, np.random.default_rng(scramble_seed))
    return build_topology(x, tris)


def fan_unit_square(nper: int, hub_on_boundary: bool = False, scramble_seed: int | None = None) -> Mesh:
    """Unit square triangulated as ONE fan: every cell joins the hub vertex with two neighbouring points of
    the perimeter (`nper` segments per side).  Hub in the interior: an interior patch with 4 nper cells; hub
    at the middle of the bottom side: a boundary patch with 3 nper cells.  All other patches have 2..4 cells.
    Stand-in for the high-valence vertices of unstructured meshes (patches with more than 16 cells)."""
    t = np.arange(nper) / nper
    one, zero = np.ones(nper), np.zeros(nper)
    sides = [np.stack([t, zero], 1), np.stack([one, t], 1), np.stack([1.0 - t, one], 1), np.stack([zero, 1.0 - t], 1)]
    if hub_on_boundary:
        ring = np.concatenate([np.array([[1.0, 0.0]]), sides[1][1:], sides[2], sides[3], np.array([[0.0, 0.0]])])
        hub = np.array([[0.5, 0.0]])
        nr = ring.shape[0]
        # the two corner points next to the hub would sit in one cell only (the reference rejects 1-cell
        # patches, se/Patch.cpp:353-359): split the first and last pair of fan cells through the midpoints
        # of the edges hub-ring[1] and hub-ring[nr-2]
        mids = 0.5 * (hub + ring[[1, nr - 2]])
        x = np.concatenate([ring, hub, mids])
        h, ma, mb = nr, nr + 1, nr + 2
        tris = [[h, i, i + 1] for i in range(2, nr - 3)]
        tris += [[h, 0, ma], [0, 1, ma], [h, ma, 2], [ma, 1, 2]]
        tris += [[h, mb, nr - 1], [mb, nr - 2, nr - 1], [h, nr - 3, mb], [nr - 3, nr - 2, mb]]
        tris = np.array(tris)
    else:
        ring = np.concatenate(sides)
        hub = np.array([[0.5, 0.5]])
        x = np.concatenate([ring, hub])
        nr = ring.shape[0]
        tris = np.stack([np.full(nr, nr), np.arange(nr), (np.arange(nr) + 1) % nr], axis=1)
    tris = np.sort(tris, axis=1)
    if scramble_seed is not None:
        tris = _scramble(tris, np.random.default_rng(scramble_seed))
    return build_topology(x, tris.astype(np.int32))


def delaunay_unit_square(nb: int, seed: int = 0, scramble_seed: int | None = None) -> Mesh:
    """Unstructured triangulation of the unit square (stand-in for the reference's gmsh fixture
    `python/test/unit/utils.py:98-137`): `nb` equal segments per side, jittered interior points at roughly the
    same spacing, Delaunay triangulation (scipy).  Vertex valences 3..9; every corner gets an interior point on its
    diagonal so that the corner patch has two cells (a 1-cell patch throws in the reference, `se/Patch.cpp:353-359`)."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    h = 1.0 / nb
    t = np.arange(nb) * h
    frame = np.concatenate([np.stack([t, 0 * t], 1), np.stack([1 + 0 * t, t], 1), np.stack([1 - t, 1 + 0 * t], 1),
                            np.stack([0 * t, 1 - t], 1)])
    gi, gj = np.meshgrid(np.arange(1, nb), np.arange(1, nb), indexing="xy")
    inner = np.stack([gi.ravel() * h, gj.ravel() * h], 1) + 0.3 * h * (rng.random(((nb - 1) ** 2, 2)) - 0.5)
    d = 0.38 * h
    corners = np.array([[d, d], [1 - d, d], [1 - d, 1 - d], [d, 1 - d]])
    # the grid points next to the corners would compete with the corner points: drop them
    keep = ~(((gi.ravel() == 1) | (gi.ravel() == nb - 1)) & ((gj.ravel() == 1) | (gj.ravel() == nb - 1)))
    x = np.concatenate([frame, inner[keep], corners])
    tri = Delaunay(x).simplices.astype(np.int32)
    # drop degenerate slivers on the frame (collinear boundary points)
    a, b, c = x[tri[:, 0]], x[tri[:, 1]], x[tri[:, 2]]
    area = 0.5 * np.abs((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1]))
    tri = np.sort(tri[area > 1e-12], axis=1)
    if scramble_seed is not None:
        tri = _scramble(tri, np.random.default_rng(scramble_seed))
    m = build_topology(x, tri)
    if np.diff(m.node_cell_off).min() < 2:
        raise RuntimeError("delaunay_unit_square: a vertex with fewer than two cells (try another seed)")
    return m


def submesh(mesh: Mesh, cells: np.ndarray):
    """Sub-mesh of the given cells with an ORDER-PRESERVING renumbering of vertices (and hence of facets and of
    the adjacency lists): a vertex whose whole patch lies in the sub-mesh sees exactly the patch it has in
    `mesh` (same start facet, same cell order, same facet orientations).  Returns (sub, nodes, facets): the
    vertex / facet ids of `mesh` for every vertex / facet of `sub`.  Used to replay windows of a full-size
    problem on the CPU oracle (tests/test_gpu_fullsize.py)."""
    cells = np.sort(np.asarray(cells, dtype=np.int64))
    cn = mesh.cell_node[cells]
    nodes = np.unique(cn)
    sub = build_topology(mesh.x[nodes], np.searchsorted(nodes, cn).astype(np.int32))
    pairs = nodes[sub.fct_node]
    key_full = mesh.fct_node[:, 0].astype(np.int64) * mesh.nnode + mesh.fct_node[:, 1]
    facets = np.searchsorted(key_full, pairs[:, 0].astype(np.int64) * mesh.nnode + pairs[:, 1])
    assert np.array_equal(mesh.fct_node[facets], pairs)
    return sub, nodes.astype(np.int64), facets.astype(np.int64)


def pk_dofmap(mesh: Mesh, q: int):
    """Cell dofmap [ncell][(q+1)(q+2)/2] of the continuous P_q space (q <= 3) in Basix' cell-local DOF order
    (vertices, edge interiors e0 e1 e2 from the lower to the higher global vertex, cell interior), numbered
    [vertex dofs | (q-1) per facet | interior per cell] - what `eqlb_set_primal_space` takes.  Returns (dofmap, ndofs)."""
    if not 1 <= q <= 3:
        raise ValueError("pk_dofmap: 1 <= q <= 3")
    nloc = (q + 1) * (q + 2) // 2
    ne = q - 1
    nint = nloc - 3 - 3 * ne
    dm = np.zeros((mesh.ncell, nloc), dtype=np.int64)
    dm[:, :3] = mesh.cell_node
    for f in range(3):
        a, b = FACET_VERTS[f]
        rev = mesh.cell_node[:, a] > mesh.cell_node[:, b]
        for i in range(ne):
            ii = np.where(rev, ne - 1 - i, i)
            dm[:, 3 + f * ne + i] = mesh.nnode + mesh.cell_fct[:, f].astype(np.int64) * ne + ii
    for i in range(nint):
        dm[:, 3 + 3 * ne + i] = mesh.nnode + mesh.nfct * ne + np.arange(mesh.ncell) * nint + i
    return dm.astype(np.int32), int(mesh.nnode + mesh.nfct * ne + mesh.ncell * nint)


def dg_dofmap(ncell: int, ndg: int) -> np.ndarray:
    """Cell dofmap of a DG_p space: dof = cell * ndg + local (DOLFINx layout)."""
    return (np.arange(ncell, dtype=np.int32)[:, None] * ndg + np.arange(ndg, dtype=np.int32)[None, :]).astype(np.int32)


def facet_types(mesh: Mesh, dirichlet_sides, neumann_sides) -> np.ndarray:
    """`facet_type[nfct]` int8 with the wire values of `base/Patch.hpp:27-33`:
    0 internal, 1 essnt_primal (Dirichlet of the primal problem), 2 essnt_dual
    (flux / Neumann BC)."""
    ft = np.zeros(mesh.nfct, dtype=np.int8)
    ft[mesh.boundary_facets(dirichlet_sides)] = 1
    if len(np.atleast_1d(neumann_sides)):
        ft[mesh.boundary_facets(neumann_sides)] = 2
    return ft
