import numpy as np
import pytest

from dolfinx_eqlb_b200 import mesh as ms


@pytest.mark.parametrize("make", [lambda: ms.crossed_unit_square(5), lambda: ms.crossed_unit_square(4, scramble_seed=1, perturb=0.3),
                                  lambda: ms.random_diagonal_square(7, scramble_seed=2)])
def test_topology_consistency(make):
    m = make()
    # Euler: V - E + F = 1 for a disc
    assert m.nnode - m.nfct + m.ncell == 1
    # facet f is opposite local vertex f
    for f in range(3):
        a, b = ms.FACET_VERTS[f]
        lo = np.minimum(m.cell_node[:, a], m.cell_node[:, b])
        hi = np.maximum(m.cell_node[:, a], m.cell_node[:, b])
        assert (m.fct_node[m.cell_fct[:, f], 0] == lo).all()
        assert (m.fct_node[m.cell_fct[:, f], 1] == hi).all()
        assert (m.fct_perms.reshape(-1, 3)[:, f] == (m.cell_node[:, a] > m.cell_node[:, b])).all()
    # facets sorted lexicographically, adjacency lists ascending
    key = m.fct_node[:, 0].astype(np.int64) * m.nnode + m.fct_node[:, 1]
    assert (np.diff(key) > 0).all()
    for off, dat in [(m.fct_cell_off, m.fct_cell), (m.node_cell_off, m.node_cell), (m.node_fct_off, m.node_fct)]:
        for i in range(len(off) - 1):
            seg = dat[off[i] : off[i + 1]]
            assert (np.diff(seg) > 0).all()
    # boundary facets = facets with one cell, all on the unit-square frame
    assert (np.diff(m.fct_cell_off)[m.bfct] == 1).all()
    assert (m.bfct_side > 0).all()
    # no 1-cell patches
    assert np.diff(m.node_cell_off).min() >= 2


def test_crossed_counts():
    n = 6
    m = ms.crossed_unit_square(n)
    assert m.ncell == 4 * n * n
    assert m.nnode == (n + 1) ** 2 + n * n
    nc = np.diff(m.node_cell_off)
    assert set(nc[(n + 1) ** 2 :]) == {4}
    assert nc.max() == 8


def test_fast_topology_builder_is_bit_identical():
    """torch-sorted builder (used for the 2048^2 / 4096^2 benchmark meshes) == numpy builder"""
    for n, scr in [(4, None), (6, 3)]:
        a = ms.crossed_unit_square(n, scr)
        b = ms.crossed_unit_square(n, scr, fast=True)
        for f in a.__dataclass_fields__:
            x, y = getattr(a, f), getattr(b, f)
            assert x.dtype == y.dtype and np.array_equal(x, y), f


def test_strong_scaling_strips_tile_the_global_mesh():
    """`dist.crossed_rows`: every rank's local mesh is the restriction of crossed_unit_square(n)
    (same cells, same geometry, order preserving), every vertex is owned exactly once."""
    from dolfinx_eqlb_b200 import dist as dd

    n, world = 12, 3
    g = ms.crossed_unit_square(n)
    owned = np.zeros(g.nnode, dtype=np.int32)
    for r in range(world):
        part, nn = dd.crossed_rows(n, r, world, fast=False)
        assert nn == g.nnode
        lm = part.mesh
        assert np.array_equal(part.node_gid[lm.cell_node], g.cell_node[part.cell_gid])
        assert np.abs(lm.x - g.x[part.node_gid]).max() < 1e-15
        np.add.at(owned, part.node_gid[part.node_owned.astype(bool)], 1)
        # boundary ids of the global square only (cut lines are not boundaries)
        gl = part.node_gid[lm.fct_node[lm.bfct]]
        gkey = g.fct_node[g.bfct, 0].astype(np.int64) * g.nnode + g.fct_node[g.bfct, 1]
        side_of = dict(zip(gkey.tolist(), g.bfct_side.tolist()))
        for (a, b), s in zip(gl, lm.bfct_side):
            assert side_of.get(int(a) * g.nnode + int(b), 0) == s
    assert (owned == 1).all()
