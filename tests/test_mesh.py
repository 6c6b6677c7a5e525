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
