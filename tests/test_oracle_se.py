"""CPU: the oracle (restatement of the reference SE path) satisfies the reference's own
acceptance invariants (test_fluxeqlb_conditions.py:145-174, test_fluxeqlb_multirhs.py)."""

import numpy as np
import pytest

import fem_mini as fm
from common import PoissonCase, make_mesh

BC_SETS = [[1, 4], [1, 3], [2], [1, 3, 4]]  # test_fluxeqlb_multirhs.py:70


@pytest.mark.parametrize("kind,n,scramble", [("crossed", 2, None), ("crossed", 3, 5), ("randdiag", 5, 2)])
@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("nsides", [[], [1, 4], [1, 3, 4]])
def test_invariants(kind, n, scramble, k, nsides):
    from oracle import pyoracle as po

    m = make_mesh(kind, n, scramble, perturb=0.3)
    case = PoissonCase(m, k, [nsides], seed=1)
    sig = po.se_run(m, case.T, case.oracle_bc(), case.G, case.F)[0]
    assert fm.check_divergence(m, case.T, sig, case.G[0], case.F[0]) < 1e-12
    assert fm.check_jump(m, case.T, sig, case.G[0]) < 1e-11
    if nsides:
        assert fm.check_bc(m, case.T, sig, case.G[0], case.bdata.bflux[0], case.neu[0]) < 1e-11


@pytest.mark.parametrize("k", [1, 2, 3])
def test_multi_rhs_equals_single_rhs(k):
    from oracle import pyoracle as po

    m = make_mesh("crossed", 4, 3, perturb=0.25)
    case = PoissonCase(m, k, BC_SETS, seed=7)
    multi = po.se_run(m, case.T, case.oracle_bc(), case.G, case.F)
    for r in range(4):
        bc = po.BCData(case.bdata.facet_type[r : r + 1], [case.bdata.bflux[r]], case.bdata.local_fct_id)
        one = po.se_run(m, case.T, bc, [case.G[r]], [case.F[r]])[0]
        assert np.abs(one - multi[r]).max() < 1e-12 * max(1.0, np.abs(one).max())


def test_lower_degree_data():
    from oracle import pyoracle as po

    m = make_mesh("crossed", 3, 2, perturb=0.2)
    for k, p in [(2, 0), (3, 1)]:
        case = PoissonCase(m, k, [[1, 4]], seed=2, p=p)
        sig = po.se_run(m, case.T, case.oracle_bc(), case.G, case.F)[0]
        assert fm.check_divergence(m, case.T, sig, case.G[0], case.F[0]) < 1e-12
        assert fm.check_jump(m, case.T, sig, case.G[0]) < 1e-11


def test_accumulates_and_is_linear():
    from oracle import pyoracle as po

    m = make_mesh("crossed", 3, None)
    case = PoissonCase(m, 2, [[]], seed=3, galerkin=False)
    bc = case.oracle_bc()
    a = po.se_run(m, case.T, bc, case.G, case.F)[0]
    b = po.se_run(m, case.T, bc, case.G, case.F, sigma0=[a])[0]
    assert np.abs(b - 2 * a).max() < 1e-13
    c = po.se_run(m, case.T, bc, [3.0 * case.G[0]], [3.0 * case.F[0]])[0]
    assert np.abs(c - 3 * a).max() < 1e-12


def test_patch_maps_structure():
    from oracle import pyoracle as po

    m = make_mesh("crossed", 3, 4)
    case = PoissonCase(m, 2, [[1, 4], [2]], seed=3, galerkin=False)
    mp = po.se_patch_maps(m, case.T, case.oracle_bc())
    nc = np.diff(m.node_cell_off)
    assert (mp["ncells"] == nc).all()
    for z in range(m.nnode):
        n = nc[z]
        cells = set(m.node_cell[m.node_cell_off[z] : m.node_cell_off[z + 1]].tolist())
        assert set(mp["cells"][z, 1 : n + 1].tolist()) == cells
        internal = mp["type"][z, 0] == 0
        if internal:
            assert mp["cells"][z, 0] == mp["cells"][z, n] and mp["cells"][z, n + 1] == mp["cells"][z, 1]
        # consecutive cells share the facet between them
        for a in range(1, n):
            f = mp["fcts"][z, a]
            assert f in m.cell_fct[mp["cells"][z, a]] and f in m.cell_fct[mp["cells"][z, a + 1]]


def test_one_cell_patch_raises():
    from oracle import pyoracle as po
    from dolfinx_eqlb_b200 import mesh as ms, tables as tb

    x = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    m = ms.build_topology(x, np.array([[0, 1, 3], [0, 2, 3]]))
    T = tb.make_tables(1)
    ft = ms.facet_types(m, [1, 2, 3, 4], [])
    with pytest.raises(RuntimeError, match="has only 1 cells"):
        po.se_run(m, T, po.BCData(ft[None, :]), [np.zeros(4)], [np.zeros(2)])
