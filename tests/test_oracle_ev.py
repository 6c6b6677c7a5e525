"""CPU: the EV oracle (dense KKT + partial-pivot LU like the reference) fulfils
div sigma = Pi f, H(div) conformity and the flux BCs (test_fluxeqlb_conditions.py)."""

import numpy as np
import pytest

import fem_mini as fm
from common import PoissonCase, make_mesh


@pytest.mark.parametrize("kind,n,scramble", [("crossed", 2, None), ("crossed", 3, 5), ("randdiag", 4, 2)])
@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("nsides,hom", [([], True), ([1, 4], True), ([1, 4], False)])
def test_invariants(kind, n, scramble, k, nsides, hom):
    from oracle import pyoracle as po

    m = make_mesh(kind, n, scramble, perturb=0.3)
    case = PoissonCase(m, k, [nsides], seed=1, hom=hom)
    sg = po.ev_run(m, case.T, case.oracle_bc(), case.G, case.F)[0]
    s = fm.conforming_to_drt(m, case.T, sg)
    z = np.zeros_like(case.G[0])
    assert fm.check_divergence(m, case.T, s, z, case.F[0]) < 1e-10
    assert fm.check_jump(m, case.T, s, z) < 1e-11
    if nsides:
        assert fm.check_bc(m, case.T, s, z, case.bdata.bflux[0], case.neu[0]) < 1e-10 or hom


def test_ev_close_to_se_plus_projected_flux():
    """EV minimises ||sigma - hat G||, SE ||sigma_eq||: same constraints, so the results
    agree up to the (small) part of hat*G outside RT_k - a sanity check, not an identity."""
    from oracle import pyoracle as po

    m = make_mesh("crossed", 4, None)
    case = PoissonCase(m, 2, [[]], seed=5)
    se = po.se_run(m, case.T, case.oracle_bc(), case.G, case.F)[0]
    ev = fm.conforming_to_drt(m, case.T, po.ev_run(m, case.T, case.oracle_bc(), case.G, case.F)[0])
    z = np.zeros_like(case.G[0])
    # both are equilibrated: div(se + G) = f, div(ev) = f
    assert fm.check_divergence(m, case.T, se, case.G[0], case.F[0]) < 1e-12
    assert fm.check_divergence(m, case.T, ev, z, case.F[0]) < 1e-10


def test_patch_maps_cover_patch():
    from oracle import pyoracle as po

    m = make_mesh("crossed", 3, 2)
    case = PoissonCase(m, 2, [[1, 4]], seed=3, galerkin=False)
    T = case.T
    for z in range(m.nnode):
        mp = po.ev_patch_maps(m, T, case.oracle_bc(), z)
        n = mp["ncells"]
        cells = set(m.node_cell[m.node_cell_off[z] : m.node_cell_off[z + 1]].tolist())
        assert set(mp["cells"][:n].tolist()) == cells
        nf = m.node_fct_off[z + 1] - m.node_fct_off[z]
        ndof = nf * T.k + n * (T.k * T.k - T.k + T.ndg)
        used = mp["dofs_patch"][: n * (T.nrt + T.ndg - T.k)]
        assert used.max() == ndof - 1 and set(used.tolist()) == set(range(ndof))
