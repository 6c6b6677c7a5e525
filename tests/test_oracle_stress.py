"""CPU: weak symmetry + Korn constants of the oracle (test_stressqlb_conditions.py:21-181)."""

import numpy as np
import pytest

import fem_mini as fm
from common import make_mesh, neumann_coeffs
from dolfinx_eqlb_b200 import eqlb, tables as tb


def run(m, k, nsides, seed=3, korn=False):
    from oracle import pyoracle as po

    rng = np.random.default_rng(seed)
    T = tb.make_tables(k)
    dsides = [s for s in (1, 2, 3, 4) if s not in nsides]
    f = [fm.random_dg(rng, m.ncell * T.ndg) for _ in range(2)]
    neu = [neumann_coeffs(m, T, nsides, rng) for _ in range(2)]
    G = fm.solve_elasticity(m, k, T, f, dsides, neu)
    bcs = []
    for n in neu:
        fc = np.array(sorted(n.keys()), dtype=np.int32)
        bcs.append([eqlb.fluxbc(fc, np.array([n[int(q)] for q in fc]))] if len(fc) else [])
    bd = eqlb.boundarydata(bcs, m, T, [m.boundary_facets(dsides)] * 2, True)
    bc = po.BCData(bd.facet_type, bd.bflux, bd.local_fct_id, bd.node_on_stress_bnd)
    out = po.se_run(m, T, bc, G, f, stress=True, korn=korn)
    return T, G, f, bd, neu, out


@pytest.mark.parametrize("k", [2, 3])
@pytest.mark.parametrize("scramble", [None, 4])
@pytest.mark.parametrize("nsides", [[], [1], [1, 2], [1, 3]])
def test_stress_invariants(k, scramble, nsides):
    m = make_mesh("crossed", 4, scramble, perturb=0.2)
    T, G, f, bd, neu, s = run(m, k, nsides)
    for r in range(2):
        assert fm.check_divergence(m, T, s[r], G[r], f[r]) < 1e-12
        assert fm.check_jump(m, T, s[r], G[r]) < 1e-11
        if nsides:
            assert fm.check_bc(m, T, s[r], G[r], bd.bflux[r], neu[r]) < 1e-11
    assert fm.check_weak_symmetry(m, T, s[0], s[1]) < 1e-11


def test_korn_constants_structured():
    """Interior patches of the crossed unit square: minimal angle 45 deg at the boundary
    nodes => c = 2 / sin^2(pi/8); every cell collects 3 patches x (dim+1)."""
    m = make_mesh("crossed", 4, None)
    T, G, f, bd, neu, (s, kc) = run(m, 2, [], korn=True)
    assert kc.shape == (m.ncell,) and (kc > 0).all()
    c_int = 2.0 / np.sin(np.pi / 8) ** 2
    # a cell whose three vertices are interior
    xc = m.x[m.cell_node].mean(axis=1)
    inner = np.all((m.x[m.cell_node][:, :, :2] > 1e-9) & (m.x[m.cell_node][:, :, :2] < 1 - 1e-9), axis=(1, 2))
    assert inner.any()
    assert np.allclose(kc[inner], 3 * 3 * c_int, rtol=1e-12)
