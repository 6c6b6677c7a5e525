"""Boundary data: the host mirror `eqlb.boundarydata` (closed-form facet moments of a
polynomial traction) against the reference's OWN `base::BoundaryData` constructor
(`base/BoundaryData.cpp:279-633`, compiled unchanged in oracle/_ref) driven by FluxBC objects whose
boundary kernel evaluates the same polynomial - interpolation branch (`:580-597`) and facet-local
projection branch (`:511-578`) - on the 12 traction layouts of `test_stressqlb_bcond.py:167-190`."""

import numpy as np
import pytest

from common import make_mesh
from dolfinx_eqlb_b200 import eqlb, tables as tb
from oracle import pyref as pr

needs_ref = pytest.mark.skipif(not pr.available(), reason="oracle/_ref not built and /root/reference absent")
LAYOUTS = [[1], [2], [3], [4], [1, 2], [1, 3], [1, 4], [2, 3], [2, 4], [3, 4], [1, 2, 3], [2, 3, 4], [1, 3, 4], [1, 2, 4]]


def traction_case(m, k, nsides, nrhs, seed, ncoef=None):
    rng = np.random.default_rng(seed)
    dsides = [s for s in (1, 2, 3, 4) if s not in nsides]
    bfp = [m.boundary_facets(dsides)] * nrhs
    bcs = []
    for _ in range(nrhs):
        fc = m.boundary_facets(nsides)
        bcs.append([eqlb.fluxbc(fc, 2.0 * (rng.random((fc.shape[0], ncoef or k)) + 0.1))] if fc.size else [])
    return bfp, bcs


@needs_ref
@pytest.mark.parametrize("kind,n,scramble", [("crossed", 3, None), ("crossed", 4, 3), ("randdiag", 5, 2)])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_host_mirror_matches_reference_constructor(kind, n, scramble, k):
    m = make_mesh(kind, n, scramble, perturb=0.2)
    T = tb.make_tables(k)
    for nsides in LAYOUTS:
        bfp, bcs = traction_case(m, k, nsides, 2, seed=3)
        bd = eqlb.boundarydata(bcs, m, T, bfp, True)
        ft, bv, nob = pr.boundary_data(m, T, bcs, bfp, stress=True)
        assert np.array_equal(ft, bd.facet_type), nsides
        assert np.array_equal(nob, bd.node_on_stress_bnd), nsides
        for r in range(2):
            ref = bv[r]
            got = bd.bflux[r] if bd.bflux[r] is not None else np.zeros_like(ref)
            assert np.abs(got - ref).max() < 1e-13 * max(np.abs(ref).max(), 1.0), (nsides, r)


@needs_ref
@pytest.mark.parametrize("k", [1, 2, 3])
def test_projection_branch_reproduces_polynomial_tractions(k):
    """traction of degree <= k-1: the facet-local L2 projection (quadrature degree 2k) is exact"""
    m = make_mesh("crossed", 4, 3, perturb=0.2)
    T = tb.make_tables(k)
    bfp, bcs = traction_case(m, k, [1, 4], 1, seed=5)
    bd = eqlb.boundarydata(bcs, m, T, bfp, False)
    ft, bv, _ = pr.boundary_data(m, T, bcs, bfp, qdegree_proj=2 * k)
    assert np.array_equal(ft, bd.facet_type)
    assert np.abs(bd.bflux[0] - bv[0]).max() < 1e-12 * np.abs(bv[0]).max()
