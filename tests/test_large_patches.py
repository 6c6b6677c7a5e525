"""Patches with more than 16 cells (high-valence vertices of unstructured meshes).  The reference has no limit
(`se/Patch.cpp:23-104` sizes its storage from the mesh); the CUDA path serves up to 32 cells per patch for flux
degrees <= 3 on the generic kernel while all other patches of the mesh stay on the lane-per-cell kernels.
Fan meshes: one hub vertex joined to every point of the perimeter (interior hub: 20 / 32 cells, boundary hub: 24)."""

import numpy as np
import pytest

from common import PoissonCase, make_mesh
from dolfinx_eqlb_b200 import eqlb
from oracle import pyref as pr

needs_ref = pytest.mark.skipif(not pr.available(), reason="oracle/_ref not available")
FANS = [("fan", 5, 3), ("halffan", 8, 2), ("fan", 8, None)]
INT_KEYS = ["ncells", "cells", "fcts", "inodes_local", "fcts_local", "type", "reversed", "reversion"]


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@needs_ref
@pytest.mark.parametrize("kind,n,scramble", FANS)
@pytest.mark.parametrize("k", [1, 2, 3])
def test_oracle_matches_reference_on_large_patches(kind, n, scramble, k):
    from oracle import pyoracle as po

    m = make_mesh(kind, n, scramble)
    case = PoissonCase(m, k, [[1, 4], [2], []], seed=5, galerkin=False)
    bc = case.oracle_bc()
    a, b = po.se_patch_maps(m, case.T, bc), pr.se_patch_maps(m, case.T, bc)
    for key in a:
        if key != "ncmax":
            assert np.array_equal(a[key], b[key]), key
    for x, y in zip(po.se_run(m, case.T, bc, case.G, case.F), pr.se_run(m, case.T, bc, case.G, case.F)):
        assert rel_err(x, y) < 1e-11
    case = PoissonCase(m, k, [[1, 4], []], seed=1, hom=True, galerkin=False)
    bc = case.oracle_bc()
    for x, y in zip(po.ev_run(m, case.T, bc, case.G, case.F), pr.ev_run(m, case.T, bc, case.G, case.F)):
        assert rel_err(x, y) < 1e-10


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("kind,n,scramble", FANS)
@pytest.mark.parametrize("k", [1, 2, 3])
def test_gpu_large_patches_vs_reference(kind, n, scramble, k):
    m = make_mesh(kind, n, scramble)
    case = PoissonCase(m, k, [[1, 4], [2], []], seed=5, galerkin=False)
    bc = case.oracle_bc()
    ref_maps = pr.se_patch_maps(m, case.T, bc)
    ref = pr.se_run(m, case.T, bc, case.G, case.F)
    eq = eqlb.FluxEqlbSE(k, m, case.F, case.G)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    got = eq.problem.patch_maps()
    for key in INT_KEYS:
        assert np.array_equal(got[key], ref_maps[key]), key
    eq.equilibrate_fluxes()
    for r in range(case.nrhs):
        assert rel_err(eq.list_flux[r], ref[r]) < 1e-10
    case = PoissonCase(m, k, [[1, 4]], seed=1, hom=True, galerkin=False)
    ref = pr.ev_run(m, case.T, case.oracle_bc(), case.G, case.F)
    ev = eqlb.FluxEqlbEV(k, m, case.F, case.G)
    ev.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    ev.equilibrate_fluxes()
    assert rel_err(ev.list_flux[0], ref[0]) < 1e-10


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("kind,n,scramble", FANS)
@pytest.mark.parametrize("k,nsides", [(2, []), (2, [1]), (3, [1, 2]), (2, [2, 3, 4]), (3, [3])])
def test_gpu_large_patches_stress_vs_reference(kind, n, scramble, k, nsides):
    from oracle import pyoracle as po
    from test_gpu_stress import elasticity_case

    m = make_mesh(kind, n, scramble)
    T, G, f, bfp, bcs, neu = elasticity_case(m, k, nsides, seed=3, galerkin=False)
    bd = eqlb.boundarydata(bcs, m, T, bfp, True)
    bc = po.BCData(bd.facet_type, bd.bflux, bd.local_fct_id, bd.node_on_stress_bnd)
    eq = eqlb.FluxEqlbSE(k, m, f, G, equilibrate_stress=True, estimate_korn_constant=True)
    try:
        ref, kref = pr.se_run(m, T, bc, G, f, stress=True, korn=True)
    except RuntimeError as e:
        # a layout the reference itself refuses (two-cell corner patches on the traction boundary,
        # se/reconstruction.hpp:186-200): same refusal, same text
        with pytest.raises(RuntimeError) as info:
            eq.set_boundary_conditions(bfp, bcs)
            eq.equilibrate_fluxes()
        assert str(e).strip() in str(info.value)
        return
    eq.set_boundary_conditions(bfp, bcs)
    eq.equilibrate_fluxes()
    for r in range(2):
        assert rel_err(eq.list_flux[r], ref[r]) < 1e-10
    assert rel_err(eq.get_korn_constants(), np.sqrt(kref)) < 1e-12


@pytest.mark.gpu
def test_gpu_degree4_rejects_more_than_16_cells_and_accepts_16():
    """degree 4 keeps the 16-cell limit (frame of the dense patch system): loud error above, parity at 16"""
    m = make_mesh("fan", 5, 1)
    case = PoissonCase(m, 4, [[]], seed=2, galerkin=False)
    eq = eqlb.FluxEqlbSE(4, m, case.F, case.G)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    with pytest.raises(RuntimeError, match="more than 16 cells"):
        eq.equilibrate_fluxes()
    if pr.available():
        m = make_mesh("fan", 4, 1)
        case = PoissonCase(m, 4, [[1]], seed=2, galerkin=False)
        eq = eqlb.FluxEqlbSE(4, m, case.F, case.G)
        eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
        eq.equilibrate_fluxes()
        ref = pr.se_run(m, case.T, case.oracle_bc(), case.G, case.F)
        assert rel_err(eq.list_flux[0], ref[0]) < 1e-10


@pytest.mark.gpu
def test_gpu_lane_classes_on_mixed_valence_mesh():
    """random-diagonal mesh: valences 4..8 in every colour -> lane classes 4 and 8 are launched separately (or
    merged when one is small); result equals the generic thread-per-patch kernel (EQLB_FLAG_GENERIC)"""
    m = make_mesh("randdiag", 40, 2, perturb=0.2)
    for k in (1, 2, 3):
        case = PoissonCase(m, k, [[1, 4]], seed=5, galerkin=False)
        out = []
        for generic in (False, True):
            eq = eqlb.FluxEqlbEV(k, m, case.F, case.G, generic=generic)
            eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
            eq.equilibrate_fluxes()
            out.append(np.array(eq.list_flux[0]))
        assert rel_err(out[0], out[1]) < 1e-11
