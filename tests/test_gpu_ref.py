"""GPU parity against the reference's OWN code (`oracle/_ref/libeqlb_ref.so`: the reference
sources compiled unchanged, see tests/test_ref_pinning.py): device patch maps / SE 4-plane
DOF maps / EV sub-DOF maps bit-exact, SE and stress flux DOFs within 1e-10 relative
(BASELINE.json north_star).  The library is built in the build container and travels with
the snapshot; the tests skip if it is absent."""

import numpy as np
import pytest

from common import PoissonCase, make_mesh
from dolfinx_eqlb_b200 import eqlb
from oracle import pyref as pr

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not pr.available(), reason="oracle/_ref not available")]
RTOL = 1e-10
MESHES = [("crossed", 4, None), ("crossed", 5, 3), ("randdiag", 6, 2)]
INT_KEYS = ["ncells", "cells", "fcts", "inodes_local", "fcts_local", "type", "reversed", "reversion"]


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("kind,n,scramble", MESHES)
@pytest.mark.parametrize("k", [1, 2, 3])
def test_se_maps_and_flux_vs_reference(kind, n, scramble, k):
    m = make_mesh(kind, n, scramble, perturb=0.2)
    case = PoissonCase(m, k, [[1, 4], [1, 3], [2], [1, 3, 4]], seed=5)
    bc = case.oracle_bc()
    ref_maps = pr.se_patch_maps(m, case.T, bc)
    ref = pr.se_run(m, case.T, bc, case.G, case.F)
    eq = eqlb.FluxEqlbSE(k, m, case.F, case.G)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    got = eq.problem.patch_maps()
    for key in INT_KEYS:
        assert np.array_equal(got[key], ref_maps[key]), key
    dm = eq.problem.se_dofmaps()
    for key in ["dofmap", "projflux_fct", "bmarkers"]:
        assert np.array_equal(dm[key], ref_maps[key]), key
    eq.equilibrate_fluxes()
    for r in range(case.nrhs):
        assert rel_err(eq.list_flux[r], ref[r]) < RTOL


@pytest.mark.parametrize("kind,n,scramble", MESHES)
@pytest.mark.parametrize("k,nsides", [(2, []), (3, []), (2, [1]), (3, [1]), (3, [1, 2]), (3, [2, 3, 4]), (2, [1, 3]), (2, [1, 2]),
                                      (2, [2, 3, 4]), (2, [3]), (2, [2, 4]), (3, [1, 2, 4]), (2, [1, 4]), (3, [3, 4]),
                                      (4, []), (4, [1]), (4, [1, 2]), (4, [2, 3, 4])])  # degree 4: test_stressqlb_conditions.py:22
def test_stress_vs_reference(kind, n, scramble, k, nsides):
    from oracle import pyoracle as po
    from test_gpu_stress import elasticity_case

    m = make_mesh(kind, n, scramble, perturb=0.2)
    T, G, f, bfp, bcs, neu = elasticity_case(m, k, nsides, seed=3, galerkin=False)
    eq = eqlb.FluxEqlbSE(k, m, f, G, equilibrate_stress=True, estimate_korn_constant=True)
    eq.set_boundary_conditions(bfp, bcs)
    bd = eq.boundary_data
    bc = po.BCData(bd.facet_type, bd.bflux, bd.local_fct_id, bd.node_on_stress_bnd)
    ref, kref = pr.se_run(m, T, bc, G, f, stress=True, korn=True)
    eq.equilibrate_fluxes()
    for r in range(2):
        assert rel_err(eq.list_flux[r], ref[r]) < RTOL
    assert rel_err(eq.get_korn_constants(), np.sqrt(kref)) < 1e-12


@pytest.mark.parametrize("kind,n,scramble", MESHES)
@pytest.mark.parametrize("k", [1, 2, 3])
def test_ev_dofmaps_vs_reference(kind, n, scramble, k):
    m = make_mesh(kind, n, scramble, perturb=0.2)
    case = PoissonCase(m, k, [[1, 4], [2]], seed=5, galerkin=False)
    eq = eqlb.FluxEqlbEV(k, m, case.F, case.G)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    got = eq.problem.ev_dofmaps()
    bc = case.oracle_bc()
    for z in range(m.nnode):
        ref = pr.ev_patch_maps(m, case.T, bc, z)
        assert got["ncells"][z] == ref["ncells"]
        for key in ("cells", "fcts", "inodes_local", "dofs_elmt", "dofs_patch", "dofs_global", "list_patch", "list_global"):
            assert np.array_equal(got[key][z], ref[key]), (z, key)


@pytest.mark.parametrize("kind,n,scramble,hom", [("crossed", 4, 3, True), ("randdiag", 6, 2, True), ("crossed", 5, None, False)])
@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("nsets", [[[]], [[1, 4], [1, 3]]])
def test_ev_flux_vs_reference(kind, n, scramble, hom, k, nsets):
    """EV (headline path): CUDA null-space kernels vs the reference's dense-KKT ev::reconstruction"""
    m = make_mesh(kind, n, scramble, perturb=0.25)
    case = PoissonCase(m, k, nsets, seed=1, hom=hom)
    ref = pr.ev_run(m, case.T, case.oracle_bc(), case.G, case.F)
    eq = eqlb.FluxEqlbEV(k, m, case.F, case.G)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    eq.equilibrate_fluxes()
    for r in range(case.nrhs):
        assert rel_err(eq.list_flux[r], ref[r]) < RTOL


@pytest.mark.parametrize("k,p", [(2, 0), (3, 1), (4, 0), (4, 1), (4, 2), (4, 3)])
def test_se_lower_degree_data_vs_reference(k, p):
    """projected data of lower degree than k-1 (`se/reconstruction.hpp:358-369` allows p <= k-1)"""
    m = make_mesh("crossed", 4, 3, perturb=0.2)
    case = PoissonCase(m, k, [[1, 4]], seed=2, p=p, galerkin=False)
    ref = pr.se_run(m, case.T, case.oracle_bc(), case.G, case.F)
    eq = eqlb.FluxEqlbSE(k, m, case.F, case.G, degree_proj=p)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    eq.equilibrate_fluxes()
    assert rel_err(eq.list_flux[0], ref[0]) < RTOL


@pytest.mark.parametrize("n,scramble", [(6, 3), (16, None), (24, 2)])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_unstructured_mesh_vs_reference(n, scramble, k):
    """Delaunay triangulations (valences 2..9 mixed in every colour -> several lane classes per colour, generic
    kernel for the multi-RHS boundary patches): SE maps + fluxes (3 RHS with different BC sets), EV flux, stress +
    Korn constants (pure Dirichlet and one traction side) against the reference's own code"""
    from oracle import pyoracle as po
    from test_gpu_stress import elasticity_case

    m = make_mesh("delaunay", n, scramble)
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
        assert rel_err(eq.list_flux[r], ref[r]) < RTOL
    case = PoissonCase(m, k, [[1, 4]], seed=1, hom=True, galerkin=False)
    ref = pr.ev_run(m, case.T, case.oracle_bc(), case.G, case.F)
    ev = eqlb.FluxEqlbEV(k, m, case.F, case.G)
    ev.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    ev.equilibrate_fluxes()
    assert rel_err(ev.list_flux[0], ref[0]) < RTOL
    if k >= 2:
        for nsides in ([], [3]):
            T, G, f, bfp, bcs, neu = elasticity_case(m, k, nsides, seed=3, galerkin=False)
            bd = eqlb.boundarydata(bcs, m, T, bfp, True)
            bc = po.BCData(bd.facet_type, bd.bflux, bd.local_fct_id, bd.node_on_stress_bnd)
            try:
                sref, kref = pr.se_run(m, T, bc, G, f, stress=True, korn=True)
            except RuntimeError:
                continue  # a layout the reference itself refuses on this mesh (2-cell traction patches)
            es = eqlb.FluxEqlbSE(k, m, f, G, equilibrate_stress=True, estimate_korn_constant=True)
            es.set_boundary_conditions(bfp, bcs)
            es.equilibrate_fluxes()
            for r in range(2):
                assert rel_err(es.list_flux[r], sref[r]) < RTOL
            assert rel_err(es.get_korn_constants(), np.sqrt(kref)) < 1e-12


@pytest.mark.parametrize("seed", range(8))
def test_random_unstructured_cases_vs_reference(seed):
    """seeded random cases: Delaunay mesh size, vertex scrambling, degree, number of RHS and their traction sides are
    drawn at random; SE + EV (+ stress for k >= 2) against the reference's compiled code"""
    from oracle import pyoracle as po
    from test_gpu_stress import elasticity_case
    from dolfinx_eqlb_b200 import mesh as ms

    rng = np.random.default_rng(1000 + seed)
    nb = int(rng.integers(5, 14))
    m = ms.delaunay_unit_square(nb, seed=int(rng.integers(0, 10_000)), scramble_seed=int(rng.integers(0, 100)))
    k = int(rng.integers(1, 4))
    sets = [sorted(rng.choice([1, 2, 3, 4], size=int(rng.integers(0, 4)), replace=False).tolist()) for _ in range(int(rng.integers(1, 4)))]
    case = PoissonCase(m, k, sets, seed=seed, galerkin=False)
    ref = pr.se_run(m, case.T, case.oracle_bc(), case.G, case.F)
    eq = eqlb.FluxEqlbSE(k, m, case.F, case.G)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs, device=bool(seed % 2))
    eq.equilibrate_fluxes()
    for r in range(case.nrhs):
        assert rel_err(eq.list_flux[r], ref[r]) < RTOL, (nb, k, sets)
    caseh = PoissonCase(m, k, sets, seed=seed, hom=True, galerkin=False)
    ref = pr.ev_run(m, caseh.T, caseh.oracle_bc(), caseh.G, caseh.F)
    ev = eqlb.FluxEqlbEV(k, m, caseh.F, caseh.G)
    ev.set_boundary_conditions(caseh.list_bfct_prime, caseh.list_bcs)
    ev.equilibrate_fluxes()
    for r in range(caseh.nrhs):
        assert rel_err(ev.list_flux[r], ref[r]) < RTOL, (nb, k, sets)
    if k >= 2:
        T, G, f, bfp, bcs, neu = elasticity_case(m, k, sets[0], seed=seed, galerkin=False)
        bd = eqlb.boundarydata(bcs, m, T, bfp, True)
        bc = po.BCData(bd.facet_type, bd.bflux, bd.local_fct_id, bd.node_on_stress_bnd)
        try:
            sref, kref = pr.se_run(m, T, bc, G, f, stress=True, korn=True)
        except RuntimeError:
            return  # layout the reference refuses on this mesh
        es = eqlb.FluxEqlbSE(k, m, f, G, equilibrate_stress=True, estimate_korn_constant=True)
        es.set_boundary_conditions(bfp, bcs)
        es.equilibrate_fluxes()
        for r in range(2):
            assert rel_err(es.list_flux[r], sref[r]) < RTOL, (nb, k, sets)
        assert rel_err(es.get_korn_constants(), np.sqrt(kref)) < 1e-12
