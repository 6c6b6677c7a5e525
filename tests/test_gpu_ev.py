"""GPU parity tests of the EV (constrained minimisation) path through the C ABI.

The CUDA kernel solves every patch KKT system by the null-space method; the oracle
assembles and LU-factorises the dense KKT matrix like the reference.  Both must give
the same minimiser."""

import numpy as np
import pytest

import fem_mini as fm
from common import PoissonCase, make_mesh
from dolfinx_eqlb_b200 import eqlb

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def run_gpu(case):
    eq = eqlb.FluxEqlbEV(case.k, case.mesh, case.F, case.G)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    eq.equilibrate_fluxes()
    return eq


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("kind,n,scramble", [("crossed", 4, None), ("crossed", 5, 3), ("randdiag", 6, 2)])
@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("nsets,hom", [([[]], True), ([[1, 4]], True), ([[1, 4]], False), ([[1, 4], [1, 3], [2], [1, 3, 4]], False)])
def test_ev_flux_parity(kind, n, scramble, k, nsets, hom):
    from oracle import pyoracle as po

    m = make_mesh(kind, n, scramble, perturb=0.25)
    case = PoissonCase(m, k, nsets, seed=7, hom=hom)
    ref = po.ev_run(m, case.T, case.oracle_bc(), case.G, case.F)
    eq = run_gpu(case)
    for r in range(case.nrhs):
        assert rel_err(eq.list_flux[r], ref[r]) < RTOL
        s = fm.conforming_to_drt(m, case.T, eq.list_flux[r])
        z = np.zeros_like(case.G[r])
        # div sigma_h = Pi f to machine precision; conformity
        assert fm.check_divergence(m, case.T, s, z, case.F[r]) < 1e-11
        assert fm.check_jump(m, case.T, s, z) < 1e-11


def test_ev_random_data_parity():
    """Benchmark input distribution (not Galerkin orthogonal): exercises the mean-value
    multiplier of the KKT system."""
    from oracle import pyoracle as po

    m = make_mesh("crossed", 8, None)
    case = PoissonCase(m, 2, [[]], seed=11, galerkin=False)
    ref = po.ev_run(m, case.T, case.oracle_bc(), case.G, case.F)
    eq = run_gpu(case)
    assert rel_err(eq.list_flux[0], ref[0]) < RTOL


def test_ev_random_data_neumann():
    from oracle import pyoracle as po

    m = make_mesh("crossed", 5, 2, perturb=0.2)
    case = PoissonCase(m, 2, [[1, 2, 3]], seed=4, galerkin=False)
    ref = po.ev_run(m, case.T, case.oracle_bc(), case.G, case.F)
    eq = run_gpu(case)
    assert rel_err(eq.list_flux[0], ref[0]) < RTOL


@pytest.mark.parametrize("kind,n,scramble", [("crossed", 4, None), ("crossed", 5, 3), ("randdiag", 6, 2)])
@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("nsets", [[[]], [[1, 4], [1, 3]]])
def test_ev_dofmaps_bit_exact(kind, n, scramble, k, nsets):
    """EV patch ordering and sub-DOF maps (ev/Patch.cpp:83-309, 482-676) built on the device
    are bit-exact against the oracle, patch by patch."""
    from oracle import pyoracle as po

    m = make_mesh(kind, n, scramble, perturb=0.25)
    case = PoissonCase(m, k, nsets, seed=1, hom=True)
    eq = eqlb.FluxEqlbEV(case.k, case.mesh, case.F, case.G)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    got = eq.problem.ev_dofmaps()
    for z in range(m.nnode):
        ref = po.ev_patch_maps(m, case.T, case.oracle_bc(), z)
        assert got["ncells"][z] == ref["ncells"]
        for key in ("cells", "fcts", "inodes_local", "dofs_elmt", "dofs_patch", "dofs_global", "list_patch", "list_global"):
            assert np.array_equal(got[key][z], ref[key]), (z, key)
