"""GPU: boundary data built on the device (`eqlb_set_bcs_poly`, csrc/bc_kernel.cu) == host mirror ==
the reference's BoundaryData constructor; equilibration with device-built BCs == with uploaded BCs."""

import numpy as np
import pytest

from common import make_mesh
from dolfinx_eqlb_b200 import eqlb, tables as tb
from oracle import pyref as pr
from test_ref_bcs import LAYOUTS, traction_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,n,scramble", [("crossed", 4, 3), ("randdiag", 5, 2)])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_device_boundary_data(kind, n, scramble, k):
    m = make_mesh(kind, n, scramble, perturb=0.2)
    T = tb.make_tables(k)
    rng = np.random.default_rng(1)
    nrhs = 2
    G = [rng.standard_normal(m.ncell * T.ndg * 2) for _ in range(nrhs)]
    F = [rng.standard_normal(m.ncell * T.ndg) for _ in range(nrhs)]
    for nsides in LAYOUTS:
        bfp, bcs = traction_case(m, k, nsides, nrhs, seed=3)
        stress = k >= 2
        host = eqlb.boundarydata(bcs, m, T, bfp, stress)
        eq = eqlb.FluxEqlbSE(k, m, F, G, equilibrate_stress=stress)
        eq.set_boundary_conditions(bfp, bcs, device=True)
        dev = eq.boundary_data
        assert np.array_equal(dev.facet_type, host.facet_type), nsides
        dual = (host.facet_type == 2).any(axis=0)
        assert np.array_equal(dev.local_fct_id[dual], host.local_fct_id[dual]), nsides
        if stress:
            assert np.array_equal(dev.node_on_stress_bnd, host.node_on_stress_bnd), nsides
        for r in range(nrhs):
            want = host.bflux[r] if host.bflux[r] is not None else np.zeros(m.ncell * T.nrt)
            assert np.abs(dev.bflux[r] - want).max() <= 1e-15 * max(np.abs(want).max(), 1.0), (nsides, r)
        if pr.available():
            ft, bv, nob = pr.boundary_data(m, T, bcs, bfp, stress=stress)
            assert np.array_equal(dev.facet_type, ft)
            for r in range(nrhs):
                assert np.abs(dev.bflux[r] - bv[r]).max() < 1e-13 * max(np.abs(bv[r]).max(), 1.0)
        # same flux as with host-built boundary data
        eq.equilibrate_fluxes()
        eq2 = eqlb.FluxEqlbSE(k, m, F, G, equilibrate_stress=stress)
        eq2.set_boundary_conditions(bfp, bcs)
        eq2.equilibrate_fluxes()
        for r in range(nrhs):
            assert np.abs(eq.list_flux[r] - eq2.list_flux[r]).max() <= 1e-13 * np.abs(eq2.list_flux[r]).max(), nsides


def test_wrong_local_fct_id_is_rejected():
    m = make_mesh("crossed", 4, None)
    T = tb.make_tables(2)
    bfp, bcs = traction_case(m, 2, [1], 1, seed=2)
    bd = eqlb.boundarydata(bcs, m, T, bfp, False)
    eq = eqlb.FluxEqlbSE(2, m, [np.zeros(m.ncell * T.ndg)], [np.zeros(m.ncell * T.ndg * 2)])
    bd.local_fct_id = (bd.local_fct_id + 1) % 3
    with pytest.raises(RuntimeError, match="local_fct_id"):
        eq.problem.set_bcs(bd)
