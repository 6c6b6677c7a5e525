"""GPU parity of the cell-wise local projector (base/local_solver.hpp) via the C ABI."""

import numpy as np
import pytest

from common import make_mesh
from dolfinx_eqlb_b200 import eqlb, tables as tb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", [1, 2, 3])
def test_local_projection_parity(k):
    from oracle import pyoracle as po

    m = make_mesh("crossed", 6, 3, perturb=0.3)
    T = tb.make_tables(k)
    rng = np.random.default_rng(2)
    qv = [rng.standard_normal(m.ncell * T.nq) for _ in range(4)]  # 4 RHS at once (test_localsolver_multilhs.py)
    ref = po.local_project(m, T, qv)
    prob = eqlb._Problem(m, T, 1)
    got = eqlb.local_projection(prob, qv)
    for a, b in zip(got, ref):
        assert np.abs(a - b).max() < 1e-12 * max(1.0, np.abs(b).max())
    # ... and against the reference's own element loop / solvers (`base::local_solver_*`, oracle/_ref)
    from oracle import pyref as pr

    if pr.available():
        for solver in ("lu", "cholesky", "cg"):
            for a, b in zip(got, pr.local_solver(m, T, qv, solver)):
                assert np.abs(a - b).max() < 1e-12 * max(1.0, np.abs(b).max())


def test_projection_reproduces_polynomials():
    """Projecting a function of the space returns its coefficients (local == global projection
    for DG, test_localsolver_projection.py)."""
    m = make_mesh("crossed", 4, None)
    T = tb.make_tables(3)
    rng = np.random.default_rng(1)
    coef = rng.standard_normal((m.ncell, T.ndg))
    qv = (coef @ T.dg_q[0].T).ravel()
    prob = eqlb._Problem(m, T, 1)
    got = eqlb.local_projection(prob, [qv])[0]
    assert np.abs(got - coef.ravel()).max() < 1e-12


@pytest.mark.parametrize("k", [1, 2, 3])
def test_flux_l2norm(k):
    """Cell-wise ||sigma||^2 of the equilibrated DRT flux (error indicator of
    demo_error_estimation.py:96-101) against quadrature of the mapped basis."""
    import fem_mini as fm
    from common import PoissonCase

    m = make_mesh("crossed", 6, 3, perturb=0.3)
    case = PoissonCase(m, k, [[1, 4], []], seed=2, galerkin=False)
    eq = eqlb.FluxEqlbSE(k, m, case.F, case.G)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    eq.equilibrate_fluxes()
    got = eq.flux_l2norm()
    for r in range(case.nrhs):
        ref = fm.cell_l2norm_sq(m, case.T, eq.list_flux[r])
        assert np.abs(got[r] - ref).max() < 1e-12 * ref.max()
        assert (got[r] >= 0.0).all()
