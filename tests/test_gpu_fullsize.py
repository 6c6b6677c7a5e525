"""Size-independent properties at BASELINE.json's full single-GPU size (configs[1]:
1024x1024 crossed mesh, 2.1 M patches; the oracle cannot run there in seconds).

* the specialised lane-per-cell kernels agree with the generic thread-per-patch kernel
  (which is the one compared against the oracle patch by patch at small sizes),
* both equilibrations are linear maps of (G, f) for a fixed mesh and BC set,
* the colour-ordered accumulation is bitwise reproducible and accumulates (`+=`) like
  `se/solve_patch_semiexplt.hpp:1159` / `ev/solve_patch.hpp:216-227`,
* the staged host pipeline gives the one-piece result."""

import numpy as np
import pytest

from dolfinx_eqlb_b200 import eqlb, mesh as ms, tables as tb

pytestmark = pytest.mark.gpu

N = 1024


def inputs(ncell, ndg, nrhs, seed):
    rng = np.random.default_rng(seed)
    G = [2.0 * (rng.random(ncell * ndg * 2) - 0.5) for _ in range(nrhs)]
    F = [2.0 * (rng.random(ncell * ndg) + 0.1) for _ in range(nrhs)]
    return G, F


@pytest.fixture(scope="module")
def big_mesh():
    return ms.crossed_unit_square(N)


def run(cls, k, m, G, F, bfct, twice=False, **kw):
    eq = cls(k, m, F, G, **kw)
    eq.set_boundary_conditions(bfct, [[] for _ in G])
    eq.equilibrate_fluxes()
    if twice:
        eq.equilibrate_fluxes()
    return eq.list_flux


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("path,k", [("ev", 2), ("se", 2), ("se", 1), ("ev", 3)])
def test_fullsize_properties(big_mesh, path, k):
    m = big_mesh
    T = tb.make_tables(k)
    cls = eqlb.FluxEqlbEV if path == "ev" else eqlb.FluxEqlbSE
    bfct = [m.boundary_facets([1, 2, 3, 4])]
    G1, F1 = inputs(m.ncell, T.ndg, 1, 1)
    G2, F2 = inputs(m.ncell, T.ndg, 1, 2)
    a = run(cls, k, m, G1, F1, bfct, host_pipeline=False)[0]
    assert np.isfinite(a).all()
    # specialised kernels == generic kernel
    g = run(cls, k, m, G1, F1, bfct, host_pipeline=False, generic=True)[0]
    assert rel(a, g) < 1e-10
    # bitwise reproducible, accumulating
    a2 = run(cls, k, m, G1, F1, bfct, host_pipeline=False)[0]
    assert np.array_equal(a, a2)
    d = run(cls, k, m, G1, F1, bfct, twice=True, host_pipeline=False)[0]
    assert rel(d, 2.0 * a) < 1e-13
    # staged host pipeline
    s = run(cls, k, m, G1, F1, bfct, host_pipeline=True)[0]
    assert rel(s, a) < 1e-12
    # linearity
    b = run(cls, k, m, G2, F2, bfct, host_pipeline=True)[0]
    ab = run(cls, k, m, [G1[0] + 0.5 * G2[0]], [F1[0] + 0.5 * F2[0]], bfct, host_pipeline=True)[0]
    assert rel(ab, a + 0.5 * b) < 1e-10


def test_fullsize_stress_fused_vs_generic(big_mesh):
    """Elasticity rows with weak symmetry (configs[3] layout) at 1024^2: fused lane-per-cell
    stage against the generic kernel."""
    m = big_mesh
    T = tb.make_tables(2)
    bfct = [m.boundary_facets([1, 2, 3, 4])] * 2
    G, F = inputs(m.ncell, T.ndg, 2, 5)
    a = run(eqlb.FluxEqlbSE, 2, m, G, F, bfct, equilibrate_stress=True, host_pipeline=False)
    g = run(eqlb.FluxEqlbSE, 2, m, G, F, bfct, equilibrate_stress=True, host_pipeline=False, generic=True)
    for r in range(2):
        assert rel(a[r], g[r]) < 1e-10
