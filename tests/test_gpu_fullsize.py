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
    # host-pointer calls: the first call of an equilibrator runs from pageable memory (copies staged through the
    # library's pinned pool), the second one from the then page-locked vectors through the stage pipeline
    s = run(cls, k, m, G1, F1, bfct, host_pipeline=True)[0]
    assert rel(s, a) < 1e-12
    s2 = run(cls, k, m, G1, F1, bfct, twice=True, host_pipeline=True)[0]
    assert rel(s2, 2.0 * a) < 1e-12
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


# ---------------------------------------------------------------------------------------------------
# Full-size results against the reference's own code on WINDOWS of the mesh.  A patch only sees its own
# cells, so the DOFs of a cell whose three vertex patches lie inside a window are the same numbers when
# the window is equilibrated alone (order-preserving sub-mesh: same start facets, cell order, facet
# orientations).  The window runs through oracle/_ref (the reference sources compiled unchanged) or, if
# that library is absent, through the oracle port.
# ---------------------------------------------------------------------------------------------------
def cpu_impl():
    from oracle import pyref as pr

    if pr.available():
        return pr
    from oracle import pyoracle as po

    return po


def window(m, i0, j0, w):
    """cells of the squares [i0, i0+w) x [j0, j0+w) of the crossed N x N mesh, the sub-mesh and the masks of
    its vertices / cells / facets that are untouched by the artificial window boundary"""
    ii, jj = np.meshgrid(np.arange(i0, i0 + w), np.arange(j0, j0 + w), indexing="xy")
    cells = (4 * (jj.ravel() * N + ii.ravel())[:, None] + np.arange(4)[None, :]).ravel()
    sub, nodes, facets = ms.submesh(m, cells)
    cells = np.sort(cells)
    artificial = sub.bfct[sub.bfct_side == 0]
    good_node = np.ones(sub.nnode, bool)
    good_node[sub.fct_node[artificial].ravel()] = False
    good_cell = good_node[sub.cell_node].all(axis=1)
    good_fct = good_node[sub.fct_node].all(axis=1)
    return sub, cells, nodes, facets, good_node, good_cell, good_fct


WINDOWS = [(500, 500, 12), (0, 0, 10), (N - 9, 300, 9), (400, N - 8, 8)]  # centre, corner, right side, top side


@pytest.mark.parametrize("i0,j0,w", WINDOWS)
def test_fullsize_ev_headline_vs_reference_on_windows(big_mesh, i0, j0, w):
    """configs[1] (EV k=2, 1024^2, pure Dirichlet): facet and cell DOFs of the CUDA result inside a window
    against the reference's ev::reconstruction on the window"""
    from oracle import pyoracle as po

    m, k = big_mesh, 2
    T = tb.make_tables(k)
    G, F = inputs(m.ncell, T.ndg, 1, 7)
    sig = run(eqlb.FluxEqlbEV, k, m, G, F, [m.boundary_facets([1, 2, 3, 4])], host_pipeline=False)[0]
    sub, cells, nodes, facets, good_node, good_cell, good_fct = window(m, i0, j0, w)
    Gs = [G[0].reshape(m.ncell, -1)[cells].ravel()]
    Fs = [F[0].reshape(m.ncell, -1)[cells].ravel()]
    ft = np.zeros((1, sub.nfct), np.int8)
    ft[0, sub.bfct] = 1
    ref = cpu_impl().ev_run(sub, T, po.BCData(ft), Gs, Fs)[0]
    nci = k * k - k
    f_full = sig[: m.nfct * k].reshape(m.nfct, k)[facets[good_fct]]
    f_ref = ref[: sub.nfct * k].reshape(sub.nfct, k)[good_fct]
    c_full = sig[m.nfct * k:].reshape(m.ncell, nci)[cells[good_cell]]
    c_ref = ref[sub.nfct * k:].reshape(sub.ncell, nci)[good_cell]
    assert good_cell.sum() > 100
    scale = max(np.abs(f_ref).max(), np.abs(c_ref).max())
    assert np.abs(f_full - f_ref).max() < 1e-10 * scale
    assert np.abs(c_full - c_ref).max() < 1e-10 * scale


@pytest.mark.parametrize("i0,j0,w", WINDOWS)
@pytest.mark.parametrize("k,stress", [(3, False), (2, True)])
def test_fullsize_se_vs_reference_on_windows(big_mesh, i0, j0, w, k, stress):
    """configs[2] (SE k=3, mixed flux BCs: tractions on sides 1 and 4) and configs[3] (elasticity rows with weak
    symmetry) at 1024^2 against the reference's se::reconstruction on windows"""
    from oracle import pyoracle as po

    m = big_mesh
    T = tb.make_tables(k)
    nrhs = 2 if stress else 1
    G, F = inputs(m.ncell, T.ndg, nrhs, 11)
    nsides = [] if stress else [1, 4]
    rng = np.random.default_rng(3)
    nf = m.boundary_facets(nsides) if nsides else np.zeros(0, np.int32)
    coeffs = rng.standard_normal((nf.shape[0], k))
    prime = m.boundary_facets([s for s in (1, 2, 3, 4) if s not in nsides])
    bcs = [[eqlb.fluxbc(nf, coeffs)] if nf.shape[0] else [] for _ in range(nrhs)]
    eq = eqlb.FluxEqlbSE(k, m, F, G, equilibrate_stress=stress, host_pipeline=False)
    eq.set_boundary_conditions([prime] * nrhs, bcs)
    eq.equilibrate_fluxes()
    sub, cells, nodes, facets, good_node, good_cell, good_fct = window(m, i0, j0, w)
    Gs = [g.reshape(m.ncell, -1)[cells].ravel() for g in G]
    Fs = [f.reshape(m.ncell, -1)[cells].ravel() for f in F]
    # boundary data of the window: the real boundary facets keep their type and data, the artificial ones are essential
    # for the primal problem (their patches are excluded from the comparison)
    sel = np.isin(facets, nf)
    sub_nf = np.nonzero(sel)[0].astype(np.int32)
    sub_coeffs = coeffs[np.searchsorted(nf, facets[sel])] if nf.shape[0] else np.zeros((0, k))
    sub_prime = np.setdiff1d(sub.bfct, sub_nf).astype(np.int32)
    sub_bcs = [[eqlb.fluxbc(sub_nf, sub_coeffs)] if sub_nf.shape[0] else [] for _ in range(nrhs)]
    bd = eqlb.boundarydata(sub_bcs, sub, T, [sub_prime] * nrhs, stress)
    bc = po.BCData(bd.facet_type, bd.bflux, bd.local_fct_id, bd.node_on_stress_bnd)
    ref = cpu_impl().se_run(sub, T, bc, Gs, Fs, stress=stress)
    assert good_cell.sum() > 100
    for r in range(nrhs):
        a = eq.list_flux[r].reshape(m.ncell, T.nrt)[cells[good_cell]]
        b = ref[r].reshape(sub.ncell, T.nrt)[good_cell]
        assert np.abs(a - b).max() < 1e-10 * np.abs(b).max()
