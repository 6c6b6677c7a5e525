"""Pinning of the oracle (and of `tables.py`) against the reference's OWN code.

`oracle/_ref/libeqlb_ref.so` = the reference's sources (`/root/reference/cpp/dolfinx_eqlb`:
se/Patch.cpp, se/KernelData.cpp, base/KernelData.cpp, base/BoundaryData.cpp, ev/Patch.cpp and
the header templates behind `se::reconstruction`) compiled unchanged against stand-in headers
(`oracle/ref_shim`, recipe `make -C oracle ref`).  The tests skip when neither the library
nor /root/reference is present.

`tests/golden/ref_rt_element.npz` = the reference's `e_raviart_thomas.py` executed
(tests/golden/make_ref_element.py): pins the hierarchic RT basis of `tables.py`.
"""

import os

import numpy as np
import pytest

import fem_mini as fm
from common import PoissonCase, make_mesh
from dolfinx_eqlb_b200 import eqlb
from dolfinx_eqlb_b200.tables import make_tables, p_eval
from oracle import pyref as pr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not pr.available(), reason="oracle/_ref not built and /root/reference absent")

MESHES = [("crossed", 2, None), ("crossed", 4, 3), ("randdiag", 5, 2)]
BC_SETS = [[1, 4], [1, 3], [2], [1, 3, 4]]  # test_fluxeqlb_multirhs.py:70
# traction layouts of test_stressqlb_bcond.py:167-190 (+ the pure-Dirichlet case)
STRESS_LAYOUTS = [[], [1], [2], [3], [4], [1, 2], [1, 3], [1, 4], [2, 3], [2, 4], [3, 4], [1, 2, 3], [2, 3, 4], [1, 3, 4],
                  [1, 2, 4]]


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_tables_match_reference_element(k):
    """RT basis of tables.py == dual basis defined by the reference's element file."""
    z = np.load(os.path.join(GOLD, "ref_rt_element.npz"))
    T = make_tables(k)
    rt = T.extra["rt_exact"]
    tab = np.array([[[p_eval(px, x, y), p_eval(py, x, y)] for (px, py) in rt] for (x, y) in z["samples"]])
    for tag in (f"{k}", f"{k}c"):
        assert np.abs(tab - z["tab_" + tag]).max() < 1e-13 * np.abs(tab).max()
    # facet interpolation matrix: M[f][j][d][n] of tables.py == rows of the reference's M
    X, M = z[f"X_{k}"], z[f"M_{k}"]
    nder = int(z[f"nder_{k}"])
    npts = X.shape[0]
    M4 = M.reshape(T.nrt, 2, npts, nder)
    for f in range(3):
        for j in range(k):
            got = M4[f * k + j, :, f * T.nqf : (f + 1) * T.nqf, 0]
            assert np.abs(got - T.M[f, j]).max() < 1e-14


@needs_ref
@pytest.mark.parametrize("kind,n,scramble", MESHES)
@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_oracle_matches_reference_se(kind, n, scramble, k):
    """integer maps bit-exact, flux DOFs <= 1e-11 relative: oracle vs the reference's se::reconstruction"""
    from oracle import pyoracle as po

    m = make_mesh(kind, n, scramble, perturb=0.2)
    case = PoissonCase(m, k, BC_SETS, seed=5, galerkin=(k < 4))
    bc = case.oracle_bc()
    a, b = po.se_patch_maps(m, case.T, bc), pr.se_patch_maps(m, case.T, bc)
    for key in a:
        if key != "ncmax":
            assert np.array_equal(a[key], b[key]), key
    so, sr = po.se_run(m, case.T, bc, case.G, case.F), pr.se_run(m, case.T, bc, case.G, case.F)
    for x, y in zip(so, sr):
        assert np.abs(x - y).max() < 1e-11 * np.abs(y).max()
    if k < 4:
        # the reference's own output satisfies its acceptance invariants
        for r in range(case.nrhs):
            assert fm.check_divergence(m, case.T, sr[r], case.G[r], case.F[r]) < 1e-12
            assert fm.check_jump(m, case.T, sr[r], case.G[r]) < 1e-11


@needs_ref
@pytest.mark.parametrize("kind,n,scramble", [("crossed", 4, None), ("crossed", 5, 3), ("randdiag", 6, 2)])
@pytest.mark.parametrize("k", [2, 3, 4])
def test_oracle_matches_reference_stress(kind, n, scramble, k):
    """weak symmetry, grouped corner patches and Korn constants on all traction layouts"""
    from oracle import pyoracle as po
    from test_gpu_stress import elasticity_case

    m = make_mesh(kind, n, scramble, perturb=0.2)
    for nsides in STRESS_LAYOUTS:
        T, G, f, bfp, bcs, neu = elasticity_case(m, k, nsides, seed=3, galerkin=False)
        bd = eqlb.boundarydata(bcs, m, T, bfp, True)
        bc = po.BCData(bd.facet_type, bd.bflux, bd.local_fct_id, bd.node_on_stress_bnd)
        so, ko = po.se_run(m, T, bc, G, f, stress=True, korn=True)
        sr, kr = pr.se_run(m, T, bc, G, f, stress=True, korn=True)
        for x, y in zip(so, sr):
            assert np.abs(x - y).max() < 1e-11 * np.abs(y).max(), nsides
        assert np.abs(ko - kr).max() < 1e-13 * np.abs(kr).max()
        a, b = po.se_patch_maps(m, T, bc, stress=True), pr.se_patch_maps(m, T, bc, stress=True)
        for key in a:
            if key != "ncmax":
                assert np.array_equal(a[key], b[key]), (nsides, key)


@needs_ref
@pytest.mark.parametrize("kind,n,scramble", MESHES)
@pytest.mark.parametrize("k", [1, 2, 3])
def test_oracle_matches_reference_ev_maps(kind, n, scramble, k):
    """EV ordering + sub-DOF maps, every patch: oracle vs the reference's ev::Patch"""
    from oracle import pyoracle as po

    m = make_mesh(kind, n, scramble, perturb=0.2)
    case = PoissonCase(m, k, [[1, 4], [2]], seed=5, galerkin=False)
    bc = case.oracle_bc()
    for z in range(m.nnode):
        a, b = po.ev_patch_maps(m, case.T, bc, z), pr.ev_patch_maps(m, case.T, bc, z)
        for key in a:
            assert np.array_equal(a[key], b[key]), (z, key)


@needs_ref
def test_reference_rejects_one_cell_patch():
    from dolfinx_eqlb_b200 import mesh as ms
    from oracle import pyoracle as po

    x = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    m = ms.build_topology(x, np.array([[0, 1, 3], [0, 2, 3]]))
    T = make_tables(1)
    ft = ms.facet_types(m, [1, 2, 3, 4], [])
    with pytest.raises(RuntimeError, match="has only 1 cells"):
        pr.se_run(m, T, po.BCData(ft[None, :]), [np.zeros(4)], [np.zeros(2)])


@needs_ref
def test_facet_orientation_convention_is_pinned_by_the_reference_invariants():
    """`basix::cell::facet_orientations` is third party; the stand-in returns {F, T, F}
    ("reference normal points outward").  With the complement the reference's own code
    violates its own acceptance invariants, so the convention is not a free choice."""
    m = make_mesh("crossed", 4, 3, perturb=0.2)
    case = PoissonCase(m, 2, [[1, 4]], seed=5)
    bc = case.oracle_bc()
    good = pr.se_run(m, case.T, bc, case.G, case.F)[0]
    pr.lib().ref_set_flip_orientations(1)
    try:
        bad = pr.se_run(m, case.T, bc, case.G, case.F)[0]
    finally:
        pr.lib().ref_set_flip_orientations(0)
    assert fm.check_divergence(m, case.T, good, case.G[0], case.F[0]) < 1e-12
    assert fm.check_jump(m, case.T, good, case.G[0]) < 1e-11
    assert fm.check_divergence(m, case.T, bad, case.G[0], case.F[0]) > 1e-3


@needs_ref
@pytest.mark.parametrize("kind,n,scramble,hom", [("crossed", 2, None, True), ("crossed", 4, 3, True), ("randdiag", 5, 2, True),
                                                 ("crossed", 4, None, False)])
@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("nsets", [[[]], [[1, 4], [1, 3]]])
def test_oracle_matches_reference_ev(kind, n, scramble, hom, k, nsets):
    """EV flux DOFs: oracle vs the reference's ev::reconstruction (its own patch loop, assemble_tangents,
    apply_lifting, dense partial-pivot LU, scatter; fixed forms and DOF transformations restated in
    oracle/ref_driver.cpp).  Inhomogeneous Neumann data only on meshes without reflected facets: the
    reference does not transform patch BCs to the global facet orientation (`base/BoundaryData.cpp:736-743`,
    DESIGN.md section 6) and tests homogeneous data only (`test_fluxeqlb_conditions.py:147,156`)."""
    from oracle import pyoracle as po

    m = make_mesh(kind, n, scramble, perturb=0.25)
    case = PoissonCase(m, k, nsets, seed=1, hom=hom)
    bc = case.oracle_bc()
    a, b = po.ev_run(m, case.T, bc, case.G, case.F), pr.ev_run(m, case.T, bc, case.G, case.F)
    for x, y in zip(a, b):
        assert np.abs(x - y).max() < 1e-11 * np.abs(y).max()
    # the reference's own EV output: divergence + conformity invariants
    for r in range(case.nrhs):
        s = fm.conforming_to_drt(m, case.T, b[r])
        z = np.zeros_like(case.G[r])
        assert fm.check_jump(m, case.T, s, z) < 1e-11


@needs_ref
@pytest.mark.parametrize("n,scramble", [(6, 3), (10, None)])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_oracle_matches_reference_on_unstructured_mesh(n, scramble, k):
    """Delaunay triangulation (stand-in for the reference's gmsh fixture `python/test/unit/utils.py:98-137`): vertex
    valences 2..9, reversed facets: SE maps bit-exact, SE / EV / stress DOFs <= 1e-11 against the reference"""
    from oracle import pyoracle as po
    from test_gpu_stress import elasticity_case

    m = make_mesh("delaunay", n, scramble)
    case = PoissonCase(m, k, [[1, 4], [2], []], seed=5, galerkin=False)
    bc = case.oracle_bc()
    a, b = po.se_patch_maps(m, case.T, bc), pr.se_patch_maps(m, case.T, bc)
    for key in a:
        if key != "ncmax":
            assert np.array_equal(a[key], b[key]), key
    for x, y in zip(po.se_run(m, case.T, bc, case.G, case.F), pr.se_run(m, case.T, bc, case.G, case.F)):
        assert np.abs(x - y).max() < 1e-11 * np.abs(y).max()
    case = PoissonCase(m, k, [[1, 4], []], seed=1, hom=True, galerkin=False)
    bc = case.oracle_bc()
    for x, y in zip(po.ev_run(m, case.T, bc, case.G, case.F), pr.ev_run(m, case.T, bc, case.G, case.F)):
        assert np.abs(x - y).max() < 1e-10 * np.abs(y).max()
    if k >= 2:
        T, G, f, bfp, bcs, neu = elasticity_case(m, k, [], seed=3, galerkin=False)
        bd = eqlb.boundarydata(bcs, m, T, bfp, True)
        bc = po.BCData(bd.facet_type, bd.bflux, bd.local_fct_id, bd.node_on_stress_bnd)
        so, ko = po.se_run(m, T, bc, G, f, stress=True, korn=True)
        sr, kr = pr.se_run(m, T, bc, G, f, stress=True, korn=True)
        for x, y in zip(so, sr):
            assert np.abs(x - y).max() < 1e-11 * np.abs(y).max()
        assert np.abs(ko - kr).max() < 1e-13 * np.abs(kr).max()


@needs_ref
@pytest.mark.parametrize("solver", ["lu", "cholesky", "cg"])
@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_oracle_projection_matches_reference_local_solver(k, solver):
    """a16: the cell-wise projector against the reference's `base::local_solver_{lu,cholesky,cg}` (its element loop,
    solver and scatter; the two FFCx cell kernels of the fixed projection forms restated in oracle/ref_driver.cpp)"""
    from dolfinx_eqlb_b200 import tables as tb
    from oracle import pyoracle as po

    m = make_mesh("crossed", 5, 3, perturb=0.25)
    T = tb.make_tables(k)
    rng = np.random.default_rng(k)
    qvals = [rng.standard_normal(m.ncell * T.nq) for _ in range(3)]
    a, b = po.local_project(m, T, qvals), pr.local_solver(m, T, qvals, solver)
    for x, y in zip(a, b):
        assert np.abs(x - y).max() < 1e-12 * np.abs(y).max()
