"""The compiled pybind11 module `dolfinx_eqlb_b200.cpp` (csrc/wrappers.cpp): successor of the reference's
`dolfinx_eqlb.cpp` (`python/dolfinx_eqlb/wrappers.cpp`) with the same Python-visible names and argument order."""

import ctypes as C
import inspect

import numpy as np
import pytest

from common import PoissonCase, make_mesh


def cpp_mesh(cpp, m):
    return cpp.Mesh(m.x, m.cell_node, m.cell_fct, m.fct_node, m.fct_cell_off, m.fct_cell, m.node_cell_off, m.node_cell,
                    m.node_fct_off, m.node_fct, m.fct_perms, m.cell_perm_info)


# names and argument order of python/dolfinx_eqlb/wrappers.cpp:54-137, 144-256
REFERENCE_SIGNATURES = {
    "reconstruct_fluxes_minimisation": ["a", "l_pen", "l", "flux_hdiv", "boundary_data"],
    "reconstruct_fluxes_semiexplt": ["flux_hdiv", "flux_dg", "rhs_dg", "boundary_data", "reconstruct_stress"],
    "reconstruct_fluxes_semiexplt_with_kornconst": ["flux_hdiv", "flux_dg", "rhs_dg", "boundary_data", "reconstruct_stress",
                                                    "cells_kornconst"],
    "local_solver_lu": ["solution", "a", "l"],
    "local_solver_cholesky": ["solution", "a", "l"],
    "local_solver_cg": ["solution", "a", "l"],
}


def test_module_exports_reference_names_and_argument_order():
    from dolfinx_eqlb_b200 import cpp

    for name, args in REFERENCE_SIGNATURES.items():
        doc = getattr(cpp, name).__doc__.splitlines()[0]
        got = [a.split(":")[0].strip() for a in doc[doc.index("(") + 1 : doc.rindex(")")].split(", ")]
        assert got == args, (name, got)
    for cls in ("FluxBC", "BoundaryData"):
        assert inspect.isclass(getattr(cpp, cls))
    doc = cpp.BoundaryData.__init__.__doc__
    for a in ["list_of_bcs", "list_of_boundary_fluxes", "V_flux_hdiv", "rtflux_is_custom", "quadrature_degree", "list_bfcts_prime",
              "reconstruct_stress"]:
        assert a in doc
    doc = cpp.FluxBC.__init__.__doc__
    for a in ["function_space", "facets", "pointer_boundary_kernel", "nevals_per_fct", "quadrature_degree", "coefficients",
              "position_of_coefficients", "constants"]:
        assert a in doc


def test_input_errors_have_the_reference_texts():
    from dolfinx_eqlb_b200 import cpp

    m = make_mesh("crossed", 2, None)
    cm = cpp_mesh(cpp, m)
    V = cpp.FunctionSpace(cm, "DRT", 2)
    Vg, Vf = cpp.FunctionSpace(cm, "DG", 1, 2), cpp.FunctionSpace(cm, "DG", 1)
    bd = cpp.BoundaryData([[]], [cpp.Function(V)], V, True, 2, [m.boundary_facets([1, 2, 3, 4])], False)
    with pytest.raises(RuntimeError, match="Input sizes does not match"):
        cpp.reconstruct_fluxes_semiexplt([cpp.Function(V), cpp.Function(V)], [cpp.Function(Vg)], [cpp.Function(Vf)], bd, False)
    with pytest.raises(RuntimeError, match="Wrong polynomial degree"):
        cpp.reconstruct_fluxes_semiexplt([cpp.Function(V)], [cpp.Function(cpp.FunctionSpace(cm, "DG", 2, 2))],
                                         [cpp.Function(cpp.FunctionSpace(cm, "DG", 2))], bd, False)
    with pytest.raises(RuntimeError, match="Specify all rows of stress tensor"):
        cpp.reconstruct_fluxes_semiexplt([cpp.Function(V)], [cpp.Function(Vg)], [cpp.Function(Vf)], bd, True)
    with pytest.raises(RuntimeError, match="Size of input data does not match"):
        cpp.BoundaryData([[]], [], V, True, 2, [m.boundary_facets([1])], False)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 2, 3])
def test_semiexplt_through_the_module(k):
    """SE through the compiled module == oracle; repeated calls reuse the cached device handle and accumulate"""
    from dolfinx_eqlb_b200 import cpp
    from oracle import pyoracle as po

    m = make_mesh("crossed", 5, 3, perturb=0.2)
    case = PoissonCase(m, k, [[1, 4], [2]], seed=3)
    ref = po.se_run(m, case.T, case.oracle_bc(), case.G, case.F)
    cm = cpp_mesh(cpp, m)
    V = cpp.FunctionSpace(cm, "DRT", k)
    Vg, Vf = cpp.FunctionSpace(cm, "DG", k - 1, 2), cpp.FunctionSpace(cm, "DG", k - 1)
    flux = [cpp.Function(V) for _ in range(2)]
    G = [cpp.Function(Vg, np.array(g)) for g in case.G]
    F = [cpp.Function(Vf, np.array(f)) for f in case.F]
    bcs = [[cpp.FluxBC(V, bc.facets, bc.coeffs) for bc in row] for row in case.list_bcs]
    bfun = [cpp.Function(V) for _ in range(2)]
    bd = cpp.BoundaryData(bcs, bfun, V, True, 2 * (k - 1), [np.asarray(p, dtype=np.int32) for p in case.list_bfct_prime], False)
    cpp.reconstruct_fluxes_semiexplt(flux, G, F, bd, False)
    assert cm.num_cached_handles == 1
    for r in range(2):
        assert np.abs(flux[r].x - ref[r]).max() < 1e-10 * np.abs(ref[r]).max()
        want = case.bdata.bflux[r] if case.bdata.bflux[r] is not None else 0.0
        assert np.abs(bfun[r].x - want).max() < 1e-13  # boundary DOFs land in the functions handed to BoundaryData
    cpp.reconstruct_fluxes_semiexplt(flux, G, F, bd, False)  # accumulates like the reference (`+=`)
    assert cm.num_cached_handles == 1
    for r in range(2):
        assert np.abs(flux[r].x - 2 * ref[r]).max() < 1e-10 * np.abs(ref[r]).max()


@pytest.mark.gpu
def test_minimisation_stress_korn_and_local_solver_through_the_module():
    from dolfinx_eqlb_b200 import cpp, eqlb
    from oracle import pyoracle as po
    from test_gpu_stress import elasticity_case

    k = 2
    m = make_mesh("crossed", 5, 2, perturb=0.2)
    cm = cpp_mesh(cpp, m)
    # constrained minimisation (EV)
    case = PoissonCase(m, k, [[]], seed=4, hom=True)
    ref = po.ev_run(m, case.T, case.oracle_bc(), case.G, case.F)
    Vrt, Vmix = cpp.FunctionSpace(cm, "RT", k), cpp.FunctionSpace(cm, "RT", k)
    Vg, Vf = cpp.FunctionSpace(cm, "DG", k - 1, 2), cpp.FunctionSpace(cm, "DG", k - 1)
    flux = [cpp.Function(Vrt)]
    l = [cpp.Form.ev_linear(Vmix, cpp.Function(Vg, np.array(case.G[0])), cpp.Function(Vf, np.array(case.F[0])))]
    bd = cpp.BoundaryData([[]], [cpp.Function(Vrt)], Vmix, False, 2, [np.asarray(case.list_bfct_prime[0], dtype=np.int32)], False)
    cpp.reconstruct_fluxes_minimisation(cpp.Form.ev_bilinear(Vmix), cpp.Form.ev_penalty(Vf), l, flux, bd)
    assert np.abs(flux[0].x - ref[0]).max() < 1e-10 * np.abs(ref[0]).max()
    # stress rows + Korn constants
    T, G, f, bfp, bcs, neu = elasticity_case(m, k, [1], seed=3, galerkin=False)
    hb = eqlb.boundarydata(bcs, m, T, bfp, True)
    sref, kref = po.se_run(m, T, po.BCData(hb.facet_type, hb.bflux, hb.local_fct_id, hb.node_on_stress_bnd), G, f, stress=True, korn=True)
    V = cpp.FunctionSpace(cm, "DRT", k)
    sig = [cpp.Function(V) for _ in range(2)]
    korn = cpp.Function(cpp.FunctionSpace(cm, "DG", 0))
    cbcs = [[cpp.FluxBC(V, bc.facets, bc.coeffs) for bc in row] for row in bcs]
    bd2 = cpp.BoundaryData(cbcs, [cpp.Function(V) for _ in range(2)], V, True, 2, [np.asarray(p, dtype=np.int32) for p in bfp], True)
    cpp.reconstruct_fluxes_semiexplt_with_kornconst(sig, [cpp.Function(Vg, np.array(g)) for g in G],
                                                    [cpp.Function(Vf, np.array(x)) for x in f], bd2, True, korn)
    for r in range(2):
        assert np.abs(sig[r].x - sref[r]).max() < 1e-10 * np.abs(sref[r]).max()
    assert np.abs(korn.x - kref).max() < 1e-12 * np.abs(kref).max()
    # local projection (local_solver_cholesky with the mass form)
    rng = np.random.default_rng(0)
    q = rng.standard_normal(m.ncell * T.nq)
    want = po.local_project(m, T, [q])[0]
    sol = [cpp.Function(Vf)]
    cpp.local_solver_cholesky(sol, cpp.Form.mass(Vf), [cpp.Form.projection_rhs(Vf, q)])
    assert np.abs(sol[0].x - want).max() < 1e-12 * np.abs(want).max()


@pytest.mark.gpu
def test_fluxbc_with_compiled_kernel_pointer():
    """the reference's FluxBC signature: a compiled boundary kernel handed over as an integer address"""
    from dolfinx_eqlb_b200 import cpp, eqlb

    k = 2
    m = make_mesh("crossed", 4, 3, perturb=0.2)
    T = __import__("dolfinx_eqlb_b200.tables", fromlist=["x"]).make_tables(k)
    fcts = m.boundary_facets([1, 4])
    rng = np.random.default_rng(2)
    coeffs = 2.0 * (rng.random((fcts.shape[0], k)) + 0.1)
    want = eqlb.boundarydata([[eqlb.fluxbc(fcts, coeffs)]], m, T, [m.boundary_facets([2, 3])], False)
    # traction = c0 + c1 s per facet, c = coefficient function (DG-like: per cell [local facet][2])
    cell = m.fct_cell[m.fct_cell_off[fcts]]
    lf = np.argmax(m.cell_fct[cell] == fcts[:, None], axis=1)
    s_ip = 0.5 + 0.5 * np.polynomial.legendre.leggauss(k + 1)[0]
    KFUN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                       C.POINTER(C.c_int), C.POINTER(C.c_uint8))

    def kern(vals, w, c, x, e, q):
        for f in range(3):
            for l_, s in enumerate(s_ip):
                vals[f * (k + 1) + l_] = w[2 * f] + w[2 * f + 1] * s

    cb = KFUN(kern)
    cm = cpp_mesh(cpp, m)
    V = cpp.FunctionSpace(cm, "DRT", k)
    Vc = cpp.FunctionSpace(cm, "DG", 2)  # 6 dofs per cell = [3 facets][2 coefficients]
    wv = np.zeros(m.ncell * 6)
    for i in range(fcts.shape[0]):
        wv[cell[i] * 6 + 2 * lf[i] : cell[i] * 6 + 2 * lf[i] + 2] = coeffs[i]
    bc = cpp.FluxBC(V, fcts, C.cast(cb, C.c_void_p).value, k + 1, [cpp.Function(Vc, wv)], [0], [])
    bfun = [cpp.Function(V)]
    bd = cpp.BoundaryData([[bc]], bfun, V, True, 2, [m.boundary_facets([2, 3])], False)
    Vg, Vf = cpp.FunctionSpace(cm, "DG", 1, 2), cpp.FunctionSpace(cm, "DG", 1)
    cpp.reconstruct_fluxes_semiexplt([cpp.Function(V)], [cpp.Function(Vg)], [cpp.Function(Vf)], bd, False)
    assert np.abs(bfun[0].x - want.bflux[0]).max() < 1e-12 * np.abs(want.bflux[0]).max()
