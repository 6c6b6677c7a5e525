"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

ctypes loader of `oracle/_ref/libeqlb_ref.so`: the reference's own C++ sources
(`/root/reference/cpp/dolfinx_eqlb`) compiled UNCHANGED against the stand-in headers of
`oracle/ref_shim` (recipe: `make -C oracle ref`).  The library is built where
`/root/reference` exists (this container; `__graft_entry__.build()`), is git-ignored and
travels to the GPU box with the snapshot.  Its flux element comes from the committed
fixture `tests/golden/ref_rt_element.npz` (the reference's `e_raviart_thomas.py`
executed by `tests/golden/make_ref_element.py`).

Only tests/ and bench.py's reference arm may import this module.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from dolfinx_eqlb_b200.cabi import EqlbMesh, PackedMesh, c_double_p, c_int8_p, c_int32_p, c_uint8_p, ptr_array

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libeqlb_ref.so")
REF_SRC = "/root/reference/cpp/dolfinx_eqlb"
ELEMENT_NPZ = os.path.join(_HERE, "..", "tests", "golden", "ref_rt_element.npz")
_lib = None


class RefElement(C.Structure):
    _fields_ = [
        ("k", C.c_int32), ("nrt", C.c_int32), ("nmono", C.c_int32), ("coef", c_double_p),
        ("npts", C.c_int32), ("mcols", C.c_int32), ("X", c_double_p), ("M", c_double_p), ("p", C.c_int32),
    ]


def available() -> bool:
    """True when the library exists or can be built here."""
    return os.path.exists(LIB) or os.path.isdir(REF_SRC)


def build(force=False):
    if not os.path.isdir(REF_SRC):
        if os.path.exists(LIB):
            return
        raise RuntimeError("oracle/_ref: /root/reference is absent and no prebuilt library travelled with the snapshot")
    srcs = [os.path.join(_HERE, "ref_driver.cpp")]
    for root, _, files in os.walk(os.path.join(_HERE, "ref_shim")):
        srcs += [os.path.join(root, f) for f in files if not f.endswith(".py")]
    if (not force) and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.ref_last_error.restype = C.c_char_p
        L.ref_se_run.restype = C.c_int
        L.ref_se_run.argtypes = [
            C.POINTER(EqlbMesh), C.POINTER(RefElement), C.c_int, c_int8_p, C.POINTER(c_double_p), c_int8_p, c_int8_p,
            C.c_int, C.POINTER(c_double_p), C.POINTER(c_double_p), C.POINTER(c_double_p), c_double_p,
        ]
        L.ref_se_patch_maps.restype = C.c_int
        L.ref_se_patch_maps.argtypes = [
            C.POINTER(EqlbMesh), C.POINTER(RefElement), C.c_int, c_int8_p, c_int8_p, C.c_int, C.c_int,
            c_int32_p, c_int32_p, c_int32_p, c_int8_p, c_int8_p, c_int8_p, c_uint8_p, c_uint8_p, c_int32_p, c_int32_p, c_int8_p,
        ]
        L.ref_set_flip_orientations.argtypes = [C.c_int]
        L.ref_ev_patch_maps.restype = C.c_int
        L.ref_ev_patch_maps.argtypes = [
            C.POINTER(EqlbMesh), C.POINTER(RefElement), C.c_int, c_int8_p, C.c_int,
            c_int32_p, c_int32_p, c_int32_p, c_int8_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p,
        ]
        L.ref_ev_run.restype = C.c_int
        L.ref_ev_run.argtypes = [
            C.POINTER(EqlbMesh), C.POINTER(RefElement), C.c_int, c_int8_p, C.POINTER(c_double_p),
            C.POINTER(c_double_p), C.POINTER(c_double_p), C.POINTER(c_double_p),
        ]
        L.ref_boundary_data.restype = C.c_int
        L.ref_boundary_data.argtypes = [
            C.POINTER(EqlbMesh), C.POINTER(RefElement), C.c_int, C.c_int, C.c_int, c_int32_p, C.POINTER(c_int32_p), c_int32_p,
            C.POINTER(c_int32_p), C.c_int, C.POINTER(c_double_p), c_int8_p, C.POINTER(c_double_p), c_int8_p,
        ]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().ref_last_error().decode())


def _i8(a):
    return a.ctypes.data_as(c_int8_p) if a is not None else c_int8_p()


class PackedElement:
    """`ref_element` of flux degree k, data degree p, from the committed fixture."""

    def __init__(self, k, p, continuous=False):
        z = np.load(ELEMENT_NPZ)
        tag = f"{k}c" if continuous else f"{k}"
        self.coef = np.ascontiguousarray(z["coef_" + tag], dtype=np.float64)
        self.X = np.ascontiguousarray(z["X_" + tag], dtype=np.float64)
        self.M = np.ascontiguousarray(z["M_" + tag], dtype=np.float64)
        nrt, _, nmono = self.coef.shape
        self.struct = RefElement(
            k, nrt, nmono, self.coef.ctypes.data_as(c_double_p), self.X.shape[0], self.M.shape[1],
            self.X.ctypes.data_as(c_double_p), self.M.ctypes.data_as(c_double_p), p,
        )


def se_run(mesh, tables, bc, G, F, stress=False, korn=False, sigma0=None):
    """`reconstruct_fluxes_semiexplt[_with_kornconst]` of the reference -> DRT vectors."""
    pm, pe = PackedMesh(mesh, tables.ndg), PackedElement(tables.k, tables.p)
    nrhs = bc.nrhs
    G = [np.ascontiguousarray(g, dtype=np.float64) for g in G]
    F = [np.ascontiguousarray(f, dtype=np.float64) for f in F]
    sig = [np.zeros(mesh.ncell * tables.nrt) if sigma0 is None else np.array(sigma0[i], dtype=np.float64) for i in range(nrhs)]
    nob = bc.node_on_stress_bnd
    if stress and nob is None:
        nob = np.zeros(mesh.nnode, dtype=np.int8)
    kc = np.zeros(mesh.ncell) if korn else None
    rc = lib().ref_se_run(
        C.byref(pm.struct), C.byref(pe.struct), nrhs, _i8(bc.facet_type), ptr_array(bc.bflux), _i8(bc.local_fct_id), _i8(nob),
        int(stress), ptr_array(G), ptr_array(F), ptr_array(sig), kc.ctypes.data_as(c_double_p) if korn else c_double_p(),
    )
    _check(rc)
    return (sig, kc) if korn else sig


def se_patch_maps(mesh, tables, bc, stress=False):
    """Same dictionary as `pyoracle.se_patch_maps`, produced by the reference's classes."""
    pm, pe = PackedMesh(mesh, tables.ndg), PackedElement(tables.k, tables.p)
    nrhs = bc.nrhs
    ncmax = int(np.diff(mesh.node_cell_off).max())
    npatch = mesh.nnode
    k = tables.k
    ndpc = 2 * k + tables.nadd + tables.ndiv + (3 if stress else 0)
    hzmax = 1 + (k - 1) * (ncmax + 1) + tables.nadd * ncmax
    out = dict(
        ncells=np.zeros(npatch, np.int32),
        cells=np.zeros((npatch, ncmax + 2), np.int32),
        fcts=np.zeros((npatch, ncmax + 2), np.int32),
        inodes_local=np.zeros((npatch, ncmax + 2), np.int8),
        fcts_local=np.zeros((npatch, 2 * (ncmax + 1)), np.int8),
        type=np.zeros((npatch, nrhs), np.int8),
        reversed=np.zeros((npatch, ncmax, 2), np.uint8),
        reversion=np.zeros((npatch, nrhs), np.uint8),
        dofmap=np.zeros((npatch, 4, ncmax + 2, ndpc), np.int32),
        projflux_fct=np.zeros((npatch, ncmax + 1, 2 * tables.ndg_fct), np.int32),
        bmarkers=np.zeros((npatch, nrhs, hzmax), np.int8),
    )
    nob = bc.node_on_stress_bnd
    if stress and nob is None:
        nob = np.zeros(mesh.nnode, dtype=np.int8)
    rc = lib().ref_se_patch_maps(
        C.byref(pm.struct), C.byref(pe.struct), nrhs, _i8(bc.facet_type), _i8(nob), int(stress), ncmax,
        out["ncells"].ctypes.data_as(c_int32_p), out["cells"].ctypes.data_as(c_int32_p), out["fcts"].ctypes.data_as(c_int32_p),
        _i8(out["inodes_local"]), _i8(out["fcts_local"]), _i8(out["type"]), out["reversed"].ctypes.data_as(c_uint8_p),
        out["reversion"].ctypes.data_as(c_uint8_p), out["dofmap"].ctypes.data_as(c_int32_p),
        out["projflux_fct"].ctypes.data_as(c_int32_p), _i8(out["bmarkers"]),
    )
    _check(rc)
    out["ncmax"] = ncmax
    return out


def ev_patch_maps(mesh, tables, bc, node):
    """Same dictionary as `pyoracle.ev_patch_maps`, produced by the reference's ev::Patch."""
    pm, pe = PackedMesh(mesh, tables.ndg), PackedElement(tables.k, tables.p, continuous=True)
    ncmax = int(np.diff(mesh.node_cell_off).max())
    k = tables.k
    nz = tables.nrt + tables.ndg - k
    nc = np.zeros(1, np.int32)
    out = dict(
        cells=np.full(ncmax, -1, np.int32), fcts=np.full(ncmax + 1, -1, np.int32), inodes_local=np.full(ncmax, -1, np.int8),
        dofs_elmt=np.full(ncmax * nz, -1, np.int32), dofs_patch=np.full(ncmax * nz, -1, np.int32),
        dofs_global=np.full(ncmax * nz, -1, np.int32),
        list_patch=np.full(ncmax * (k * k - k) + (ncmax + 1) * k, -1, np.int32),
        list_global=np.full(ncmax * (k * k - k) + (ncmax + 1) * k, -1, np.int32),
    )
    rc = lib().ref_ev_patch_maps(
        C.byref(pm.struct), C.byref(pe.struct), bc.nrhs, _i8(bc.facet_type), int(node), nc.ctypes.data_as(c_int32_p),
        out["cells"].ctypes.data_as(c_int32_p), out["fcts"].ctypes.data_as(c_int32_p), _i8(out["inodes_local"]),
        out["dofs_elmt"].ctypes.data_as(c_int32_p), out["dofs_patch"].ctypes.data_as(c_int32_p),
        out["dofs_global"].ctypes.data_as(c_int32_p), out["list_patch"].ctypes.data_as(c_int32_p),
        out["list_global"].ctypes.data_as(c_int32_p),
    )
    _check(rc)
    out["ncells"] = int(nc[0])
    return out


def ev_run(mesh, tables, bc, G, F, sigma0=None):
    """`reconstruct_fluxes_minimisation` of the reference (ev::reconstruction with its own patch
    assembly, lifting, dense partial-pivot LU and scatter) -> conforming hierarchic-RT vectors."""
    pm, pe = PackedMesh(mesh, tables.ndg), PackedElement(tables.k, tables.p, continuous=True)
    nrhs = bc.nrhs
    G = [np.ascontiguousarray(g, dtype=np.float64) for g in G]
    F = [np.ascontiguousarray(f, dtype=np.float64) for f in F]
    k = tables.k
    n = mesh.nfct * k + mesh.ncell * (k * k - k)
    sig = [np.zeros(n) if sigma0 is None else np.array(sigma0[i], dtype=np.float64) for i in range(nrhs)]
    rc = lib().ref_ev_run(
        C.byref(pm.struct), C.byref(pe.struct), nrhs, _i8(bc.facet_type), ptr_array(bc.bflux), ptr_array(G), ptr_array(F),
        ptr_array(sig),
    )
    _check(rc)
    return sig


def boundary_data(mesh, tables, list_bcs, list_bfct_prime, stress=False, qdegree_proj=-1):
    """The reference's `base::BoundaryData` constructor with polynomial FluxBC kernels
    (`list_bcs[r]` = list of objects with `.facets`, `.coeffs`).  qdegree_proj < 0: interpolation branch,
    else the facet-local projection branch with that quadrature degree.
    Returns (facet_type [nrhs][nfct], bflux list, node_on_stress_bnd or None)."""
    pm, pe = PackedMesh(mesh, tables.ndg), PackedElement(tables.k, tables.p)
    nrhs = len(list_bcs)
    fc, cf = [], []
    ncoef = 1
    for bcs in list_bcs:
        for bc in bcs:
            ncoef = max(ncoef, bc.coeffs.shape[1])
    for bcs in list_bcs:
        f = np.concatenate([bc.facets for bc in bcs]).astype(np.int32) if bcs else np.zeros(0, np.int32)
        c = np.zeros((f.shape[0], ncoef))
        o = 0
        for bc in bcs:
            c[o : o + bc.facets.shape[0], : bc.coeffs.shape[1]] = bc.coeffs
            o += bc.facets.shape[0]
        fc.append(np.ascontiguousarray(f))
        cf.append(np.ascontiguousarray(c))
    pr_ = [np.ascontiguousarray(p, dtype=np.int32) for p in list_bfct_prime]
    i32 = lambda xs: (C.c_int32 * len(xs))(*xs)
    ft = np.zeros((nrhs, mesh.nfct), np.int8)
    bv = [np.zeros(mesh.ncell * tables.nrt) for _ in range(nrhs)]
    nob = np.zeros(mesh.nnode, np.int8)
    rc = lib().ref_boundary_data(
        C.byref(pm.struct), C.byref(pe.struct), nrhs, int(stress), int(qdegree_proj), i32([p.shape[0] for p in pr_]),
        ptr_array(pr_, C.c_int32), i32([f.shape[0] for f in fc]), ptr_array(fc, C.c_int32), ncoef, ptr_array(cf),
        _i8(ft), ptr_array(bv), _i8(nob),
    )
    _check(rc)
    return ft, bv, (nob if stress else None)


def local_solver(mesh, tables, qvals, solver="cholesky"):
    """The reference's `base::local_solver_{lu,cholesky,cg}` (`base/local_solver.hpp`) on the fixed projection forms
    of `lsolver/projection.py:17-77`: a = (u, v) on DG_p, l_i = (data_i, v), data_i at the cell quadrature points
    (`qvals[i][cell*nq + q]`).  Returns the DG_p coefficient vectors."""
    from dolfinx_eqlb_b200.cabi import EqlbTables, PackedTables

    pm, pt = PackedMesh(mesh, tables.ndg), PackedTables(tables)
    L = lib()
    L.ref_local_solver.restype = C.c_int
    L.ref_local_solver.argtypes = [C.POINTER(EqlbMesh), C.POINTER(EqlbTables), C.c_int, C.c_int, C.POINTER(c_double_p), C.POINTER(c_double_p)]
    qv = [np.ascontiguousarray(q, dtype=np.float64) for q in qvals]
    out = [np.zeros(mesh.ncell * tables.ndg) for _ in qv]
    _check(L.ref_local_solver(C.byref(pm.struct), C.byref(pt.struct), len(qv), {"lu": 0, "cholesky": 1, "cg": 2}[solver],
                              ptr_array(qv), ptr_array(out)))
    return out
