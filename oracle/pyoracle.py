"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

ctypes loader of the CPU oracle (`oracle/liboracle.so`, built by
`oracle/Makefile`).  Only tests/, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of bench.py may import this module.
Parity pinned: checked against the reference's own sources compiled unchanged (`oracle/_ref`, loaded by
`oracle/pyref.py`) in tests/test_ref_pinning.py, and against the golden vectors generated from it.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from dolfinx_eqlb_b200.cabi import (
    EqlbMesh,
    EqlbTables,
    PackedMesh,
    PackedTables,
    c_double_p,
    c_int8_p,
    c_int32_p,
    c_uint8_p,
    ptr_array,
)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", ".inc"))]
    srcs.append(os.path.join(_HERE, "..", "include", "eqlb_b200.h"))
    if (not force) and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_se_run.restype = C.c_int
        L.oracle_se_run.argtypes = [
            C.POINTER(EqlbMesh), C.POINTER(EqlbTables), C.c_int, c_int8_p, C.POINTER(c_double_p), c_int8_p, c_int8_p,
            C.c_int, C.POINTER(c_double_p), C.POINTER(c_double_p), C.POINTER(c_double_p), c_double_p,
        ]
        L.oracle_se_patch_maps.restype = C.c_int
        L.oracle_se_patch_maps.argtypes = [
            C.POINTER(EqlbMesh), C.POINTER(EqlbTables), C.c_int, c_int8_p, c_int8_p, C.c_int, C.c_int,
            c_int32_p, c_int32_p, c_int32_p, c_int8_p, c_int8_p, c_int8_p, c_uint8_p, c_uint8_p, c_int32_p, c_int32_p, c_int8_p,
        ]
        L.oracle_ev_run.restype = C.c_int
        L.oracle_ev_run.argtypes = [
            C.POINTER(EqlbMesh), C.POINTER(EqlbTables), C.c_int, c_int8_p, C.POINTER(c_double_p),
            C.POINTER(c_double_p), C.POINTER(c_double_p), C.POINTER(c_double_p),
        ]
        L.oracle_ev_patch_maps.restype = C.c_int
        L.oracle_ev_patch_maps.argtypes = [
            C.POINTER(EqlbMesh), C.POINTER(EqlbTables), C.c_int, c_int8_p, C.c_int,
            c_int32_p, c_int32_p, c_int32_p, c_int8_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p,
        ]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().oracle_last_error().decode())


def _i8(a):
    return a.ctypes.data_as(c_int8_p) if a is not None else c_int8_p()


class BCData:
    """What the hot path reads from `base::BoundaryData` (host arrays)."""

    def __init__(self, facet_type, bflux=None, local_fct_id=None, node_on_stress_bnd=None):
        self.facet_type = np.ascontiguousarray(facet_type, dtype=np.int8)  # [nrhs][nfct]
        self.nrhs = self.facet_type.shape[0]
        self.bflux = bflux if bflux is not None else [None] * self.nrhs
        self.local_fct_id = None if local_fct_id is None else np.ascontiguousarray(local_fct_id, dtype=np.int8)
        self.node_on_stress_bnd = None if node_on_stress_bnd is None else np.ascontiguousarray(node_on_stress_bnd, dtype=np.int8)


def se_run(mesh, tables, bc: BCData, G, F, stress=False, korn=False, sigma0=None, node_owned=None):
    """oracle of `reconstruct_fluxes_semiexplt` -> list of DRT vectors [ncell*nrt]."""
    pm, pt = PackedMesh(mesh, tables.ndg, node_owned), PackedTables(tables)
    nrhs = bc.nrhs
    G = [np.ascontiguousarray(g, dtype=np.float64) for g in G]
    F = [np.ascontiguousarray(f, dtype=np.float64) for f in F]
    sig = [np.zeros(mesh.ncell * tables.nrt) if sigma0 is None else np.array(sigma0[i], dtype=np.float64) for i in range(nrhs)]
    lfi = bc.local_fct_id if bc.local_fct_id is not None else np.zeros(mesh.nfct, dtype=np.int8)
    nob = bc.node_on_stress_bnd
    if stress and nob is None:
        nob = np.zeros(mesh.nnode, dtype=np.int8)
    kc = np.zeros(mesh.ncell) if korn else None
    rc = lib().oracle_se_run(
        C.byref(pm.struct), C.byref(pt.struct), nrhs, _i8(bc.facet_type), ptr_array(bc.bflux), _i8(lfi), _i8(nob),
        int(stress), ptr_array(G), ptr_array(F), ptr_array(sig), kc.ctypes.data_as(c_double_p) if korn else c_double_p(),
    )
    _check(rc)
    return (sig, kc) if korn else sig


def se_patch_maps(mesh, tables, bc: BCData, stress=False):
    pm, pt = PackedMesh(mesh, tables.ndg), PackedTables(tables)
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
    rc = lib().oracle_se_patch_maps(
        C.byref(pm.struct), C.byref(pt.struct), nrhs, _i8(bc.facet_type), _i8(nob), int(stress), ncmax,
        out["ncells"].ctypes.data_as(c_int32_p), out["cells"].ctypes.data_as(c_int32_p), out["fcts"].ctypes.data_as(c_int32_p),
        _i8(out["inodes_local"]), _i8(out["fcts_local"]), _i8(out["type"]), out["reversed"].ctypes.data_as(c_uint8_p),
        out["reversion"].ctypes.data_as(c_uint8_p), out["dofmap"].ctypes.data_as(c_int32_p),
        out["projflux_fct"].ctypes.data_as(c_int32_p), _i8(out["bmarkers"]),
    )
    _check(rc)
    out["ncmax"] = ncmax
    return out


def ev_ndofs(mesh, tables):
    k = tables.k
    return mesh.nfct * k + mesh.ncell * (k * k - k)


def ev_run(mesh, tables, bc: BCData, G, F, sigma0=None, node_owned=None):
    """oracle of `reconstruct_fluxes_minimisation` -> conforming hierarchic-RT vectors."""
    pm, pt = PackedMesh(mesh, tables.ndg, node_owned), PackedTables(tables)
    nrhs = bc.nrhs
    G = [np.ascontiguousarray(g, dtype=np.float64) for g in G]
    F = [np.ascontiguousarray(f, dtype=np.float64) for f in F]
    n = ev_ndofs(mesh, tables)
    sig = [np.zeros(n) if sigma0 is None else np.array(sigma0[i], dtype=np.float64) for i in range(nrhs)]
    rc = lib().oracle_ev_run(
        C.byref(pm.struct), C.byref(pt.struct), nrhs, _i8(bc.facet_type), ptr_array(bc.bflux), ptr_array(G), ptr_array(F),
        ptr_array(sig),
    )
    _check(rc)
    return sig


def ev_patch_maps(mesh, tables, bc: BCData, node):
    pm, pt = PackedMesh(mesh, tables.ndg), PackedTables(tables)
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
    rc = lib().oracle_ev_patch_maps(
        C.byref(pm.struct), C.byref(pt.struct), bc.nrhs, _i8(bc.facet_type), int(node), nc.ctypes.data_as(c_int32_p),
        out["cells"].ctypes.data_as(c_int32_p), out["fcts"].ctypes.data_as(c_int32_p), _i8(out["inodes_local"]),
        out["dofs_elmt"].ctypes.data_as(c_int32_p), out["dofs_patch"].ctypes.data_as(c_int32_p),
        out["dofs_global"].ctypes.data_as(c_int32_p), out["list_patch"].ctypes.data_as(c_int32_p),
        out["list_global"].ctypes.data_as(c_int32_p),
    )
    _check(rc)
    out["ncells"] = int(nc[0])
    return out


def local_project(mesh, tables, qvals):
    """oracle of `local_projection` (lsolver/projection.py:17-77) into DG_p."""
    pm, pt = PackedMesh(mesh, tables.ndg), PackedTables(tables)
    L = lib()
    L.oracle_local_project.restype = C.c_int
    L.oracle_local_project.argtypes = [C.POINTER(EqlbMesh), C.POINTER(EqlbTables), C.c_int, C.POINTER(c_double_p), C.POINTER(c_double_p)]
    qv = [np.ascontiguousarray(q, dtype=np.float64) for q in qvals]
    out = [np.zeros(mesh.ncell * tables.ndg) for _ in qv]
    _check(L.oracle_local_project(C.byref(pm.struct), C.byref(pt.struct), len(qv), ptr_array(qv), ptr_array(out)))
    return out
