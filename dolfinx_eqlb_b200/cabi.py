"""ctypes mirror of `include/eqlb_b200.h` (structs + library loader).

This is the Python stand-in for the reference's pybind11 layer
(`python/dolfinx_eqlb/wrappers.cpp`): it turns mesh/table objects into the plain
C structs of the C ABI.  The product library is `csrc/libeqlb_b200.so`; there is
no fallback - loading fails loudly when the CUDA library has not been built.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libeqlb_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int8_p = C.POINTER(C.c_int8)
c_uint8_p = C.POINTER(C.c_uint8)
c_uint32_p = C.POINTER(C.c_uint32)


class EqlbMesh(C.Structure):
    _fields_ = [
        ("nnode", C.c_int32),
        ("ncell", C.c_int32),
        ("nfct", C.c_int32),
        ("x", c_double_p),
        ("cell_node", c_int32_p),
        ("cell_fct", c_int32_p),
        ("fct_node", c_int32_p),
        ("fct_cell_off", c_int32_p),
        ("fct_cell", c_int32_p),
        ("node_cell_off", c_int32_p),
        ("node_cell", c_int32_p),
        ("node_fct_off", c_int32_p),
        ("node_fct", c_int32_p),
        ("fct_perms", c_uint8_p),
        ("cell_perm_info", c_uint32_p),
        ("dg_dofmap", c_int32_p),
        ("node_owned", c_uint8_p),
    ]


_TABLE_DOUBLES = ["qpts", "qwts", "fpts_s", "fwts", "M", "rt_q", "rt_f", "dg_q", "dg_f", "hat_q", "hat_f", "trafo"]
_TABLE_INTS = ["fct_closure", "div_lm"]
_TABLE_REF = ["rt_mass", "fct_mom", "cell_mom_f", "cell_mom_g", "bc_mat", "rt_p1", "dg_mono", "hat_dg_rt", "mono_int",
              "rt_basix_fct", "rt_basix_int", "pk_grad_dg", "pk_to_dg", "pk_q", "pk_gq", "pk_hq", "rt_div_q"]


class EqlbTables(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in ["k", "p", "nrt", "ndg", "ndg_fct", "nq", "nqf", "ndiv", "nadd", "npk"]]
        + [(n, c_double_p) for n in _TABLE_DOUBLES]
        + [(n, c_int32_p) for n in _TABLE_INTS]
        + [(n, c_double_p) for n in _TABLE_REF]
    )


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class PackedMesh:
    """Owns contiguous copies of the mesh arrays and the C struct view."""

    def __init__(self, mesh, ndg, node_owned=None, identity_dg_as_null=False):
        from .mesh import dg_dofmap

        self.mesh = mesh
        self._keep = {}

        def keep(name, arr, dt):
            a = np.ascontiguousarray(arr, dtype=dt)
            self._keep[name] = a
            return a

        s = EqlbMesh()
        s.nnode, s.ncell, s.nfct = mesh.nnode, mesh.ncell, mesh.nfct
        s.x = _ptr(keep("x", mesh.x, np.float64), C.c_double)
        for name in ["cell_node", "cell_fct", "fct_node", "fct_cell_off", "fct_cell", "node_cell_off", "node_cell", "node_fct_off", "node_fct"]:
            setattr(s, name, _ptr(keep(name, getattr(mesh, name), np.int32), C.c_int32))
        s.fct_perms = _ptr(keep("fct_perms", mesh.fct_perms, np.uint8), C.c_uint8)
        s.cell_perm_info = _ptr(keep("cell_perm_info", mesh.cell_perm_info, np.uint32), C.c_uint32)
        # DG_p dofmap: the mirror's spaces use the DOLFINx layout cell*ndg + i.  The CUDA library assumes that layout for
        # a NULL pointer (`identity_dg_as_null`: saves building and checking a 12 M entry array per handle; a binding
        # passes V_dg.dofmap()->list()); the CPU oracles take the explicit map.
        custom = getattr(mesh, "dg_dofmap", None)
        if custom is not None:
            s.dg_dofmap = _ptr(keep("dg_dofmap", custom, np.int32), C.c_int32)
        elif not identity_dg_as_null:
            s.dg_dofmap = _ptr(keep("dg_dofmap", dg_dofmap(mesh.ncell, ndg), np.int32), C.c_int32)
        if node_owned is not None:
            s.node_owned = _ptr(keep("node_owned", node_owned, np.uint8), C.c_uint8)
        self.struct = s


class PackedTables:
    def __init__(self, tables):
        self.tables = tables
        self._keep = {}
        s = EqlbTables()
        for n in ["k", "p", "nrt", "ndg", "ndg_fct", "nq", "nqf", "ndiv", "nadd", "npk"]:
            setattr(s, n, getattr(tables, n))
        for n in _TABLE_DOUBLES + _TABLE_REF:
            a = np.ascontiguousarray(getattr(tables, n), dtype=np.float64)
            if a.size == 0:
                a = np.zeros(2)
            self._keep[n] = a
            setattr(s, n, _ptr(a, C.c_double))
        for n in _TABLE_INTS:
            a = np.ascontiguousarray(getattr(tables, n), dtype=np.int32)
            if a.size == 0:
                a = np.zeros(2, dtype=np.int32)
            self._keep[n] = a
            setattr(s, n, _ptr(a, C.c_int32))
        self.struct = s


def ptr_array(arrays, ctype=C.c_double):
    """Array of pointers (e.g. `const double* const*`) from a list of numpy arrays
    (entries may be None -> NULL)."""
    PT = C.POINTER(ctype)
    arr = (PT * len(arrays))()
    for i, a in enumerate(arrays):
        arr[i] = _ptr(a, ctype) if a is not None else PT()
    return arr


class EqlbFluxBC(C.Structure):
    _fields_ = [("nfct", C.c_int32), ("facets", c_int32_p), ("ncoef", C.c_int32), ("coeffs", c_double_p)]


_lib = None


def load_library():
    """Load the CUDA product library; raise loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"CUDA extension {LIB_PATH} not built - run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.eqlb_create.argtypes = [C.POINTER(EqlbMesh), C.POINTER(EqlbTables), C.c_int, C.c_uint32, C.POINTER(H)]
    lib.eqlb_create.restype = C.c_int
    lib.eqlb_destroy.argtypes = [H]
    lib.eqlb_destroy.restype = None
    lib.eqlb_set_stream.argtypes = [H, C.c_void_p]
    lib.eqlb_set_stream.restype = C.c_int
    lib.eqlb_set_bcs.argtypes = [H, c_int8_p, C.POINTER(c_double_p), c_int8_p, c_int8_p]
    lib.eqlb_set_bcs.restype = C.c_int
    lib.eqlb_se_run.argtypes = [H, C.POINTER(c_double_p), C.POINTER(c_double_p), C.POINTER(c_double_p), c_double_p, C.c_int]
    lib.eqlb_se_run.restype = C.c_int
    lib.eqlb_ev_run.argtypes = [H, C.POINTER(c_double_p), C.POINTER(c_double_p), C.POINTER(c_double_p), C.c_int]
    lib.eqlb_ev_run.restype = C.c_int
    lib.eqlb_local_project.argtypes = [H, C.c_int, C.POINTER(c_double_p), C.POINTER(c_double_p), C.c_int]
    lib.eqlb_local_project.restype = C.c_int
    lib.eqlb_patch_dims.argtypes = [H, c_int32_p, c_int32_p, c_int32_p]
    lib.eqlb_patch_dims.restype = C.c_int
    lib.eqlb_get_patch_maps.argtypes = [H, c_int32_p, c_int32_p, c_int32_p, c_int8_p, c_int8_p, c_int8_p, c_uint8_p, c_uint8_p, c_int32_p]
    lib.eqlb_get_patch_maps.restype = C.c_int
    lib.eqlb_get_se_dofmaps.argtypes = [H, c_int32_p, c_int32_p, c_int8_p, c_int32_p, c_int32_p]
    lib.eqlb_get_se_dofmaps.restype = C.c_int
    lib.eqlb_get_ev_dofmaps.argtypes = [H, c_int32_p, c_int32_p, c_int32_p, c_int8_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p,
                                        c_int32_p]
    lib.eqlb_get_ev_dofmaps.restype = C.c_int
    lib.eqlb_flux_l2norm.argtypes = [H, C.c_int, C.POINTER(c_double_p), C.POINTER(c_double_p), C.c_int]
    lib.eqlb_flux_l2norm.restype = C.c_int
    lib.eqlb_pin_host.argtypes = [C.c_void_p, C.c_size_t]
    lib.eqlb_pin_host.restype = C.c_int
    lib.eqlb_unpin_host.argtypes = [C.c_void_p]
    lib.eqlb_unpin_host.restype = C.c_int
    lib.eqlb_set_part.argtypes = [H, C.c_int]
    lib.eqlb_set_part.restype = C.c_int
    c_int64_p = C.POINTER(C.c_int64)
    c_ubyte_p = C.POINTER(C.c_ubyte)
    lib.eqlb_halo_create.argtypes = [C.c_int, c_int64_p, C.POINTER(c_int64_p), C.c_int, C.POINTER(H), c_ubyte_p, c_int64_p]
    lib.eqlb_halo_create.restype = C.c_int
    lib.eqlb_halo_connect.argtypes = [H, C.c_int, c_ubyte_p, C.c_int64, C.c_int]
    lib.eqlb_halo_connect.restype = C.c_int
    lib.eqlb_halo_apply.argtypes = [H, C.POINTER(c_double_p), C.c_int, C.c_void_p]
    lib.eqlb_halo_apply.restype = C.c_int
    lib.eqlb_halo_status.argtypes = [H, C.c_void_p]
    lib.eqlb_halo_status.restype = C.c_int
    lib.eqlb_get_launch_order.argtypes = [C.c_void_p, c_int32_p, C.POINTER(C.c_int32), c_int32_p]
    lib.eqlb_get_launch_order.restype = C.c_int
    lib.eqlb_get_staged_flux.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    lib.eqlb_get_staged_flux.restype = C.c_int
    lib.eqlb_halo_destroy.argtypes = [H]
    lib.eqlb_halo_destroy.restype = None
    lib.eqlb_launch_count.argtypes = [H]
    lib.eqlb_launch_count.restype = C.c_int64
    lib.eqlb_set_bcs_poly.argtypes = [H, c_int32_p, C.POINTER(c_int32_p), c_int32_p, C.POINTER(C.POINTER(EqlbFluxBC))]
    lib.eqlb_set_bcs_poly.restype = C.c_int
    lib.eqlb_get_boundary_data.argtypes = [H, c_int8_p, C.POINTER(c_double_p), c_int8_p, c_int8_p]
    lib.eqlb_get_boundary_data.restype = C.c_int
    lib.eqlb_ev_to_basix_rt.argtypes = [H, C.c_int, C.POINTER(c_double_p), C.POINTER(c_double_p), C.c_int]
    lib.eqlb_ev_to_basix_rt.restype = C.c_int
    PP = C.POINTER(c_double_p)
    lib.eqlb_set_primal_space.argtypes = [H, c_int32_p, C.c_int64]
    lib.eqlb_set_primal_space.restype = C.c_int
    lib.eqlb_project_primal.argtypes = [H, C.c_int, PP, PP, PP, PP, C.c_int]
    lib.eqlb_project_primal.restype = C.c_int
    lib.eqlb_ev_run_primal.argtypes = [H, PP, PP, PP, C.c_int]
    lib.eqlb_ev_run_primal.restype = C.c_int
    lib.eqlb_se_run_primal.argtypes = [H, PP, PP, PP, c_double_p, C.c_int]
    lib.eqlb_se_run_primal.restype = C.c_int
    lib.eqlb_estimate_poisson.argtypes = [H, C.c_int, PP, PP, PP, PP, PP, C.c_int, C.c_int]
    lib.eqlb_estimate_poisson.restype = C.c_int
    lib.eqlb_estimate_elasticity.argtypes = [H, PP, PP, PP, c_double_p, C.c_double, PP, C.c_int]
    lib.eqlb_estimate_elasticity.restype = C.c_int
    lib.eqlb_measure_fp64_peak.argtypes = [C.c_int, C.c_int, c_double_p]
    lib.eqlb_measure_fp64_peak.restype = C.c_int
    lib.eqlb_last_error.restype = C.c_char_p
    lib.eqlb_version.restype = C.c_char_p
    _lib = lib
    return lib
