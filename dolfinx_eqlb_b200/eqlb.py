"""Host-side mirror of the reference's equilibration API for the hot path.

Same names, argument meaning and error behaviour as
`python/dolfinx_eqlb/eqlb/FluxEqlbSE.py`, `FluxEqlbEV.py`, `bcs.py` and the
pybind entry points of `python/dolfinx_eqlb/wrappers.cpp:83-138`; DOLFINx
objects are replaced by the plain arrays the pybind layer would extract from
them (`Mesh` of `mesh.py`, DG coefficient vectors as numpy arrays or device
pointers).  All numerics run in the CUDA library behind the C ABI
(`include/eqlb_b200.h`); there is no CPU path in this module.
"""

from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import cabi
from .mesh import FACET_VERTS, Mesh
from .tables import Tables, make_tables

NORMAL_IS_OUTWARD = np.array([False, True, False])


def _check(lib, rc):
    if rc != 0:
        raise RuntimeError(lib.eqlb_last_error().decode())


class FluxBC:
    """Essential flux boundary condition on a set of facets.

    Stand-in for `dolfinx_eqlb.cpp.FluxBC` (`wrappers.cpp:144-232`) for the case the
    reference evaluates by interpolation (`base/BoundaryData.cpp:580-597`): the
    prescribed outward normal flux on facet i is the polynomial
    sum_j coeffs[i, j] s^j in the facet parameter s of the adjacent cell."""

    def __init__(self, facets, coeffs):
        self.facets = np.ascontiguousarray(facets, dtype=np.int32)
        self.coeffs = np.atleast_2d(np.asarray(coeffs, dtype=np.float64))
        if self.coeffs.shape[0] != self.facets.shape[0]:
            raise RuntimeError("FluxBC: one coefficient row per facet required")


def fluxbc(facets, coeffs) -> FluxBC:
    """`dolfinx_eqlb.eqlb.fluxbc` (`bcs.py:25-162`) for polynomial tractions."""
    return FluxBC(facets, coeffs)


class BoundaryData:
    """State of `base::BoundaryData` after its constructor
    (`base/BoundaryData.cpp:279-633`): facet types, boundary DOFs of the flux
    function (hierarchic facet moments), local facet ids, node markers."""

    def __init__(self, list_bcs, mesh: Mesh, tables: Tables, list_bfcts_prime, reconstruct_stress=False):
        nrhs = len(list_bcs)
        if len(list_bfcts_prime) != nrhs:
            raise RuntimeError("Mismatching inputs!")
        k, nrt = tables.k, tables.nrt
        self.num_rhs = nrhs
        self.facet_type = np.zeros((nrhs, mesh.nfct), dtype=np.int8)
        self.bflux = [None] * nrhs
        self.local_fct_id = np.zeros(mesh.nfct, dtype=np.int8)
        cnt = np.zeros(mesh.nnode, dtype=np.int32)
        x = mesh.x[:, :2]
        for r in range(nrhs):
            self.facet_type[r, np.asarray(list_bfcts_prime[r], dtype=np.int64)] = 1
            for bc in list_bcs[r]:
                if self.bflux[r] is None:
                    self.bflux[r] = np.zeros(mesh.ncell * nrt)
                f = bc.facets.astype(np.int64)
                c = mesh.fct_cell[mesh.fct_cell_off[f]].astype(np.int64)
                lf = np.argmax(mesh.cell_fct[c] == f[:, None], axis=1)
                va = mesh.cell_node[c, FACET_VERTS[lf, 0]]
                vb = mesh.cell_node[c, FACET_VERTS[lf, 1]]
                length = np.linalg.norm(x[va] - x[vb], axis=1)
                cn = mesh.cell_node[c]
                J00 = x[cn[:, 1], 0] - x[cn[:, 0], 0]
                J01 = x[cn[:, 2], 0] - x[cn[:, 0], 0]
                J10 = x[cn[:, 1], 1] - x[cn[:, 0], 1]
                J11 = x[cn[:, 2], 1] - x[cn[:, 0], 1]
                sgn = np.sign(J00 * J11 - J01 * J10)
                pre = np.where(NORMAL_IS_OUTWARD[lf], 1.0, -1.0) * sgn
                ng = bc.coeffs.shape[1]
                for j in range(k):
                    mom = sum(bc.coeffs[:, i] / (i + j + 1) for i in range(ng))
                    self.bflux[r][c * nrt + lf * k + j] = pre * length * mom
                self.facet_type[r, f] = 2
                self.local_fct_id[f] = lf
                if reconstruct_stress and r < 2:
                    np.add.at(cnt, mesh.fct_node[f].ravel(), 1)
        self.node_on_stress_bnd = (cnt == 4).astype(np.int8) if reconstruct_stress else None


def boundarydata(list_bcs, mesh, tables, list_bfcts_prime, equilibrate_stress=False) -> BoundaryData:
    """`dolfinx_eqlb.eqlb.boundarydata` (`bcs.py:165-215`)."""
    return BoundaryData(list_bcs, mesh, tables, list_bfcts_prime, equilibrate_stress)


class _Problem:
    """Owns the device-resident problem (C-ABI handle)."""

    def __init__(self, mesh: Mesh, tables: Tables, nrhs: int, stress=False, atomic=False, node_owned=None,
                 host_pipeline=False, generic=False, interface_first=False):
        self.lib = cabi.load_library()
        self.mesh, self.tables, self.nrhs = mesh, tables, nrhs
        self._pm = cabi.PackedMesh(mesh, tables.ndg, node_owned, identity_dg_as_null=True)
        self._pt = cabi.PackedTables(tables)
        flags = ((1 if stress else 0) | (2 if atomic else 0) | (4 if generic else 0) | (16 if host_pipeline else 0)
                 | (32 if interface_first else 0))
        self.stress = stress
        h = C.c_void_p()
        t0 = time.perf_counter()
        _check(self.lib, self.lib.eqlb_create(C.byref(self._pm.struct), C.byref(self._pt.struct), nrhs, flags, C.byref(h)))
        self.create_seconds = time.perf_counter() - t0  # mesh upload, Jacobians (C ABI call only)
        self.set_bcs_seconds = 0.0
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.eqlb_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_part(self, part: int):
        """0: all patches, 1: interface patches of a partitioned mesh, 2: interior patches (`eqlb_set_part`)."""
        _check(self.lib, self.lib.eqlb_set_part(self.h, int(part)))

    def set_stream(self, stream_ptr):
        _check(self.lib, self.lib.eqlb_set_stream(self.h, C.c_void_p(stream_ptr)))

    def set_bcs(self, bd: BoundaryData):
        nob = bd.node_on_stress_bnd
        t0 = time.perf_counter()
        _check(
            self.lib,
            self.lib.eqlb_set_bcs(
                self.h,
                bd.facet_type.ctypes.data_as(cabi.c_int8_p),
                cabi.ptr_array(bd.bflux),
                bd.local_fct_id.ctypes.data_as(cabi.c_int8_p),
                nob.ctypes.data_as(cabi.c_int8_p) if nob is not None else cabi.c_int8_p(),
            ),
        )
        self.set_bcs_seconds = time.perf_counter() - t0  # colouring, device patch builder (C ABI call only)

    def set_bcs_device(self, list_bcs, list_bfcts_prime):
        """Boundary data built ON THE DEVICE from polynomial tractions (`eqlb_set_bcs_poly`):
        facet types, cell-local facet ids, boundary DOFs and node markers by one thread per boundary facet -
        replaces the host loop of `BoundaryData` above / `base/BoundaryData.cpp:279-633`."""
        nrhs = self.nrhs
        if len(list_bcs) != nrhs or len(list_bfcts_prime) != nrhs:
            raise RuntimeError("Mismatching inputs!")
        prime = [np.ascontiguousarray(p, dtype=np.int32) for p in list_bfcts_prime]
        nprime = (C.c_int32 * nrhs)(*[p.shape[0] for p in prime])
        nbc = (C.c_int32 * nrhs)(*[len(b) for b in list_bcs])
        keep, rows = [], (C.POINTER(cabi.EqlbFluxBC) * nrhs)()
        for r, bcs in enumerate(list_bcs):
            arr = (cabi.EqlbFluxBC * max(len(bcs), 1))()
            for i, bc in enumerate(bcs):
                f = np.ascontiguousarray(bc.facets, dtype=np.int32)
                co = np.ascontiguousarray(bc.coeffs, dtype=np.float64)
                keep += [f, co]
                arr[i] = cabi.EqlbFluxBC(f.shape[0], f.ctypes.data_as(cabi.c_int32_p), co.shape[1], co.ctypes.data_as(cabi.c_double_p))
            keep.append(arr)
            rows[r] = C.cast(arr, C.POINTER(cabi.EqlbFluxBC))
        t0 = time.perf_counter()
        _check(self.lib, self.lib.eqlb_set_bcs_poly(self.h, nprime, cabi.ptr_array(prime, C.c_int32), nbc, rows))
        self.set_bcs_seconds = time.perf_counter() - t0

    def boundary_data(self):
        """Boundary data resident on the device (`eqlb_get_boundary_data`), as a `BoundaryData`-like object."""
        m, T = self.mesh, self.tables

        class _BD:
            pass

        bd = _BD()
        bd.facet_type = np.zeros((self.nrhs, m.nfct), np.int8)
        bd.bflux = [np.zeros(m.ncell * T.nrt) for _ in range(self.nrhs)]
        bd.local_fct_id = np.zeros(m.nfct, np.int8)
        bd.node_on_stress_bnd = np.zeros(m.nnode, np.int8) if self.stress else None
        nob = bd.node_on_stress_bnd
        _check(self.lib, self.lib.eqlb_get_boundary_data(
            self.h, bd.facet_type.ctypes.data_as(cabi.c_int8_p), cabi.ptr_array(bd.bflux),
            bd.local_fct_id.ctypes.data_as(cabi.c_int8_p), nob.ctypes.data_as(cabi.c_int8_p) if nob is not None else cabi.c_int8_p()))
        bd.num_rhs = self.nrhs
        return bd

    # ---- fused input stage and error estimator ----
    def set_primal_space(self, pk_dofmap, ndofs):
        """Cell dofmap [ncell][npk] of the continuous P_k space of the primal solution (Basix DOF order)."""
        dm = np.ascontiguousarray(pk_dofmap, dtype=np.int32)
        if dm.shape != (self.mesh.ncell, self.tables.npk):
            raise RuntimeError("set_primal_space: dofmap must be [ncell][npk]")
        self.pk_ndofs = int(ndofs)
        _check(self.lib, self.lib.eqlb_set_primal_space(self.h, dm.ctypes.data_as(cabi.c_int32_p), int(ndofs)))

    def project_primal(self, uh=None, fh=None):
        """`lsolver/projection.py:17-77` on the device: (G, F) = (Pi(-grad u_h), Pi f_h) from P_k coefficient vectors."""
        n = len(uh) if uh is not None else len(fh)
        T, m = self.tables, self.mesh
        G = [np.zeros(m.ncell * T.ndg * 2) for _ in range(n)] if uh is not None else None
        F = [np.zeros(m.ncell * T.ndg) for _ in range(n)] if fh is not None else None
        pa = lambda xs: cabi.ptr_array(_as_ptr_list(xs)) if xs is not None else C.POINTER(cabi.c_double_p)()
        keep = (_as_ptr_list(uh) if uh is not None else None, _as_ptr_list(fh) if fh is not None else None)
        _check(self.lib, self.lib.eqlb_project_primal(
            self.h, n, cabi.ptr_array(keep[0]) if keep[0] is not None else C.POINTER(cabi.c_double_p)(),
            cabi.ptr_array(keep[1]) if keep[1] is not None else C.POINTER(cabi.c_double_p)(),
            cabi.ptr_array(G) if G is not None else C.POINTER(cabi.c_double_p)(),
            cabi.ptr_array(F) if F is not None else C.POINTER(cabi.c_double_p)(), 0))
        return G, F

    def run_primal(self, ev, uh, fh, sigma, korn=None, zeroed=False):
        """projection + equilibration in one call (`eqlb_ev_run_primal` / `eqlb_se_run_primal`), host vectors"""
        u, f = _as_ptr_list(uh), _as_ptr_list(fh)
        ms = 2 if zeroed else 0
        if ev:
            _check(self.lib, self.lib.eqlb_ev_run_primal(self.h, cabi.ptr_array(u), cabi.ptr_array(f), cabi.ptr_array(sigma), ms))
        else:
            kp = korn.ctypes.data_as(cabi.c_double_p) if korn is not None else cabi.c_double_p()
            _check(self.lib, self.lib.eqlb_se_run_primal(self.h, cabi.ptr_array(u), cabi.ptr_array(f), cabi.ptr_array(sigma), kp, ms))

    def estimate_poisson(self, sigma, uh, fh, is_ev):
        """cell-wise (eta_sig^2, eta_osc^2) of `demo/poisson/demo_error_estimation.py:52-122` on the device"""
        n = len(sigma)
        s, u, f = _as_ptr_list(sigma), _as_ptr_list(uh), _as_ptr_list(fh)
        e1 = [np.zeros(self.mesh.ncell) for _ in range(n)]
        e2 = [np.zeros(self.mesh.ncell) for _ in range(n)]
        _check(self.lib, self.lib.eqlb_estimate_poisson(self.h, n, cabi.ptr_array(s), cabi.ptr_array(u), cabi.ptr_array(f),
                                                        cabi.ptr_array(e1), cabi.ptr_array(e2), int(bool(is_ev)), 0))
        return e1, e2

    def estimate_elasticity(self, dsig, sigma_h, fh, korn, pi_1):
        """cell-wise indicators of `demo/elasticity/demo_error_estimation.py:49-135` (displacement formulation)"""
        s, sh, f = _as_ptr_list(dsig), _as_ptr_list(sigma_h), _as_ptr_list(fh)
        kc = np.ascontiguousarray(korn, dtype=np.float64)
        eta = [np.zeros(self.mesh.ncell) for _ in range(3)]
        _check(self.lib, self.lib.eqlb_estimate_elasticity(self.h, cabi.ptr_array(s), cabi.ptr_array(sh), cabi.ptr_array(f),
                                                           kc.ctypes.data_as(cabi.c_double_p), float(pi_1), cabi.ptr_array(eta), 0))
        return eta

    def launch_count(self):
        return int(self.lib.eqlb_launch_count(self.h))

    # ---- integer maps (parity evidence) ----
    def patch_dims(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        _check(self.lib, self.lib.eqlb_patch_dims(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def patch_maps(self):
        npatch, ncmax, ncol = self.patch_dims()
        out = dict(
            ncells=np.zeros(npatch, np.int32),
            cells=np.zeros((npatch, ncmax + 2), np.int32),
            fcts=np.zeros((npatch, ncmax + 2), np.int32),
            inodes_local=np.zeros((npatch, ncmax + 2), np.int8),
            fcts_local=np.zeros((npatch, 2 * (ncmax + 1)), np.int8),
            type=np.zeros((npatch, self.nrhs), np.int8),
            reversed=np.zeros((npatch, ncmax, 2), np.uint8),
            reversion=np.zeros((npatch, self.nrhs), np.uint8),
            colour=np.zeros(npatch, np.int32),
        )
        p = lambda a, t: a.ctypes.data_as(t)
        _check(
            self.lib,
            self.lib.eqlb_get_patch_maps(
                self.h, p(out["ncells"], cabi.c_int32_p), p(out["cells"], cabi.c_int32_p), p(out["fcts"], cabi.c_int32_p),
                p(out["inodes_local"], cabi.c_int8_p), p(out["fcts_local"], cabi.c_int8_p), p(out["type"], cabi.c_int8_p),
                p(out["reversed"], cabi.c_uint8_p), p(out["reversion"], cabi.c_uint8_p), p(out["colour"], cabi.c_int32_p),
            ),
        )
        out["ncmax"], out["ncolours"] = ncmax, ncol
        return out

    def se_dofmaps(self):
        npatch, ncmax, _ = self.patch_dims()
        T = self.tables
        ndpc = 2 * T.k + T.nadd + T.ndiv + (3 if self.stress else 0)
        hzmax = 1 + (T.k - 1) * (ncmax + 1) + T.nadd * ncmax
        dm = np.zeros((npatch, 4, ncmax + 2, ndpc), np.int32)
        pf = np.zeros((npatch, ncmax + 1, 2 * T.ndg_fct), np.int32)
        bm = np.zeros((npatch, self.nrhs, hzmax), np.int8)
        a, b = C.c_int32(), C.c_int32()
        _check(
            self.lib,
            self.lib.eqlb_get_se_dofmaps(
                self.h, dm.ctypes.data_as(cabi.c_int32_p), pf.ctypes.data_as(cabi.c_int32_p),
                bm.ctypes.data_as(cabi.c_int8_p), C.byref(a), C.byref(b),
            ),
        )
        assert a.value == ndpc and b.value == hzmax
        return dict(dofmap=dm, projflux_fct=pf, bmarkers=bm)


    def ev_dofmaps(self):
        """EV patch ordering + sub-DOF maps (`ev/Patch.cpp:482-676`) of all patches, built on the device."""
        npatch, ncmax, _ = self.patch_dims()
        T = self.tables
        nz = T.nrt + T.ndg - T.k
        lenf = ncmax * (T.k * T.k - T.k) + (ncmax + 1) * T.k
        out = dict(
            ncells=np.zeros(npatch, np.int32), cells=np.zeros((npatch, ncmax), np.int32),
            fcts=np.zeros((npatch, ncmax + 1), np.int32), inodes_local=np.zeros((npatch, ncmax), np.int8),
            dofs_elmt=np.zeros((npatch, ncmax * nz), np.int32), dofs_patch=np.zeros((npatch, ncmax * nz), np.int32),
            dofs_global=np.zeros((npatch, ncmax * nz), np.int32), list_patch=np.zeros((npatch, lenf), np.int32),
            list_global=np.zeros((npatch, lenf), np.int32),
        )
        p32 = lambda a: a.ctypes.data_as(cabi.c_int32_p)
        _check(
            self.lib,
            self.lib.eqlb_get_ev_dofmaps(
                self.h, p32(out["ncells"]), p32(out["cells"]), p32(out["fcts"]), out["inodes_local"].ctypes.data_as(cabi.c_int8_p),
                p32(out["dofs_elmt"]), p32(out["dofs_patch"]), p32(out["dofs_global"]), p32(out["list_patch"]),
                p32(out["list_global"]),
            ),
        )
        return out


def _as_ptr_list(arrs):
    return [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]


def _check_inplace(vectors, size, what):
    """Result vectors are accumulated in place by the C ABI (which carries no sizes): contiguous float64 arrays of
    exactly the expected length, else the reference's 'Input sizes does not match'."""
    for s in vectors:
        if not (isinstance(s, np.ndarray) and s.dtype == np.float64 and s.flags.c_contiguous):
            raise RuntimeError(f"{what} must be contiguous float64 arrays (accumulated in place)")
        if s.shape != (size,):
            raise RuntimeError("Equilibration: Input sizes does not match")


def _check_inputs(problem, G, F):
    m, T = problem.mesh, problem.tables
    for g in G:
        if g.shape != (m.ncell * T.ndg * 2,):
            raise RuntimeError("Equilibration: Input sizes does not match")
    for f in F:
        if f.shape != (m.ncell * T.ndg,):
            raise RuntimeError("Equilibration: Input sizes does not match")


def reconstruct_fluxes_semiexplt(problem: _Problem, flux_hdiv, flux_dg, rhs_dg, korn=None, zeroed=False):
    """`cpp.reconstruct_fluxes_semiexplt[_with_kornconst]` (`wrappers.cpp:97-137`):
    host numpy vectors, accumulated in place into `flux_hdiv` (`zeroed`: the caller
    guarantees `flux_hdiv` is zero on entry, its upload is skipped)."""
    lib = problem.lib
    G, F = _as_ptr_list(flux_dg), _as_ptr_list(rhs_dg)
    if not (len(G) == len(F) == len(flux_hdiv) == problem.nrhs):
        raise RuntimeError("Equilibration: Input sizes does not match")
    _check_inplace(flux_hdiv, problem.mesh.ncell * problem.tables.nrt, "flux_hdiv")
    _check_inputs(problem, G, F)
    if korn is not None:
        _check_inplace([korn], problem.mesh.ncell, "cells_kornconst")
    kp = korn.ctypes.data_as(cabi.c_double_p) if korn is not None else cabi.c_double_p()
    _check(lib, lib.eqlb_se_run(problem.h, cabi.ptr_array(G), cabi.ptr_array(F), cabi.ptr_array(flux_hdiv), kp, 2 if zeroed else 0))


def reconstruct_fluxes_minimisation(problem: _Problem, flux_hdiv, flux_dg, rhs_dg, zeroed=False):
    """`cpp.reconstruct_fluxes_minimisation` (`wrappers.cpp:85-95`) for the fixed
    forms of `FluxEqlbEV.py:116-133`."""
    lib = problem.lib
    G, F = _as_ptr_list(flux_dg), _as_ptr_list(rhs_dg)
    if not (len(G) == len(F) == len(flux_hdiv) == problem.nrhs):
        raise RuntimeError("Equilibration: Input sizes does not match")
    k = problem.tables.k
    _check_inplace(flux_hdiv, problem.mesh.nfct * k + problem.mesh.ncell * (k * k - k), "flux_hdiv")
    _check_inputs(problem, G, F)
    _check(lib, lib.eqlb_ev_run(problem.h, cabi.ptr_array(G), cabi.ptr_array(F), cabi.ptr_array(flux_hdiv), 2 if zeroed else 0))


class FluxEquilibrator:
    """`eqlb/FluxEquilibrator.py:15-96`."""

    def __init__(self, degree_flux: int, n_eqlbs: int, equilibrate_stress: bool):
        self.degree_flux = degree_flux
        self.n_fluxes = n_eqlbs
        self.equilibrate_stresses = equilibrate_stress
        self.list_flux = []
        self.list_bfunctions = []
        self.boundary_data = None
        self._pinned = []
        self._pin_pending = None  # host vectors to page-lock before the second call (see _note_call)

    def _pin(self, arrays, min_bytes=4 << 20):
        """Page-lock large host vectors once (`eqlb_pin_host`) so that the host-pointer calls copy
        at PCIe speed; small vectors are not worth the registration cost."""
        lib = cabi.load_library()
        todo = [a for a in arrays
                if isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.nbytes >= min_bytes]
        if not todo:
            return
        # the registrations of different vectors run side by side (ctypes releases the GIL)
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=min(4, len(todo))) as ex:
            rcs = list(ex.map(lambda a: lib.eqlb_pin_host(a.ctypes.data, a.nbytes), todo))
        self._pinned.extend(a for a, rc in zip(todo, rcs) if rc == 0)

    def _note_call(self):
        """The first call on a new problem runs from pageable memory (the library stages the copies through its
        pinned pool); registering the caller's vectors costs about one such call and is done when a second call
        shows that the problem is reused (time loops) - then the stage pipeline takes over."""
        if self._pin_pending is not None:
            if self._pin_pending[0] >= 1:
                self._pin(self._pin_pending[1])
                self._pin_pending = None
            else:
                self._pin_pending[0] += 1

    def __del__(self):
        try:
            lib = cabi.load_library()
            for a in self._pinned:
                lib.eqlb_unpin_host(a.ctypes.data)
            self._pinned = []
        except Exception:
            pass


class FluxEqlbSE(FluxEquilibrator):
    """`eqlb/FluxEqlbSE.py:24-198` on top of the CUDA hot path."""

    def __init__(self, degree_flux, msh: Mesh, list_rhs, list_proj_flux, equilibrate_stress=False,
                 estimate_korn_constant=False, degree_proj=None, atomic=False, node_owned=None, host_pipeline=True,
                 generic=False, interface_first=False):
        super().__init__(degree_flux, len(list_rhs), equilibrate_stress)
        if len(list_proj_flux) != self.n_fluxes:
            raise RuntimeError("Mismatching inputs!")
        self.mesh = msh
        self.tables = make_tables(degree_flux, degree_proj)
        self.list_rhs, self.list_proj_flux = list_rhs, list_proj_flux
        self.estimate_korn_constant = estimate_korn_constant
        self.korn_constants = np.zeros(msh.ncell) if estimate_korn_constant else None
        self.problem = _Problem(msh, self.tables, self.n_fluxes, equilibrate_stress, atomic, node_owned, host_pipeline, generic,
                                interface_first)
        self.list_flux = [np.zeros(msh.ncell * self.tables.nrt) for _ in range(self.n_fluxes)]
        self._fresh = True  # list_flux still holds the zeros it was created with
        if host_pipeline:
            self._pin_pending = [0, self.list_flux + list(list_rhs) + list(list_proj_flux)]

    def set_boundary_conditions(self, list_bfct_prime, list_bcs_flux, device=False):
        """`device=True`: the boundary data are built on the GPU (`eqlb_set_bcs_poly`); else by the host mirror
        `BoundaryData` and uploaded (`eqlb_set_bcs`).  Same result (tests/test_gpu_bcs.py)."""
        if self.n_fluxes != len(list_bfct_prime) or self.n_fluxes != len(list_bcs_flux):
            raise RuntimeError("Mismatching inputs!")
        if device:
            self.problem.set_bcs_device(list_bcs_flux, list_bfct_prime)
            self.boundary_data = self.problem.boundary_data()
            self.list_bfunctions = self.boundary_data.bflux
            return
        self.boundary_data = boundarydata(list_bcs_flux, self.mesh, self.tables, list_bfct_prime, self.equilibrate_stresses)
        self.list_bfunctions = self.boundary_data.bflux
        self.problem.set_bcs(self.boundary_data)

    def equilibrate_fluxes(self):
        self._note_call()
        reconstruct_fluxes_semiexplt(self.problem, self.list_flux, self.list_proj_flux, self.list_rhs, self.korn_constants,
                                     zeroed=self._fresh)
        self._fresh = False
        if self.estimate_korn_constant:
            self.korn_constants[:] = np.sqrt(self.korn_constants)

    def flux_l2norm(self):
        """Cell-wise ||sigma_i||^2_{L2(T)} of the equilibrated (DRT) fluxes: the flux part of the error
        indicators of `demo_error_estimation.py:96-101`, evaluated on the GPU (`eqlb_flux_l2norm`)."""
        out = [np.zeros(self.mesh.ncell) for _ in range(self.n_fluxes)]
        _check(self.problem.lib, self.problem.lib.eqlb_flux_l2norm(self.problem.h, self.n_fluxes, cabi.ptr_array(self.list_flux),
                                                                   cabi.ptr_array(out), 0))
        return out

    def get_korn_constants(self):
        if self.estimate_korn_constant:
            return self.korn_constants
        raise RuntimeError("Korn constants are not estimated!")


class FluxEqlbEV(FluxEquilibrator):
    """`eqlb/FluxEqlbEV.py:20-188` on top of the CUDA hot path; the flux lives in
    the conforming hierarchic RT_k space ([facet dofs nfct*k][cell dofs])."""

    def __init__(self, degree_flux, msh: Mesh, list_rhs, list_proj_flux, node_owned=None, host_pipeline=True, generic=False,
                 interface_first=False):
        super().__init__(degree_flux, len(list_rhs), False)
        if len(list_proj_flux) != self.n_fluxes:
            raise RuntimeError("Missmatching inputs!")
        self.mesh = msh
        self.tables = make_tables(degree_flux)
        self.list_rhs, self.list_proj_flux = list_rhs, list_proj_flux
        self.problem = _Problem(msh, self.tables, self.n_fluxes, False, False, node_owned, host_pipeline, generic, interface_first)
        k = degree_flux
        self.ndofs = msh.nfct * k + msh.ncell * (k * k - k)
        self.list_flux = [np.zeros(self.ndofs) for _ in range(self.n_fluxes)]
        self._fresh = True
        if host_pipeline:
            self._pin_pending = [0, self.list_flux + list(list_rhs) + list(list_proj_flux)]

    def set_boundary_conditions(self, list_bfct_prime, list_bcs_flux, device=False):
        if self.n_fluxes != len(list_bfct_prime) or self.n_fluxes != len(list_bcs_flux):
            raise RuntimeError("Mismatching inputs!")
        if device:
            self.problem.set_bcs_device(list_bcs_flux, list_bfct_prime)
            self.boundary_data = self.problem.boundary_data()
            self.list_bfunctions = self.boundary_data.bflux
            return
        self.boundary_data = boundarydata(list_bcs_flux, self.mesh, self.tables, list_bfct_prime, False)
        self.list_bfunctions = self.boundary_data.bflux
        self.problem.set_bcs(self.boundary_data)

    def equilibrate_fluxes(self):
        self._note_call()
        reconstruct_fluxes_minimisation(self.problem, self.list_flux, self.list_proj_flux, self.list_rhs, zeroed=self._fresh)
        self._fresh = False

    def fluxes_in_basix_rt(self):
        """The equilibrated fluxes as DOF vectors of Basix' "RT" element (Legendre variant) - the space the
        reference's FluxEqlbEV returns (`FluxEqlbEV.py:95`); layout [facet dofs nfct*k][cell dofs] (`eqlb_ev_to_basix_rt`)."""
        out = [np.zeros(self.ndofs) for _ in range(self.n_fluxes)]
        _check(self.problem.lib, self.problem.lib.eqlb_ev_to_basix_rt(self.problem.h, self.n_fluxes, cabi.ptr_array(self.list_flux),
                                                                      cabi.ptr_array(out), 0))
        return out


def local_projection(problem: _Problem, qvals):
    """`dolfinx_eqlb.lsolver.local_projection` (`lsolver/projection.py:17-77`) for DG_p
    targets on affine cells: `qvals[i][cell*nq + q]` are the values of the i-th function at
    the cell quadrature points; returns the DG_p coefficient vectors (CUDA)."""
    lib = problem.lib
    qv = _as_ptr_list(qvals)
    T, m = problem.tables, problem.mesh
    for q in qv:
        if q.shape[0] != m.ncell * T.nq:
            raise RuntimeError("local_projection: one value per cell quadrature point required")
    out = [np.zeros(m.ncell * T.ndg) for _ in qv]
    _check(lib, lib.eqlb_local_project(problem.h, len(qv), cabi.ptr_array(qv), cabi.ptr_array(out), 0))
    return out
