"""GPU parity tests of the SE path (call through the C ABI)."""

import numpy as np
import pytest

import fem_mini as fm
from common import PoissonCase, make_mesh
from dolfinx_eqlb_b200 import eqlb

pytestmark = pytest.mark.gpu

RTOL = 1e-10  # BASELINE.json north_star: flux DOFs within 1e-10 relative in FP64

INT_KEYS = ["ncells", "cells", "fcts", "inodes_local", "fcts_local", "type", "reversed", "reversion"]


def run_gpu(case, atomic=False):
    eq = eqlb.FluxEqlbSE(case.k, case.mesh, case.F, case.G, degree_proj=case.T.p, atomic=atomic)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    eq.equilibrate_fluxes()
    return eq


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("kind,n,scramble", [("crossed", 4, None), ("crossed", 5, 3), ("randdiag", 6, 2)])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_patch_maps_bit_exact(kind, n, scramble, k):
    from oracle import pyoracle as po

    m = make_mesh(kind, n, scramble, perturb=0.2)
    case = PoissonCase(m, k, [[1, 4], [1, 3], [2], [1, 3, 4]], seed=5, galerkin=False)
    ref = po.se_patch_maps(m, case.T, case.oracle_bc())
    eq = eqlb.FluxEqlbSE(k, m, case.F, case.G)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    got = eq.problem.patch_maps()
    for key in INT_KEYS:
        assert np.array_equal(got[key], ref[key]), key
    dm = eq.problem.se_dofmaps()
    for key in ["dofmap", "projflux_fct", "bmarkers"]:
        assert np.array_equal(dm[key], ref[key]), key
    # colouring: no two patches of a colour share a cell
    col = got["colour"]
    assert (col[m.cell_node[:, 0]] != col[m.cell_node[:, 1]]).all()
    assert (col[m.cell_node[:, 0]] != col[m.cell_node[:, 2]]).all()
    assert (col[m.cell_node[:, 1]] != col[m.cell_node[:, 2]]).all()


@pytest.mark.parametrize("kind,n,scramble", [("crossed", 4, None), ("crossed", 5, 3), ("randdiag", 6, 2)])
@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("nsets", [[[]], [[1, 4]], [[1, 4], [1, 3], [2], [1, 3, 4]]])
def test_se_flux_parity(kind, n, scramble, k, nsets):
    from oracle import pyoracle as po

    m = make_mesh(kind, n, scramble, perturb=0.25)
    case = PoissonCase(m, k, nsets, seed=7)
    ref = po.se_run(m, case.T, case.oracle_bc(), case.G, case.F)
    eq = run_gpu(case)
    for r in range(case.nrhs):
        assert rel_err(eq.list_flux[r], ref[r]) < RTOL
        # the reference's acceptance invariants hold for the GPU result itself
        assert fm.check_divergence(m, case.T, eq.list_flux[r], case.G[r], case.F[r]) < 1e-12
        assert fm.check_jump(m, case.T, eq.list_flux[r], case.G[r]) < 1e-11
        if len(case.neu[r]):
            assert fm.check_bc(m, case.T, eq.list_flux[r], case.G[r], case.bdata.bflux[r], case.neu[r]) < 1e-11


def test_se_random_data_parity():
    """Non-Galerkin random data (the benchmark's input distribution): parity with the
    oracle must hold for any input, closure of the patch fluxes is not required."""
    from oracle import pyoracle as po

    m = make_mesh("crossed", 8, None)
    case = PoissonCase(m, 2, [[]], seed=11, galerkin=False)
    ref = po.se_run(m, case.T, case.oracle_bc(), case.G, case.F)
    eq = run_gpu(case)
    assert rel_err(eq.list_flux[0], ref[0]) < RTOL


def test_se_atomic_matches_coloured():
    m = make_mesh("crossed", 6, 4, perturb=0.2)
    case = PoissonCase(m, 2, [[1, 4]], seed=3)
    a = run_gpu(case, atomic=False).list_flux[0]
    b = run_gpu(case, atomic=True).list_flux[0]
    assert rel_err(a, b) < 1e-13


def test_se_accumulates_like_reference():
    """Output is += (se/solve_patch_semiexplt.hpp:1159): a second call doubles it."""
    m = make_mesh("crossed", 4, None)
    case = PoissonCase(m, 2, [[]], seed=1)
    eq = run_gpu(case)
    once = eq.list_flux[0].copy()
    eq.equilibrate_fluxes()
    assert rel_err(eq.list_flux[0], 2 * once) < 1e-14


def test_lower_degree_data():
    """degree of projected data below k-1 (reference test matrix
    test_fluxeqlb_conditions.py:62-64)."""
    from oracle import pyoracle as po

    m = make_mesh("crossed", 4, 2, perturb=0.2)
    for k, p in [(2, 0), (3, 1), (3, 0)]:
        case = PoissonCase(m, k, [[1, 4]], seed=2, p=p)
        ref = po.se_run(m, case.T, case.oracle_bc(), case.G, case.F)
        eq = run_gpu(case)
        assert rel_err(eq.list_flux[0], ref[0]) < RTOL


def test_errors_like_reference():
    from dolfinx_eqlb_b200 import mesh as ms

    # "right" diagonal mesh has 1-cell corner patches -> reference throws (se/Patch.cpp:353-359)
    x = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    m = ms.build_topology(x, np.array([[0, 1, 3], [0, 2, 3]]))
    with pytest.raises(RuntimeError, match="has only 1 cells"):
        eqlb.FluxEqlbSE(1, m, [np.zeros(2)], [np.zeros(4)])


@pytest.mark.parametrize("env", [{"EQLB_KW": "1"}, {"EQLB_RED": "0"}, {"EQLB_STRESS_GENERIC": "1"}])
def test_alternate_kernel_switches(env):
    """The environment switches are read once per process, so the alternate paths run in a
    child process: degree 2 on the general-degree kernel, load-add-store accumulation,
    generic weak-symmetry stage - all against the oracle."""
    import os
    import subprocess
    import sys

    code = r"""
import sys, numpy as np
sys.path.insert(0, 'tests')
from common import PoissonCase, make_mesh
from test_gpu_stress import elasticity_case
from dolfinx_eqlb_b200 import eqlb
from oracle import pyoracle as po
m = make_mesh('crossed', 6, 3, perturb=0.2)
case = PoissonCase(m, 2, [[1, 4]], seed=7)
for cls, ref in ((eqlb.FluxEqlbSE, po.se_run), (eqlb.FluxEqlbEV, po.ev_run)):
    eq = cls(2, m, case.F, case.G)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    eq.equilibrate_fluxes()
    r = ref(m, case.T, case.oracle_bc(), case.G, case.F)[0]
    assert np.abs(eq.list_flux[0] - r).max() < 1e-10 * np.abs(r).max()
T, G, f, bfp, bcs, neu = elasticity_case(m, 2, [], seed=3)
eq = eqlb.FluxEqlbSE(2, m, f, G, equilibrate_stress=True)
eq.set_boundary_conditions(bfp, bcs)
bd = eq.boundary_data
r = po.se_run(m, T, po.BCData(bd.facet_type, bd.bflux, bd.local_fct_id, bd.node_on_stress_bnd), G, f, stress=True)
eq.equilibrate_fluxes()
for i in range(2):
    assert np.abs(eq.list_flux[i] - r[i]).max() < 1e-10 * np.abs(r[i]).max()
print('ok')
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def fan_mesh(nfan, seed=0):
    """Disc triangulated as two rings: a central vertex of valence `nfan` (one interior patch
    with nfan cells -> the 16-lane variants of the lane-per-cell kernels for nfan > 8)."""
    from dolfinx_eqlb_b200 import mesh as ms

    rng = np.random.default_rng(seed)
    ang = 2 * np.pi * (np.arange(nfan) + 0.15 * rng.random(nfan)) / nfan
    inner = 0.5 * np.stack([np.cos(ang), np.sin(ang)], 1)
    outer = 1.0 * np.stack([np.cos(ang + np.pi / nfan), np.sin(ang + np.pi / nfan)], 1)
    x = np.concatenate([[[0.0, 0.0]], inner, outer])
    tris = []
    for i in range(nfan):
        a, b = 1 + i, 1 + (i + 1) % nfan
        o = 1 + nfan + i
        tris += [[0, a, b], [a, o, b], [b, o, 1 + nfan + (i + 1) % nfan]]
    perm = rng.permutation(x.shape[0])  # scrambled vertex numbering: reversed facets, negative Jacobians
    inv = np.argsort(perm)
    return ms.build_topology(x[perm], inv[np.array(tris)].astype(np.int32))


@pytest.mark.parametrize("nfan", [9, 13, 16, 20, 32])  # > 16: generic kernel for the hub, lane kernels for the rest
@pytest.mark.parametrize("k", [1, 2, 3])
def test_high_valence_patches(nfan, k):
    """Patches with 9..16 cells take the S = 16 lane variants; parity against the oracle for SE, EV
    and (k = 2) the fused weak-symmetry stage."""
    from oracle import pyoracle as po
    from dolfinx_eqlb_b200 import tables as tb

    m = fan_mesh(nfan, seed=nfan)
    T = tb.make_tables(k)
    rng = np.random.default_rng(k)
    nrhs = 2
    G = [rng.standard_normal(m.ncell * T.ndg * 2) for _ in range(nrhs)]
    F = [rng.standard_normal(m.ncell * T.ndg) for _ in range(nrhs)]
    bf = [m.bfct.astype(np.int32)] * nrhs
    ft = np.zeros((nrhs, m.nfct), np.int8)
    ft[:, m.bfct] = 1
    bc = po.BCData(ft)
    for cls, ref in ((eqlb.FluxEqlbSE, po.se_run), (eqlb.FluxEqlbEV, po.ev_run)):
        eq = cls(k, m, F, G)
        eq.set_boundary_conditions(bf, [[], []])
        eq.equilibrate_fluxes()
        r = ref(m, T, bc, G, F)
        for i in range(nrhs):
            assert np.abs(eq.list_flux[i] - r[i]).max() < 1e-10 * np.abs(r[i]).max()
    if k == 2:
        eq = eqlb.FluxEqlbSE(k, m, F, G, equilibrate_stress=True)
        eq.set_boundary_conditions(bf, [[], []])
        eq.equilibrate_fluxes()
        r = po.se_run(m, T, po.BCData(ft, None, None, np.zeros(m.nnode, np.int8)), G, F, stress=True)
        for i in range(nrhs):
            assert np.abs(eq.list_flux[i] - r[i]).max() < 1e-10 * np.abs(r[i]).max()


def test_too_many_cells_is_an_error():
    """More than 32 cells around a vertex: rejected loudly by eqlb_create (no silent fallback); 17..32 cells
    are served by the generic kernel (tests/test_large_patches.py)."""
    from dolfinx_eqlb_b200 import tables as tb

    m = fan_mesh(33, seed=1)
    T = tb.make_tables(2)
    z = [np.zeros(m.ncell * T.ndg)]
    with pytest.raises(RuntimeError, match="more than 32 cells"):
        eqlb.FluxEqlbSE(2, m, z, [np.zeros(m.ncell * T.ndg * 2)])


@pytest.mark.parametrize("kind,n,scramble", [("crossed", 24, None), ("crossed", 17, 3), ("randdiag", 40, 2), ("fan", 8, 1)])
def test_device_colouring_is_the_sequential_first_fit(kind, n, scramble):
    """The colouring computed on the device (wavefront of dependent decisions, `greedy_colour_kernel`) is exactly
    the sequential first-fit colouring in vertex order - evaluated here by a plain Python loop."""
    from common import make_mesh
    from dolfinx_eqlb_b200 import tables as tb

    m = make_mesh(kind, n, scramble)
    T = tb.make_tables(1)
    z = [np.zeros(m.ncell * T.ndg)]
    eq = eqlb.FluxEqlbSE(1, m, z, [np.zeros(m.ncell * T.ndg * 2)])
    eq.set_boundary_conditions([m.bfct.astype(np.int32)], [[]])
    got = eq.problem.patch_maps()
    colour = -np.ones(m.nnode, dtype=np.int64)
    for v in range(m.nnode):
        cells = m.node_cell[m.node_cell_off[v]:m.node_cell_off[v + 1]]
        used = set(colour[m.cell_node[cells].ravel()].tolist())
        c = 0
        while c in used:
            c += 1
        colour[v] = c
    assert np.array_equal(got["colour"], colour)
    assert got["ncolours"] == colour.max() + 1


@pytest.mark.parametrize("k", [1, 2, 3])
def test_permuted_dg_dofmap(k):
    """`eqlb_mesh.dg_dofmap` other than the DOLFINx layout: G and f are read through the map (generic kernel);
    same fluxes as the identity layout with the vectors permuted back"""
    from common import PoissonCase, make_mesh

    m = make_mesh("crossed", 6, 3, perturb=0.2)
    case = PoissonCase(m, k, [[1, 4]], seed=2, galerkin=False)
    ndg = case.T.ndg
    ref = {}
    for cls in (eqlb.FluxEqlbSE, eqlb.FluxEqlbEV):
        eq = cls(k, m, case.F, case.G)
        eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
        eq.equilibrate_fluxes()
        ref[cls] = np.array(eq.list_flux[0])
    perm = np.random.default_rng(0).permutation(m.ncell * ndg).astype(np.int32)  # dof of (cell, i) = perm[cell*ndg + i]
    Fp, Gp = np.zeros_like(case.F[0]), np.zeros_like(case.G[0])
    Fp[perm] = case.F[0]
    Gp.reshape(-1, 2)[perm] = case.G[0].reshape(-1, 2)
    m.dg_dofmap = perm.reshape(m.ncell, ndg)
    try:
        for cls in (eqlb.FluxEqlbSE, eqlb.FluxEqlbEV):
            eq = cls(k, m, [Fp], [Gp])
            eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
            eq.equilibrate_fluxes()
            assert np.abs(eq.list_flux[0] - ref[cls]).max() < 1e-11 * np.abs(ref[cls]).max()
    finally:
        del m.dg_dofmap


@pytest.mark.parametrize("kind,n,scramble,pipeline", [("randdiag", 40, 2, False), ("crossed", 40, 3, True), ("fan", 8, 1, False)])
def test_device_launch_order(kind, n, scramble, pipeline):
    """The launch order computed on the device (raw keys + histogram + stable radix sort) is the order the rule
    defines: patches sorted by (segment = chunk * ncolours + colour, lane class with small classes merged into the
    next wider one, node id); chunk of a patch = chunk of its last cell (stages of the host pipeline)."""
    import ctypes as C

    from common import make_mesh
    from dolfinx_eqlb_b200 import cabi, tables as tb

    m = make_mesh(kind, n, scramble)
    T = tb.make_tables(2)
    eq = eqlb.FluxEqlbEV(2, m, [np.zeros(m.ncell * T.ndg)], [np.zeros(m.ncell * T.ndg * 2)], host_pipeline=pipeline)
    eq.set_boundary_conditions([m.bfct.astype(np.int32)], [[]])
    got = eq.problem.patch_maps()
    colour, ncol = got["colour"], got["ncolours"]
    order = np.zeros(m.nnode, np.int32)
    nchunk = C.c_int32()
    seg_off = np.zeros(64 * 16 + 1, np.int32)
    lib = eq.problem.lib
    assert lib.eqlb_get_launch_order(eq.problem.h, order.ctypes.data_as(cabi.c_int32_p), C.byref(nchunk),
                                     seg_off.ctypes.data_as(cabi.c_int32_p)) == 0
    nchunk = nchunk.value
    assert nchunk == (4 if pipeline else 1)
    ncells = np.diff(m.node_cell_off)
    nf = np.diff(m.node_fct_off)
    last_cell = np.array([m.node_cell[m.node_cell_off[z]:m.node_cell_off[z + 1]].max() for z in range(m.nnode)])
    seg = (last_cell.astype(np.int64) * nchunk // m.ncell) * ncol + colour
    cls = np.where(nf > 16, 3, np.where(nf <= 4, 0, np.where(nf <= 8, 1, 2)))  # single RHS: every patch is eligible
    merged = cls.copy()
    for s_ in range(nchunk * ncol):
        sel = seg == s_
        cc = [int((sel & (cls == c)).sum()) for c in range(3)]
        thr = max(8192, sum(cc) // 16)
        cm = [0, 1, 2]
        if 0 < cc[1] < thr and cc[2] > 0:
            cm[1] = 2
            cc[2] += cc[1]
            cc[1] = 0
        if 0 < cc[0] < thr and cc[1] + cc[2] > 0:
            cm[0] = 1 if cc[1] > 0 else 2
        for c in range(3):
            merged[sel & (cls == c)] = cm[c]
    want = np.lexsort((np.arange(m.nnode), merged, seg))
    assert np.array_equal(order, want)
    assert np.array_equal(seg_off[: nchunk * ncol + 1], np.concatenate([[0], np.cumsum(np.bincount(seg, minlength=nchunk * ncol))]))
