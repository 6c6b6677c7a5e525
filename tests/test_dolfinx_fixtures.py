"""Parity against fixtures exported from a LIVE dolfinx_eqlb installation
(`tools/export_dolfinx_fixture.py`).  Skipped while `tests/golden/dolfinx/` holds no
fixture - the build image has no DOLFINx; dropping a file there pins the
oracle and the CUDA path against the reference itself (SE path; the reference's EV flux
lives in Basix' Legendre-variant RT space, which is not reproducible offline)."""

import glob
import os

import numpy as np
import pytest

from dolfinx_eqlb_b200 import eqlb, mesh as ms, tables as tb

FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dolfinx", "*.npz")))
needs_fixture = pytest.mark.skipif(not FIXTURES, reason="no DOLFINx fixture in tests/golden/dolfinx (see its README)")


def load(path):
    d = np.load(path)
    k, is_ev, stress, nrhs = (int(v) for v in d["meta"])
    m = ms.build_topology(np.asarray(d["x"], dtype=np.float64), np.asarray(d["cell_node"], dtype=np.int32))
    key = m.fct_node[:, 0].astype(np.int64) * m.nnode + m.fct_node[:, 1]

    def facets(pairs):
        pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
        return np.searchsorted(key, pairs[:, 0] * m.nnode + pairs[:, 1]).astype(np.int32)

    T = tb.make_tables(k)
    G = [np.ascontiguousarray(d["G"][r]).ravel() for r in range(nrhs)]
    F = [np.ascontiguousarray(d["f"][r]).ravel() for r in range(nrhs)]
    return d, m, T, k, bool(is_ev), bool(stress), nrhs, G, F, facets(d["bfct_prime"])


def test_loader_roundtrip(tmp_path):
    """The fixture reader itself (always runs): a fixture written from this repository's own
    mesh/oracle in the exporter's layout is read back and reproduced."""
    from oracle import pyoracle as po

    m0 = ms.crossed_unit_square(3, scramble_seed=4, perturb=0.1)
    T = tb.make_tables(2)
    rng = np.random.default_rng(0)
    G = rng.standard_normal((1, m0.ncell, T.ndg, 2))
    f = rng.standard_normal((1, m0.ncell, T.ndg))
    bf = m0.boundary_facets([1, 2, 3, 4])
    ft = ms.facet_types(m0, [1, 2, 3, 4], [])[None, :]
    sig = po.se_run(m0, T, po.BCData(ft), [G[0].ravel()], [f[0].ravel()])
    p = tmp_path / "self.npz"
    np.savez(p, x=m0.x[:, :2], cell_node=m0.cell_node, G=G, f=f, sigma_cells=np.array(sig).reshape(1, m0.ncell, T.nrt),
             bfct_prime=m0.fct_node[bf], bfct_flux=np.zeros((0, 2), np.int64), meta=np.array([2, 0, 0, 1]))
    d, m, T2, k, is_ev, stress, nrhs, G2, F2, prime = load(str(p))
    assert np.array_equal(m.cell_node, m0.cell_node) and np.array_equal(np.sort(prime), np.sort(bf))
    ft2 = np.zeros((1, m.nfct), np.int8)
    ft2[0, prime] = 1
    sig2 = po.se_run(m, T2, po.BCData(ft2), G2, F2)
    assert np.abs(np.array(sig2).reshape(d["sigma_cells"].shape) - d["sigma_cells"]).max() < 1e-13


@needs_fixture
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_matches_dolfinx(path):
    from oracle import pyoracle as po

    d, m, T, k, is_ev, stress, nrhs, G, F, prime = load(path)
    if is_ev:
        pytest.skip("EV fixtures need Basix' RT basis (DESIGN.md section 6)")
    ft = np.zeros((nrhs, m.nfct), np.int8)
    ft[:, prime] = 1
    sig = po.se_run(m, T, po.BCData(ft), G, F, stress=stress)
    ref = d["sigma_cells"]
    assert np.abs(np.array(sig).reshape(ref.shape) - ref).max() < 1e-10 * np.abs(ref).max()


@needs_fixture
@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_gpu_matches_dolfinx(path):
    d, m, T, k, is_ev, stress, nrhs, G, F, prime = load(path)
    if is_ev:
        pytest.skip("EV fixtures need Basix' RT basis (DESIGN.md section 6)")
    eq = eqlb.FluxEqlbSE(k, m, F, G, equilibrate_stress=stress)
    eq.set_boundary_conditions([prime] * nrhs, [[] for _ in range(nrhs)])
    eq.equilibrate_fluxes()
    ref = d["sigma_cells"]
    assert np.abs(np.array(eq.list_flux).reshape(ref.shape) - ref).max() < 1e-10 * np.abs(ref).max()
