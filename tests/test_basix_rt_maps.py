"""Change of basis hierarchic RT -> Basix "RT" (Legendre variant), `tables.py::basix_rt_legendre_maps`.

Basix is not available offline, so its functionals are restated from their definition (normal moments against
the orthonormal Legendre basis on the facets, moments against the orthonormal Dubiner basis in the interior).
The exact-rational Gram-Schmidt construction of tables.py is checked here against an independent formulation
(numpy Legendre / scipy Jacobi polynomials + numerical quadrature); the GPU kernel against numpy."""

import numpy as np
import pytest
from scipy.special import eval_jacobi, eval_legendre

from dolfinx_eqlb_b200.tables import REF_NORMALS, cell_quadrature, facet_point, gauss_legendre_01, make_tables, p_eval


def dubiner(p, q, x, y):
    xi = np.where(np.abs(1 - y) > 1e-14, 2 * x / (1 - y) - 1, 0.0)
    return eval_legendre(p, xi) * (1 - y) ** p * eval_jacobi(q, 2 * p + 1, 0, 2 * y - 1)


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_maps_are_the_basix_functionals_of_the_hierarchic_basis(k):
    T = make_tables(k)
    rt = T.extra["rt_exact"]
    A, B = T.rt_basix_fct, T.rt_basix_int
    qp, qw = cell_quadrature(2 * k + 2)
    fs, fw = gauss_legendre_01(k + 2)
    rng = np.random.default_rng(k)
    c = rng.standard_normal(T.nrt)

    def v(x, y):
        return np.array([sum(c[j] * p_eval(rt[j][d], x, y) for j in range(T.nrt)) for d in range(2)])

    # facet functionals: int v.n sqrt(2i+1) P_i(2s-1) ds  ==  (A c_facet)_i
    for f in range(3):
        for i in range(k):
            val = sum(w * (v(*facet_point(f, s)) @ np.array(REF_NORMALS[f], float)) * np.sqrt(2 * i + 1) * eval_legendre(i, 2 * s - 1)
                      for s, w in zip(fs, fw))
            assert abs(val - (A @ c[f * k : (f + 1) * k])[i]) < 1e-11 * max(1.0, abs(val))
    # interior functionals: int v_d P_{p,q} (orthonormal, Basix polyset order, direction-major)  ==  (B c)_i
    if k > 1:
        idx = [(p, q) for n in range(k - 1) for q in range(n + 1) for p in [n - q]]
        nsc = len(idx)
        vq = np.array([v(x, y) for x, y in qp])
        for d in range(2):
            for i, (p, q) in enumerate(idx):
                ph = dubiner(p, q, qp[:, 0], qp[:, 1])
                ph = ph / np.sqrt(np.sum(qw * ph * ph))
                val = np.sum(qw * vq[:, d] * ph)
                assert abs(val - (B @ c)[d * nsc + i]) < 1e-10 * max(1.0, abs(val)), (d, p, q)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 2, 3])
def test_device_conversion(k):
    from common import PoissonCase, make_mesh
    from dolfinx_eqlb_b200 import eqlb

    m = make_mesh("crossed", 5, 3, perturb=0.2)  # scrambled: reflected facets
    case = PoissonCase(m, k, [[1, 4]], seed=2, hom=True)
    eq = eqlb.FluxEqlbEV(k, m, case.F, case.G)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    eq.equilibrate_fluxes()
    got = eq.fluxes_in_basix_rt()[0]
    T = case.T
    hier = eq.list_flux[0]
    ncd = k * k - k
    fd = hier[: m.nfct * k].reshape(m.nfct, k)
    want_f = fd @ T.rt_basix_fct.T
    assert np.abs(got[: m.nfct * k].reshape(m.nfct, k) - want_f).max() < 1e-13 * np.abs(want_f).max()
    if ncd:
        R = T.trafo.T  # c_loc = R c_glob on reflected facets
        loc = np.zeros((m.ncell, T.nrt))
        for f in range(3):
            g = fd[m.cell_fct[:, f]]
            refl = m.fct_perms.reshape(m.ncell, 3)[:, f].astype(bool)
            loc[:, f * k : (f + 1) * k] = np.where(refl[:, None], g @ R.T, g)
        loc[:, 3 * k :] = hier[m.nfct * k :].reshape(m.ncell, ncd)
        want_c = loc @ T.rt_basix_int.T
        assert np.abs(got[m.nfct * k :].reshape(m.ncell, ncd) - want_c).max() < 1e-13 * np.abs(want_c).max()
