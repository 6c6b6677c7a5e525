"""Reference-element tables (CPU): the role of the reference's test_hierarchic_rt.py
(`python/test/unit/test_hierarchic_rt.py:114-159`): the basis is dual to the defining
functionals, and the exact reference matrices agree with quadrature."""

import numpy as np
import pytest

from dolfinx_eqlb_b200 import tables as tb


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_rt_basis_is_dual_to_functionals(k):
    T = tb.make_tables(k)
    nqf = T.nqf
    # facet moments  sum_n M[f][j][:, n] . phi_i(x_n) = delta
    D = np.einsum("fjdn,fnid->fji", T.M, T.rt_f.reshape(3, nqf, T.nrt, 2))
    E = np.zeros_like(D)
    for f in range(3):
        for j in range(k):
            E[f, j, f * k + j] = 1
    assert np.abs(D - E).max() < 1e-12
    # cell moments: int div(phi_i) x^l y^m and int phi_i . e2 x^l y^m
    rt = T.extra["rt_exact"]
    div_idx, add_idx = tb.rt_moment_indices(k)
    for i, (px, py) in enumerate(rt):
        dv = tb.p_add(tb.p_dx(px), tb.p_dy(py))
        for t, (l, m) in enumerate(div_idx):
            val = float(tb.p_int_cell(tb.p_mul(dv, {(l, m): 1})))
            assert abs(val - (1.0 if i == 3 * k + t else 0.0)) < 1e-14
        for t, (l, m) in enumerate(add_idx):
            val = float(tb.p_int_cell(tb.p_mul(py, {(l, m): 1})))
            assert abs(val - (1.0 if i == 3 * k + len(div_idx) + t else 0.0)) < 1e-14


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_dimensions(k):
    T = tb.make_tables(k)
    assert T.nrt == k * (k + 2)
    assert T.ndg == k * (k + 1) // 2
    assert T.ndiv == k * (k + 1) // 2 - 1
    assert T.nadd == (k - 1) * (k - 2) // 2
    assert T.nqf == (1 if k == 1 else k + 1)


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_quadrature_exactness(k):
    T = tb.make_tables(k)
    deg = 2 if k == 1 else 2 * k + 1
    from math import factorial

    for a in range(deg + 1):
        for b in range(deg + 1 - a):
            exact = factorial(a) * factorial(b) / factorial(a + b + 2)
            assert abs(np.sum(T.qwts * T.qpts[:, 0] ** a * T.qpts[:, 1] ** b) - exact) < 1e-14


@pytest.mark.parametrize("k,p", [(1, 0), (2, 1), (2, 0), (3, 2), (3, 1), (4, 3)])
def test_reference_matrices_match_quadrature(k, p):
    T = tb.make_tables(k, p)
    w = T.qwts
    m00 = np.einsum("q,qi,qj->ij", w, T.rt_q[:, :, 0], T.rt_q[:, :, 0])
    m01 = np.einsum("q,qi,qj->ij", w, T.rt_q[:, :, 0], T.rt_q[:, :, 1])
    m11 = np.einsum("q,qi,qj->ij", w, T.rt_q[:, :, 1], T.rt_q[:, :, 1])
    scale = np.abs(T.rt_mass).max()
    assert np.abs(T.rt_mass[0] - m00).max() < 1e-13 * scale
    assert np.abs(T.rt_mass[1] - (m01 + m01.T)).max() < 1e-13 * scale
    assert np.abs(T.rt_mass[2] - m11).max() < 1e-13 * scale
    # facet moments of hat * dg basis
    nqf = T.nqf
    for f in range(3):
        for v in range(3):
            if v == f:
                continue
            for j in range(k):
                q = np.einsum("n,n,ni->i", T.fwts * T.fpts_s**j, T.hat_f[f * nqf : (f + 1) * nqf, v], T.dg_f[f * nqf : (f + 1) * nqf])
                assert np.abs(q - T.fct_mom[f, v, j]).max() < 1e-14
    # cell moments
    lm = [(0, 0)] + [tuple(x) for x in T.div_lm[: T.ndiv]]
    for v in range(3):
        for t, (l, m) in enumerate(lm):
            wq = w * T.hat_q[:, v] * T.qpts[:, 0] ** l * T.qpts[:, 1] ** m
            if k > 1:
                assert np.abs(wq @ T.dg_q[0] - T.cell_mom_f[v, t]).max() < 1e-14
                assert np.abs(wq @ T.dg_q[1] - T.cell_mom_g[v, t, :, 0]).max() < 1e-13
                assert np.abs(wq @ T.dg_q[2] - T.cell_mom_g[v, t, :, 1]).max() < 1e-13
    # EV tables
    for t, (l, m) in enumerate(lm):
        if k > 1:
            wq = w * T.qpts[:, 0] ** l * T.qpts[:, 1] ** m
            assert np.abs(wq @ T.dg_q[0] - T.dg_mono[t]).max() < 1e-14
    if k > 1:
        h = np.einsum("q,qv,qm,qid->vmid", w, T.hat_q, T.dg_q[0], T.rt_q)
        assert np.abs(h - T.hat_dg_rt).max() < 1e-13 * max(1.0, np.abs(h).max())


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_reversed_facet_transform_is_involution(k):
    R = tb.make_tables(k).trafo
    assert np.abs(R @ R - np.eye(k)).max() < 1e-14


@pytest.mark.parametrize("p", [0, 1, 2, 3])
def test_lagrange_is_nodal(p):
    basis = tb.lagrange_basis(p)
    nodes = tb.lagrange_nodes(p)
    V = np.array([[tb.p_eval(ph, float(x), float(y)) for ph in basis] for (x, y) in nodes])
    assert np.abs(V - np.eye(len(nodes))).max() < 1e-14
    cl = tb.lagrange_facet_closure_dofs(p)
    # closure dofs are exactly the nodes on the facet
    for f in range(3):
        on = []
        for i, (x, y) in enumerate(nodes):
            on_f = [x + y == 1, x == 0, y == 0][f]
            if on_f and p > 0:
                on.append(i)
        if p > 0:
            assert sorted(cl[f]) == on


def test_bad_degrees_raise():
    with pytest.raises(ValueError):
        tb.make_tables(0)
    with pytest.raises(ValueError):
        tb.make_tables(2, 2)
