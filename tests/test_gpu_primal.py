"""GPU: fused input stage (projection of the primal solution on the device) and error estimators
 against independent numpy evaluations of the same definitions
(`lsolver/projection.py:17-77`, `demo/poisson/demo_error_estimation.py:52-122`,
`demo/elasticity/demo_error_estimation.py:49-135`)."""

import numpy as np
import pytest

import fem_mini as fm
from common import make_mesh
from dolfinx_eqlb_b200 import eqlb, tables as tb

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def l2_project_pk_to_dg(m, T, V, fh):
    """numpy L2 projection of a P_k function into DG_p, cell by cell, by quadrature"""
    qp, qw = tb.cell_quadrature(2 * T.k + 2)
    phi, _ = fm.tabulate_scalar(V.basis, qp)
    dgv, _ = fm.tabulate_scalar(T.extra["dg_exact"], qp)
    M = np.einsum("q,qi,qj->ij", qw, dgv, dgv)
    b = np.einsum("q,qi,qn,cn->ci", qw, dgv, phi, fh[V.dofmap])
    return np.linalg.solve(M, b.T).T.reshape(-1)


def primal_case(k, seed=1, scramble=3):
    m = make_mesh("crossed", 5, scramble, perturb=0.2)
    T = tb.make_tables(k)
    rng = np.random.default_rng(seed)
    f_dg = fm.random_dg(rng, m.ncell * T.ndg)
    G, u, V = fm.solve_poisson(m, k, T, f_dg, [1, 2, 3, 4])
    fh = rng.standard_normal(V.ndof)
    return m, T, V, G, u, fh


@pytest.mark.parametrize("k", [1, 2, 3])
def test_projection_of_the_primal_solution(k):
    m, T, V, G, u, fh = primal_case(k)
    eq = eqlb.FluxEqlbSE(k, m, [np.zeros(m.ncell * T.ndg)], [np.zeros(m.ncell * T.ndg * 2)])
    eq.problem.set_primal_space(V.dofmap, V.ndof)
    Gd, Fd = eq.problem.project_primal([u], [fh])
    assert rel(Gd[0], G) < 1e-12
    assert rel(Fd[0], l2_project_pk_to_dg(m, T, V, fh)) < 1e-12


@pytest.mark.parametrize("k,ev", [(2, True), (2, False), (3, True), (1, False)])
def test_fused_run_equals_projection_then_equilibration(k, ev):
    m, T, V, G, u, fh = primal_case(k, seed=2)
    F = l2_project_pk_to_dg(m, T, V, fh)
    cls = eqlb.FluxEqlbEV if ev else eqlb.FluxEqlbSE
    eq = cls(k, m, [F], [G])
    eq.set_boundary_conditions([m.boundary_facets([1, 2, 3, 4])], [[]])
    eq.equilibrate_fluxes()
    eq.problem.set_primal_space(V.dofmap, V.ndof)
    sig = [np.zeros_like(eq.list_flux[0])]
    eq.problem.run_primal(ev, [u], [fh], sig, zeroed=True)
    assert rel(sig[0], eq.list_flux[0]) < 1e-11
    eq.problem.run_primal(ev, [u], [fh], sig)  # accumulates like the reference
    assert rel(sig[0], 2 * eq.list_flux[0]) < 1e-11


@pytest.mark.parametrize("k,ev", [(2, True), (2, False), (3, False)])
def test_poisson_estimator(k, ev):
    m, T, V, G, u, fh = primal_case(k, seed=3)
    F = l2_project_pk_to_dg(m, T, V, fh)
    cls = eqlb.FluxEqlbEV if ev else eqlb.FluxEqlbSE
    eq = cls(k, m, [F], [G])
    eq.set_boundary_conditions([m.boundary_facets([1, 2, 3, 4])], [[]])
    eq.equilibrate_fluxes()
    eq.problem.set_primal_space(V.dofmap, V.ndof)
    e_sig, e_osc = eq.problem.estimate_poisson(eq.list_flux, [u], [fh], ev)
    # numpy evaluation of the same definitions
    sig = fm.conforming_to_drt(m, T, eq.list_flux[0]) if ev else eq.list_flux[0]
    J, K, det = fm.jacobians(m)
    qp, qw = tb.cell_quadrature(2 * k + 2)
    rt, rdiv = fm.tabulate_rt(T, qp)
    phi, dphi = fm.tabulate_scalar(V.basis, qp)
    c = sig.reshape(m.ncell, T.nrt)
    sref = np.einsum("ci,qid->cqd", c, rt)
    s = np.einsum("cab,cqb->cqa", J, sref) / det[:, None, None]
    dv = np.einsum("ci,qi->cq", c, rdiv) / det[:, None]
    gu = np.einsum("cji,qnj,cn->cqi", K, dphi, u[V.dofmap])
    fq = np.einsum("qn,cn->cq", phi, fh[V.dofmap])
    # Laplacian of u_h cell-wise
    lap = np.zeros((m.ncell, len(qw)))
    for n_, ph in enumerate(V.basis):
        hx, hy = tb.p_dx(ph), tb.p_dy(ph)
        H = np.array([[[tb.p_eval(tb.p_dx(hx), x, y), tb.p_eval(tb.p_dy(hx), x, y)],
                       [tb.p_eval(tb.p_dx(hy), x, y), tb.p_eval(tb.p_dy(hy), x, y)]] for x, y in qp])  # [q][2][2] reference Hessian
        Hp = np.einsum("cai,qab,cbj->cqij", K, H, K)
        lap += (Hp[:, :, 0, 0] + Hp[:, :, 1, 1]) * u[V.dofmap[:, n_]][:, None]
    if ev:
        err = gu + s
        res = fq - dv
    else:
        err = s
        res = fq - dv + lap
    w = qw[None, :] * np.abs(det)[:, None]
    x = m.x[m.cell_node][:, :, :2]
    hT = np.sqrt(np.max([((x[:, a] - x[:, b]) ** 2).sum(axis=1) for a, b in ((0, 1), (1, 2), (2, 0))], axis=0))
    want_sig = (w * (err**2).sum(axis=2)).sum(axis=1)
    want_osc = (hT / np.pi) ** 2 * (w * res**2).sum(axis=1)
    assert rel(e_sig[0], want_sig) < 1e-11
    assert rel(e_osc[0], want_osc) < 1e-10
    if not ev:
        # consistency with eqlb_flux_l2norm
        assert rel(e_sig[0], eq.flux_l2norm()[0]) < 1e-11


def test_elasticity_estimator():
    from test_gpu_stress import elasticity_case

    k = 2
    m = make_mesh("crossed", 5, 2, perturb=0.2)
    T, G, f, bfp, bcs, neu = elasticity_case(m, k, [], seed=3)
    eq = eqlb.FluxEqlbSE(k, m, f, G, equilibrate_stress=True, estimate_korn_constant=True)
    eq.set_boundary_conditions(bfp, bcs)
    eq.equilibrate_fluxes()
    V = fm.LagrangeSpace(m, k)
    rng = np.random.default_rng(4)
    fh = [rng.standard_normal(V.ndof) for _ in range(2)]
    eq.problem.set_primal_space(V.dofmap, V.ndof)
    korn = eq.get_korn_constants()
    pi_1 = 1.5
    sig_h = [-g for g in G]  # projected stress rows (the equilibration is fed with G = -sigma_h)
    eta = eq.problem.estimate_elasticity(eq.list_flux, sig_h, fh, korn, pi_1)
    J, K, det = fm.jacobians(m)
    qp, qw = tb.cell_quadrature(2 * k + 2)
    rt, rdiv = fm.tabulate_rt(T, qp)
    phi, _ = fm.tabulate_scalar(V.basis, qp)
    dgv, dgg = fm.tabulate_scalar(T.extra["dg_exact"], qp)
    w = qw[None, :] * np.abs(det)[:, None]
    r, dv, dsh, fq = [], [], [], []
    for row in range(2):
        c = eq.list_flux[row].reshape(m.ncell, T.nrt)
        sref = np.einsum("ci,qid->cqd", c, rt)
        r.append(np.einsum("cab,cqb->cqa", J, sref) / det[:, None, None])
        dv.append(np.einsum("ci,qi->cq", c, rdiv) / det[:, None])
        sh = sig_h[row].reshape(m.ncell, T.ndg, 2)
        gphys = np.einsum("cji,qnj->cqni", K, dgg)
        dsh.append(np.einsum("cqni,cni->cq", gphys, sh))
        fq.append(np.einsum("qn,cn->cq", phi, fh[row][V.dofmap]))
    tr = r[0][:, :, 0] + r[1][:, :, 1]
    ctr = pi_1 / (2 + 2 * pi_1)
    e0 = (w * 0.5 * ((r[0] ** 2).sum(axis=2) + (r[1] ** 2).sum(axis=2) - ctr * tr**2)).sum(axis=1)
    e1 = (w * (0.5 * korn[:, None] * (r[0][:, :, 1] - r[1][:, :, 0])) ** 2).sum(axis=1)
    x = m.x[m.cell_node][:, :, :2]
    hT = np.sqrt(np.max([((x[:, a] - x[:, b]) ** 2).sum(axis=1) for a, b in ((0, 1), (1, 2), (2, 0))], axis=0))
    e2 = (korn * hT / np.pi) ** 2 * (w * ((fq[0] + dsh[0] + dv[0]) ** 2 + (fq[1] + dsh[1] + dv[1]) ** 2)).sum(axis=1)
    assert rel(eta[0], e0) < 1e-11
    assert rel(eta[1], e1) < 1e-10
    assert rel(eta[2], e2) < 1e-10
