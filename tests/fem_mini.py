"""Tiny FE toolbox for the tests (numpy/scipy, small meshes only).

Produces Galerkin-orthogonal input data (a real P_q Poisson / elasticity solve)
and evaluates the acceptance invariants of the reference's test-suite
(`python/dolfinx_eqlb/eqlb/check_eqlb_conditions.py:183-291` divergence,
`:294-359` jump, `:90-180` boundary conditions, `:476-521` weak symmetry).
"""

from __future__ import annotations

from fractions import Fraction as Fr

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from dolfinx_eqlb_b200 import tables as tb
from dolfinx_eqlb_b200.mesh import FACET_VERTS

REFV = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
NOUT = np.array([False, True, False])


def tabulate_scalar(basis, pts):
    """values [npts][n], grads [npts][n][2] of exact scalar polynomials."""
    n = len(basis)
    v = np.zeros((len(pts), n))
    g = np.zeros((len(pts), n, 2))
    for i, ph in enumerate(basis):
        px, py = tb.p_dx(ph), tb.p_dy(ph)
        for q, (x, y) in enumerate(pts):
            v[q, i] = tb.p_eval(ph, x, y)
            g[q, i, 0] = tb.p_eval(px, x, y)
            g[q, i, 1] = tb.p_eval(py, x, y)
    return v, g


def tabulate_rt(T, pts):
    rt = T.extra["rt_exact"]
    v = np.zeros((len(pts), len(rt), 2))
    d = np.zeros((len(pts), len(rt)))
    for i, (px, py) in enumerate(rt):
        dv = tb.p_add(tb.p_dx(px), tb.p_dy(py))
        for q, (x, y) in enumerate(pts):
            v[q, i, 0] = tb.p_eval(px, x, y)
            v[q, i, 1] = tb.p_eval(py, x, y)
            d[q, i] = tb.p_eval(dv, x, y)
    return v, d


def jacobians(mesh):
    x = mesh.x[:, :2]
    cn = mesh.cell_node
    x0, x1, x2 = x[cn[:, 0]], x[cn[:, 1]], x[cn[:, 2]]
    J = np.stack([x1 - x0, x2 - x0], axis=2)  # [ncell][2][2], columns = edges
    det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
    K = np.linalg.inv(J)
    return J, K, det


class LagrangeSpace:
    """Continuous P_q (q <= 3) on a Mesh, Basix cell-local DOF order."""

    def __init__(self, mesh, q):
        assert 1 <= q <= 3
        self.mesh, self.q = mesh, q
        self.basis = tb.lagrange_basis(q)
        self.nloc = len(self.basis)
        ne = q - 1
        nint = self.nloc - 3 - 3 * ne
        dm = np.zeros((mesh.ncell, self.nloc), dtype=np.int64)
        dm[:, :3] = mesh.cell_node
        for f in range(3):
            a, b = FACET_VERTS[f]
            rev = mesh.cell_node[:, a] > mesh.cell_node[:, b]
            for i in range(ne):
                ii = np.where(rev, ne - 1 - i, i)
                dm[:, 3 + f * ne + i] = mesh.nnode + mesh.cell_fct[:, f].astype(np.int64) * ne + ii
        for i in range(nint):
            dm[:, 3 + 3 * ne + i] = mesh.nnode + mesh.nfct * ne + np.arange(mesh.ncell) * nint + i
        self.dofmap = dm
        self.ndof = mesh.nnode + mesh.nfct * ne + mesh.ncell * nint

    def boundary_dofs(self, sides):
        m = self.mesh
        fcts = m.boundary_facets(sides)
        dofs = set(m.fct_node[fcts].ravel().tolist())
        ne = self.q - 1
        for f in fcts:
            for i in range(ne):
                dofs.add(m.nnode + int(f) * ne + i)
        return np.array(sorted(dofs), dtype=np.int64)


def solve_poisson(mesh, q, T, f_dg, dirichlet_sides, neumann=None, seed=0):
    """-div grad u = f with u = 0 on the Dirichlet sides and sigma.n = g on the
    Neumann facets (sigma = -grad u).  f_dg: DG_p coefficients [ncell*ndg].
    neumann: dict facet -> polynomial coefficients of g in the facet parameter of
    the adjacent cell.  Returns G = -grad u_h as DG_p coefficients [ncell*ndg*2]."""
    V = LagrangeSpace(mesh, q)
    J, K, det = jacobians(mesh)
    qp, qw = tb.cell_quadrature(2 * max(q, T.p + 1))
    phi, dphi = tabulate_scalar(V.basis, qp)
    dgv, _ = tabulate_scalar(T.extra["dg_exact"], qp)
    # physical gradients: K^T grad_ref
    gphys = np.einsum("cji,qnj->cqni", K, dphi)  # [c][q][n][2]
    Ke = np.einsum("q,c,cqni,cqmi->cnm", qw, np.abs(det), gphys, gphys)
    fq = np.einsum("qi,ci->cq", dgv, f_dg.reshape(mesh.ncell, T.ndg))
    be = np.einsum("q,c,cq,qn->cn", qw, np.abs(det), fq, phi)
    rows = np.repeat(V.dofmap, V.nloc, axis=1).ravel()
    cols = np.tile(V.dofmap, (1, V.nloc)).ravel()
    A = sp.csr_matrix((Ke.ravel(), (rows, cols)), shape=(V.ndof, V.ndof))
    b = np.zeros(V.ndof)
    np.add.at(b, V.dofmap.ravel(), be.ravel())
    if neumann:
        gs, gwt = tb.gauss_legendre_01(q + T.k + 1)
        for fct, gc in neumann.items():
            c = mesh.fct_cell[mesh.fct_cell_off[fct]]
            lf = int(np.nonzero(mesh.cell_fct[c] == fct)[0][0])
            a_, b_ = FACET_VERTS[lf]
            pts = [tb.facet_point(lf, s) for s in gs]
            pv, _ = tabulate_scalar(V.basis, pts)
            length = np.linalg.norm(mesh.x[mesh.cell_node[c, a_], :2] - mesh.x[mesh.cell_node[c, b_], :2])
            gval = sum(gc[j] * gs**j for j in range(len(gc)))
            b[V.dofmap[c]] -= length * (gwt * gval) @ pv
    bd = V.boundary_dofs(dirichlet_sides)
    free = np.setdiff1d(np.arange(V.ndof), bd)
    u = np.zeros(V.ndof)
    u[free] = spla.spsolve(A[free][:, free].tocsc(), b[free])
    # G = -grad u_h, nodal interpolation into DG_p (exact when p >= q-1)
    nodes = [(float(a), float(b)) for a, b in tb.lagrange_nodes(T.p)]
    _, dn = tabulate_scalar(V.basis, nodes)
    gn = np.einsum("cji,qnj->cqni", K, dn)
    G = -np.einsum("cqni,cn->cqi", gn, u[V.dofmap])
    return G.reshape(-1), u, V


def neumann_bflux(mesh, T, neumann):
    """Boundary function (DRT vector) + local facet ids for prescribed outward normal
    fluxes g(s): dof_j = prefactor * |E| int_0^1 g s^j ds, prefactor = +-sgn(detJ)
    (normal-orientation, `se/solve_patch_semiexplt.hpp:392-396`)."""
    _, _, det = jacobians(mesh)
    k = T.k
    bfl = np.zeros(mesh.ncell * T.nrt)
    lfi = np.zeros(mesh.nfct, dtype=np.int8)
    for fct, gc in neumann.items():
        c = mesh.fct_cell[mesh.fct_cell_off[fct]]
        lf = int(np.nonzero(mesh.cell_fct[c] == fct)[0][0])
        a_, b_ = FACET_VERTS[lf]
        length = np.linalg.norm(mesh.x[mesh.cell_node[c, a_], :2] - mesh.x[mesh.cell_node[c, b_], :2])
        pre = (1.0 if NOUT[lf] else -1.0) * np.sign(det[c])
        lfi[fct] = lf
        for j in range(k):
            bfl[c * T.nrt + lf * k + j] = pre * length * sum(gc[i] / (i + j + 1) for i in range(len(gc)))
    return bfl, lfi


# ------------------------------------------------------------------ invariants


def check_divergence(mesh, T, sigma, G, f):
    """max |div(sigma + G) - f| over cell quadrature points, relative."""
    J, K, det = jacobians(mesh)
    pts = T.qpts
    _, dref = tabulate_rt(T, pts)
    dgv, dgg = tabulate_scalar(T.extra["dg_exact"], pts)
    div_s = np.einsum("qi,ci->cq", dref, sigma.reshape(mesh.ncell, T.nrt)) / det[:, None]
    gphys = np.einsum("cji,qnj->cqni", K, dgg)
    div_g = np.einsum("cqni,cni->cq", gphys, G.reshape(mesh.ncell, T.ndg, 2))
    fq = np.einsum("qi,ci->cq", dgv, f.reshape(mesh.ncell, T.ndg))
    scale = max(np.abs(fq).max(), np.abs(div_g).max(), 1e-30)
    return np.abs(div_s + div_g - fq).max() / scale


def _flux_at(mesh, T, sigma, G, cell, ref_pts, J, det):
    rv, _ = tabulate_rt(T, ref_pts)
    dgv, _ = tabulate_scalar(T.extra["dg_exact"], ref_pts)
    s_ref = np.einsum("qid,i->qd", rv, sigma.reshape(mesh.ncell, T.nrt)[cell])
    s = np.einsum("ij,qj->qi", J[cell], s_ref) / det[cell]
    g = np.einsum("qi,id->qd", dgv, G.reshape(mesh.ncell, T.ndg, 2)[cell])
    return s + g


def check_jump(mesh, T, sigma, G, spts=(0.15, 0.5, 0.8)):
    """max normal jump of (sigma + G) over interior facets, relative."""
    J, K, det = jacobians(mesh)
    worst, scale = 0.0, 1e-30
    for fct in range(mesh.nfct):
        cells = mesh.fct_cell[mesh.fct_cell_off[fct] : mesh.fct_cell_off[fct + 1]]
        if len(cells) != 2:
            continue
        lo, hi = mesh.fct_node[fct]
        tvec = mesh.x[hi, :2] - mesh.x[lo, :2]
        nvec = np.array([tvec[1], -tvec[0]])
        vals = []
        for c in cells:
            lf = int(np.nonzero(mesh.cell_fct[c] == fct)[0][0])
            a_, b_ = FACET_VERTS[lf]
            if mesh.cell_node[c, a_] == lo:
                ref = [REFV[a_] + s * (REFV[b_] - REFV[a_]) for s in spts]
            else:
                ref = [REFV[b_] + s * (REFV[a_] - REFV[b_]) for s in spts]
            vals.append(_flux_at(mesh, T, sigma, G, c, ref, J, det) @ nvec)
        worst = max(worst, np.abs(vals[0] - vals[1]).max())
        scale = max(scale, np.abs(vals[0]).max())
    return worst / scale


def facet_moments_of_G(mesh, T, G, cell, lf, J, K, det):
    """L_j(G|cell) = sum_n M[lf][j][:, n] . (detJ K G(x_n))  (pull-back + M)."""
    nqf = T.nqf
    gv = np.einsum("ni,id->nd", T.dg_f[lf * nqf : (lf + 1) * nqf], G.reshape(mesh.ncell, T.ndg, 2)[cell])
    gref = det[cell] * np.einsum("ij,nj->ni", K[cell], gv)
    return np.einsum("jdn,nd->j", T.M[lf], gref)


def check_bc(mesh, T, sigma, G, bflux, neumann):
    J, K, det = jacobians(mesh)
    worst, scale = 0.0, 1e-30
    k = T.k
    for fct in neumann:
        c = mesh.fct_cell[mesh.fct_cell_off[fct]]
        lf = int(np.nonzero(mesh.cell_fct[c] == fct)[0][0])
        mom = facet_moments_of_G(mesh, T, G, c, lf, J, K, det)
        sl = slice(c * T.nrt + lf * k, c * T.nrt + lf * k + k)
        worst = max(worst, np.abs(sigma[sl] + mom - bflux[sl]).max())
        scale = max(scale, np.abs(bflux[sl]).max(), np.abs(mom).max())
    return worst / scale


def random_dg(rng, n):
    """`2*(U[0,1)+0.1)` like `python/test/unit/testcase_general.py:118-131`."""
    return 2.0 * (rng.random(n) + 0.1)


def conforming_to_drt(mesh, T, sig_g):
    """Conforming hierarchic-RT vector ([fct*k+j][nfct*k + cell*(k^2-k)+i]) -> cell-local
    DRT coefficients: c_loc = R^T c_glob on reflected facets."""
    k, nrt = T.k, T.nrt
    ncd = k * k - k
    out = np.zeros((mesh.ncell, nrt))
    R = T.trafo
    for f in range(3):
        cg = sig_g[: mesh.nfct * k].reshape(mesh.nfct, k)[mesh.cell_fct[:, f]]
        refl = mesh.fct_perms.reshape(-1, 3)[:, f].astype(bool)
        out[:, f * k : (f + 1) * k] = np.where(refl[:, None], cg @ R, cg)
    if ncd:
        out[:, 3 * k :] = sig_g[mesh.nfct * k :].reshape(mesh.ncell, ncd)
    return out.reshape(-1)


def solve_elasticity(mesh, q, T, f_dg, dirichlet_sides, neumann=None, mu=1.0, lam=1.5):
    """-div sigma(u) = f, sigma = 2 mu eps(u) + lam tr(eps) I, u = 0 on the Dirichlet sides,
    (-sigma) n = g on the Neumann facets (rows of the 'flux' -sigma).  f_dg: two DG_p
    vectors; neumann: two dicts facet -> coefficients (one per row).  Returns the two
    projected flux rows G_r = -sigma_r(u_h) as DG_p coefficient vectors."""
    V = LagrangeSpace(mesh, q)
    J, K, det = jacobians(mesh)
    qp, qw = tb.cell_quadrature(2 * max(q, T.p + 1))
    phi, dphi = tabulate_scalar(V.basis, qp)
    dgv, _ = tabulate_scalar(T.extra["dg_exact"], qp)
    g = np.einsum("cji,qnj->cqni", K, dphi)  # [c][q][n][2] physical gradients
    nl = V.nloc
    wdet = np.einsum("q,c->cq", qw, np.abs(det))
    # B-matrix free assembly: eps(u):eps(v) and div div
    Ke = np.zeros((mesh.ncell, 2 * nl, 2 * nl))
    for a in range(2):
        for b in range(2):
            # u component a, v component b
            blk = lam * np.einsum("cq,cqn,cqm->cnm", wdet, g[..., b], g[..., a])  # div v * div u
            blk += mu * np.einsum("cq,cqn,cqm->cnm", wdet, g[..., a], g[..., b])  # grad_a v_b * grad_b u_a
            if a == b:
                blk += mu * np.einsum("cq,cqni,cqmi->cnm", wdet, g, g)
            Ke[:, b::2, a::2] = blk
    dm2 = np.stack([2 * V.dofmap, 2 * V.dofmap + 1], axis=2).reshape(mesh.ncell, 2 * nl)
    rows = np.repeat(dm2, 2 * nl, axis=1).ravel()
    cols = np.tile(dm2, (1, 2 * nl)).ravel()
    A = sp.csr_matrix((Ke.ravel(), (rows, cols)), shape=(2 * V.ndof, 2 * V.ndof))
    b = np.zeros(2 * V.ndof)
    for r in range(2):
        fq = np.einsum("qi,ci->cq", dgv, f_dg[r].reshape(mesh.ncell, T.ndg))
        be = np.einsum("cq,cq,qn->cn", wdet, fq, phi)
        np.add.at(b, (2 * V.dofmap + r).ravel(), be.ravel())
        if neumann and neumann[r]:
            gs, gwt = tb.gauss_legendre_01(q + T.k + 1)
            for fct, gc in neumann[r].items():
                c = mesh.fct_cell[mesh.fct_cell_off[fct]]
                lf = int(np.nonzero(mesh.cell_fct[c] == fct)[0][0])
                a_, b_ = FACET_VERTS[lf]
                pts = [tb.facet_point(lf, s) for s in gs]
                pv, _ = tabulate_scalar(V.basis, pts)
                length = np.linalg.norm(mesh.x[mesh.cell_node[c, a_], :2] - mesh.x[mesh.cell_node[c, b_], :2])
                gval = sum(gc[j] * gs**j for j in range(len(gc)))
                b[2 * V.dofmap[c] + r] -= length * (gwt * gval) @ pv
    bd = V.boundary_dofs(dirichlet_sides)
    bd2 = np.concatenate([2 * bd, 2 * bd + 1])
    free = np.setdiff1d(np.arange(2 * V.ndof), bd2)
    u = np.zeros(2 * V.ndof)
    u[free] = spla.spsolve(A[free][:, free].tocsc(), b[free])
    nodes = [(float(a), float(b_)) for a, b_ in tb.lagrange_nodes(T.p)]
    _, dn = tabulate_scalar(V.basis, nodes)
    gn = np.einsum("cji,qnj->cqni", K, dn)  # [c][node][n][2]
    ux = np.einsum("cqni,cn->cqi", gn, u[2 * V.dofmap])  # grad u_x
    uy = np.einsum("cqni,cn->cqi", gn, u[2 * V.dofmap + 1])
    tr = ux[..., 0] + uy[..., 1]
    s00 = 2 * mu * ux[..., 0] + lam * tr
    s11 = 2 * mu * uy[..., 1] + lam * tr
    s01 = mu * (ux[..., 1] + uy[..., 0])
    G0 = -np.stack([s00, s01], axis=2).reshape(-1)
    G1 = -np.stack([s01, s11], axis=2).reshape(-1)
    return [G0, G1]


def check_weak_symmetry(mesh, T, sig0, sig1):
    """max_z |int (sigma_01 - sigma_10) hat_z| relative (check_eqlb_conditions.py:476-521)."""
    J, K, det = jacobians(mesh)
    rv, _ = tabulate_rt(T, T.qpts)
    hat = T.hat_q
    c0 = sig0.reshape(mesh.ncell, T.nrt)
    c1 = sig1.reshape(mesh.ncell, T.nrt)
    s0 = np.einsum("cij,qnj,cn->cqi", J, rv, c0) / det[:, None, None]
    s1 = np.einsum("cij,qnj,cn->cqi", J, rv, c1) / det[:, None, None]
    asym = s0[..., 1] - s1[..., 0]
    loc = np.einsum("q,c,cq,qv->cv", T.qwts, np.abs(det), asym, hat)
    tot = np.zeros(mesh.nnode)
    np.add.at(tot, mesh.cell_node.ravel(), loc.ravel())
    scale = np.einsum("q,c,cq->c", T.qwts, np.abs(det), np.abs(s0[..., 1]) + np.abs(s1[..., 0])).max()
    return np.abs(tot).max() / max(scale, 1e-300)


def cell_l2norm_sq(mesh, T, sigma):
    """Cell-wise ||sigma||^2_{L2(T)} of a DRT_k function by quadrature of the Piola-mapped basis
    (J phi_hat / detJ): the UFL form `dot(sigma, sigma) * v * dx` with v in DG0."""
    J, K, det = jacobians(mesh)
    qp, qw = tb.cell_quadrature(2 * T.k)
    phi, _ = tabulate_rt(T, qp)  # [q][nrt][2] reference values
    c = sigma.reshape(mesh.ncell, T.nrt)
    ref = np.einsum("ci,qid->cqd", c, phi)  # reference-cell combination
    phys = np.einsum("cde,cqe->cqd", J, ref) / det[:, None, None]
    return np.einsum("q,cqd,cqd->c", qw, phys, phys) * np.abs(det)
