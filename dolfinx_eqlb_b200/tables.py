"""Reference-element tables for the equilibration hot path (setup-time, host).

The reference obtains these tables from Basix 0.6.0 (third party, absent from the
reference tree): quadrature `base/QuadratureRule.hpp:50-74`, tabulation of the
hierarchic RT / DG / P1 elements `se/KernelData.cpp:84-197`, the facet
interpolation matrix `base/KernelData.cpp:191-268` and the reversed-facet
transform `se/KernelData.cpp:46-64`.  Here everything is built from the
*definition* of the elements in exact rational arithmetic:

* hierarchic RT_k: span(RT_k) + the functionals of
  `python/dolfinx_eqlb/elmtlib/e_raviart_thomas.py:82-124`
* DG_p / P1: equispaced nodal Lagrange in Basix' DOF order
  (vertices, edge interiors e0,e1,e2 low->high vertex, interior)
* reference triangle (0,0),(1,0),(0,1); facet f opposite vertex f, parameterised
  (1-s,s),(0,s),(s,0); reference normals (-1,-1),(-1,0),(0,1)
  (`e_raviart_thomas.py:77-82`); "normal is outward" = {False, True, False}
  (`base/KernelData.cpp:61`).

Two families of tables come out:

1. *quadrature-style* tables (basis functions tabulated at points) - exactly
   what `se::KernelData` holds; consumed by the CPU oracle which keeps the
   reference's quadrature loops.
2. *reference-matrix* tables (exact integrals of products of reference basis
   functions) - consumed by the CUDA kernels, which use the affine-cell
   identities  M_c = 1/|detJ| sum_ab (J^T J)_ab Mhat_ab  etc. instead of
   quadrature loops ("SE kernel").
"""

from __future__ import annotations

from dataclasses import dataclass, field
from fractions import Fraction as Fr
from math import comb, factorial

import numpy as np

# --------------------------------------------------------------------------
# exact bivariate polynomials: dict {(a, b): Fraction}  ==  sum c x^a y^b
# --------------------------------------------------------------------------


def p_add(p, q, s=Fr(1)):
    r = dict(p)
    for key, c in q.items():
        r[key] = r.get(key, Fr(0)) + s * c
    return {k: v for k, v in r.items() if v != 0}


def p_scale(p, s):
    return {k: v * s for k, v in p.items() if v * s != 0}


def p_mul(p, q):
    r = {}
    for (a, b), c in p.items():
        for (d, e), f in q.items():
            key = (a + d, b + e)
            r[key] = r.get(key, Fr(0)) + c * f
    return {k: v for k, v in r.items() if v != 0}


def p_dx(p):
    return {(a - 1, b): c * a for (a, b), c in p.items() if a > 0}


def p_dy(p):
    return {(a, b - 1): c * b for (a, b), c in p.items() if b > 0}


def p_int_cell(p):
    """Integral over the reference triangle: int x^a y^b = a! b! / (a+b+2)!"""
    s = Fr(0)
    for (a, b), c in p.items():
        s += c * Fr(factorial(a) * factorial(b), factorial(a + b + 2))
    return s


def p_eval(p, x, y):
    """Evaluate at the float point (x, y) in exact rational arithmetic and round
    once, so tabulated values carry no cancellation error (the hierarchic basis
    has large monomial coefficients for k >= 3)."""
    xf, yf = Fr(float(x)), Fr(float(y))
    s = Fr(0)
    for (a, b), c in p.items():
        s += c * xf**a * yf**b
    return float(s)


def p_on_facet(p, f):
    """Restrict to reference facet f -> univariate {deg: Fraction} in s.

    f0: (1-s, s), f1: (0, s), f2: (s, 0)  (`e_raviart_thomas.py:77-79`)."""
    r = {}
    for (a, b), c in p.items():
        if f == 0:
            # (1-s)^a s^b
            for i in range(a + 1):
                d = b + i
                r[d] = r.get(d, Fr(0)) + c * comb(a, i) * (-1) ** i
        elif f == 1:
            if a == 0:
                r[b] = r.get(b, Fr(0)) + c
        else:
            if b == 0:
                r[a] = r.get(a, Fr(0)) + c
    return {k: v for k, v in r.items() if v != 0}


def u_mul(p, q):
    r = {}
    for a, c in p.items():
        for b, d in q.items():
            r[a + b] = r.get(a + b, Fr(0)) + c * d
    return r


def u_int01(p):
    return sum((c * Fr(1, a + 1) for a, c in p.items()), Fr(0))


def solve_exact(A, B):
    """Solve A X = B over the rationals (Gauss-Jordan with row pivoting)."""
    n = len(A)
    m = len(B[0])
    M = [list(A[i]) + list(B[i]) for i in range(n)]
    for c in range(n):
        piv = next(r for r in range(c, n) if M[r][c] != 0)
        M[c], M[piv] = M[piv], M[c]
        inv = Fr(1) / M[c][c]
        M[c] = [v * inv for v in M[c]]
        for r in range(n):
            if r != c and M[r][c] != 0:
                fac = M[r][c]
                M[r] = [vr - fac * vc for vr, vc in zip(M[r], M[c])]
    return [row[n:] for row in M]


# --------------------------------------------------------------------------
# reference geometry
# --------------------------------------------------------------------------
REF_NORMALS = ((-1, -1), (-1, 0), (0, 1))
NORMAL_IS_OUTWARD = (False, True, False)
FACET_VERTS = ((1, 2), (0, 2), (0, 1))  # local vertices of facet f, s=0 -> s=1
HAT = ({(0, 0): Fr(1), (1, 0): Fr(-1), (0, 1): Fr(-1)}, {(1, 0): Fr(1)}, {(0, 1): Fr(1)})


def facet_point(f, s):
    return ((1.0 - s, s), (0.0, s), (s, 0.0))[f]


# --------------------------------------------------------------------------
# element definitions
# --------------------------------------------------------------------------


def lagrange_nodes(p):
    """Equispaced P_p nodes in Basix DOF order (as exact Fractions)."""
    if p == 0:
        return [(Fr(1, 3), Fr(1, 3))]
    v = [(Fr(0), Fr(0)), (Fr(1), Fr(0)), (Fr(0), Fr(1))]
    nodes = list(v)
    for f in range(3):
        a, b = FACET_VERTS[f]
        for i in range(1, p):
            t = Fr(i, p)
            nodes.append((v[a][0] + t * (v[b][0] - v[a][0]), v[a][1] + t * (v[b][1] - v[a][1])))
    for j in range(1, p):
        for i in range(1, p - j):
            nodes.append((Fr(i, p), Fr(j, p)))
    return nodes


def lagrange_basis(p):
    """Nodal basis of P_p as exact polynomials."""
    nodes = lagrange_nodes(p)
    monos = [(a, d - a) for d in range(p + 1) for a in range(d, -1, -1)]
    n = len(nodes)
    assert n == len(monos)
    V = [[x**a * y**b for (a, b) in monos] for (x, y) in nodes]
    eye = [[Fr(int(i == j)) for j in range(n)] for i in range(n)]
    C = solve_exact(V, eye)  # C[m][i] = coeff of monomial m in basis i
    return [{monos[m]: C[m][i] for m in range(n) if C[m][i] != 0} for i in range(n)]


def lagrange_facet_closure_dofs(p):
    """entity_closure_dofs[1][f] of P_p (`se/Patch.hpp:420,434,843-850`)."""
    if p == 0:
        return [[0], [0], [0]]
    out = []
    for f in range(3):
        a, b = FACET_VERTS[f]
        out.append([a, b] + [3 + f * (p - 1) + i for i in range(p - 1)])
    return out


def rt_moment_indices(k):
    """(l, m) exponents of the divergence moments and of the e_2 moments in the
    order of `e_raviart_thomas.py:100-124`."""
    div = [(l, m) for l in range(k) for m in range(k - l) if l + m >= 1]
    add = [(l, m) for l in range(1, k - 1) for m in range(k - 1 - l)]
    return div, add


def hierarchic_rt_basis(k):
    """Hierarchic RT_k basis (list of (px, py) exact polynomials) in the cell-
    local DOF order [f0: s^0..s^(k-1)] [f1] [f2] [div moments] [e2 moments]."""
    span = []
    for d in range(2):
        for deg in range(k):
            for a in range(deg, -1, -1):
                mono = {(a, deg - a): Fr(1)}
                span.append((mono, {}) if d == 0 else ({}, mono))
    for a in range(k - 1, -1, -1):
        b = k - 1 - a
        span.append(({(a + 1, b): Fr(1)}, {(a, b + 1): Fr(1)}))
    n = k * (k + 2)
    assert len(span) == n

    div_idx, add_idx = rt_moment_indices(k)

    def functionals(v):
        px, py = v
        vals = []
        for f in range(3):
            nx, ny = REF_NORMALS[f]
            vn = p_add(p_scale(px, Fr(nx)), p_scale(py, Fr(ny)))
            vn_s = p_on_facet(vn, f)
            for j in range(k):
                vals.append(u_int01(u_mul(vn_s, {j: Fr(1)})))
        dv = p_add(p_dx(px), p_dy(py))
        for l, m in div_idx:
            vals.append(p_int_cell(p_mul(dv, {(l, m): Fr(1)})))
        for l, m in add_idx:
            vals.append(p_int_cell(p_mul(py, {(l, m): Fr(1)})))
        return vals

    D = [functionals(v) for v in span]  # D[j][l] = L_l(m_j)
    # basis_i = sum_j c[j][i] m_j with L_l(basis_i) = delta_li  ->  D^T c = I
    DT = [[D[j][l] for j in range(n)] for l in range(n)]
    eye = [[Fr(int(i == j)) for j in range(n)] for i in range(n)]
    C = solve_exact(DT, eye)
    basis = []
    for i in range(n):
        px, py = {}, {}
        for j in range(n):
            if C[j][i] != 0:
                px = p_add(px, span[j][0], C[j][i])
                py = p_add(py, span[j][1], C[j][i])
        basis.append((px, py))
    return basis


# --------------------------------------------------------------------------
# quadrature
# --------------------------------------------------------------------------


def gauss_legendre_01(m):
    x, w = np.polynomial.legendre.leggauss(m)
    return 0.5 + 0.5 * x, 0.5 * w


def cell_quadrature(degree):
    """Rule on the reference triangle exact to `degree`.

    Basix' default (Xiao-Gimbutas) tables are not available offline; any rule of
    sufficient exactness gives the same sums to rounding because all integrands on
    the hot path are polynomials.  degree<=2: 3-point rule; degree<=5:
    Radon's 7-point rule; else a collapsed Gauss-Jacobi (Stroud conical) rule."""
    if degree <= 2:
        pts = np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]])
        wts = np.full(3, 1 / 6)
        return pts, wts
    if degree <= 5:
        s15 = np.sqrt(15.0)
        a, b = (6 - s15) / 21, (6 + s15) / 21
        wa, wb = (155 - s15) / 2400, (155 + s15) / 2400
        pts = np.array(
            [[1 / 3, 1 / 3], [a, a], [1 - 2 * a, a], [a, 1 - 2 * a], [b, b], [1 - 2 * b, b], [b, 1 - 2 * b]]
        )
        wts = np.array([9 / 80, wa, wa, wa, wb, wb, wb])
        return pts, wts
    from scipy.special import roots_jacobi

    m = (degree + 2) // 2
    xa, wa = roots_jacobi(m, 1.0, 0.0)  # weight (1-x) on [-1,1]
    xb, wb = np.polynomial.legendre.leggauss(m)
    pts, wts = [], []
    for i in range(m):
        u = 0.5 * (1 + xa[i])
        for j in range(m):
            v = 0.5 * (1 + xb[j])
            pts.append([u, (1 - u) * v])
            wts.append(wa[i] * wb[j] * 0.125)
    return np.array(pts), np.array(wts)


# --------------------------------------------------------------------------
# the table bundle
# --------------------------------------------------------------------------


@dataclass
class Tables:
    k: int  # flux degree (Basix RT degree)
    p: int  # degree of the projected flux / RHS (DG_p)
    nrt: int
    ndg: int
    ndg_fct: int
    nq: int
    nqf: int
    ndiv: int
    nadd: int
    # quadrature style (oracle)
    qpts: np.ndarray = None  # [nq][2]
    qwts: np.ndarray = None  # [nq]
    fpts_s: np.ndarray = None  # [nqf]
    fwts: np.ndarray = None  # [nqf]
    M: np.ndarray = None  # [3][k][2][nqf]
    rt_q: np.ndarray = None  # [nq][nrt][2]
    rt_f: np.ndarray = None  # [3*nqf][nrt][2]
    dg_q: np.ndarray = None  # [3][nq][ndg] value, d/dx, d/dy
    dg_f: np.ndarray = None  # [3*nqf][ndg]
    hat_q: np.ndarray = None  # [nq][3]
    hat_f: np.ndarray = None  # [3*nqf][3]
    trafo: np.ndarray = None  # [k][k]
    fct_closure: np.ndarray = None  # [3][ndg_fct] int32
    div_lm: np.ndarray = None  # [ndiv][2] int32
    # reference-matrix style (CUDA)
    rt_mass: np.ndarray = None  # [3][nrt][nrt]: M00, M01+M01^T, M11
    fct_mom: np.ndarray = None  # [3 f][3 v][k j][ndg]   int s^j phi_i lam_v ds
    cell_mom_f: np.ndarray = None  # [3 v][1+ndiv][ndg]     int phi_i lam_v x^l y^m
    cell_mom_g: np.ndarray = None  # [3 v][1+ndiv][ndg][2]  int d_d phi_i lam_v x^l y^m
    bc_mat: np.ndarray = None  # [3 f][3 v][k j][k i]   int s^j (phi_{f,i}.n_f) lam_v ds
    # weak symmetry (stress) reference matrices
    rt_p1: np.ndarray = None  # [nrt][2][3]  int phi_i^d lam_v
    # constrained minimisation (EV) reference matrices
    dg_mono: np.ndarray = None  # [1+ndiv][ndg]       int psi_i x^l y^m
    hat_dg_rt: np.ndarray = None  # [3 v][ndg][nrt][2]  int lam_v psi_m phi_i^d
    mono_int: np.ndarray = None  # [1+ndiv]            int x^l y^m
    # primal space P_k (continuous Lagrange, Basix DOF order) for the fused input stage and the error estimator
    npk: int = 0
    pk_grad_dg: np.ndarray = None  # [ndg][npk][2]  reference gradient of the P_k basis at the DG_p nodes
    pk_to_dg: np.ndarray = None  # [ndg][npk]      L2 projection P_k -> DG_p on the reference cell
    pk_q: np.ndarray = None  # [nq][npk]       P_k basis at the cell quadrature points
    pk_gq: np.ndarray = None  # [nq][npk][2]    reference gradient
    pk_hq: np.ndarray = None  # [nq][npk][3]    reference second derivatives xx, xy, yy
    rt_div_q: np.ndarray = None  # [nq][nrt]       reference divergence of the RT basis
    # change of basis hierarchic RT -> Basix "RT" (Legendre variant), see basix_rt_legendre_maps
    rt_basix_fct: np.ndarray = None  # [k][k]        facet dofs:    c_basix = A c_hier (global facet orientation)
    rt_basix_int: np.ndarray = None  # [k*k-k][nrt]  interior dofs: c_basix = B c_hier (cell-local dofs)
    extra: dict = field(default_factory=dict)


def dubiner_basis(n):
    """Orthonormal polynomials of degree <= n on the reference triangle in Basix' polyset order
    idx(p, q) = (p+q)(p+q+1)/2 + q (Dubiner basis: P_{p,q} has x-degree p, leading coefficient > 0), as
    (exact polynomial, squared norm) pairs: Gram-Schmidt over x^p y^q, degree by degree, q descending."""
    out = {}
    done = []
    for deg in range(n + 1):
        for q in range(deg, -1, -1):
            pdeg = deg - q
            f = {(pdeg, q): Fr(1)}
            for g, gg in done:
                c = p_int_cell(p_mul(f, g)) / gg
                f = p_add(f, g, -c)
            gg = p_int_cell(p_mul(f, f))
            done.append((f, gg))
            out[(deg * (deg + 1)) // 2 + q] = (f, gg)
    return [out[i] for i in range(len(out))]


def legendre_01(n):
    """Legendre polynomials P_j(2s-1), j < n, on [0, 1] as exact univariate polynomials with their squared norms
    1/(2j+1) (Basix' orthonormal facet basis is sqrt(2j+1) P_j(2s-1))."""
    polys = []
    for j in range(n):
        f = {j: Fr(1)}
        for g, gg in polys:
            c = u_int01(u_mul(f, g)) / gg
            for kk, v in g.items():
                f[kk] = f.get(kk, Fr(0)) - c * v
        # normalise to P_j(1) = 1
        v1 = sum(f.values(), Fr(0))
        f = {kk: v / v1 for kk, v in f.items() if v != 0}
        polys.append((f, u_int01(u_mul(f, f))))
    return polys


def basix_rt_legendre_maps(k, rt):
    """Change of basis from the hierarchic RT_k element to Basix' "RT" element with the Legendre variant - the
    space `FluxEqlbEV` creates (`python/dolfinx_eqlb/eqlb/FluxEqlbEV.py:95`, `ufl.FiniteElement("RT", cell, k)`;
    Basix 0.6 `create_rt`: facet DOFs = normal moments against the orthonormal Legendre basis of P_{k-1} on the
    facet, interior DOFs = moments against the orthonormal basis of (P_{k-2})^2, direction-major).  Basix is third
    party and absent here: the functionals are restated from its definition (PARITY UNPINNED against a Basix
    install; for k >= 3 the ordering of the interior moments is the part that could differ).  The DOF vector of the
    same function is c_basix = L_basix(phi_hier) c_hier:
        A[i][j]  = coefficient of the facet moment int v.n s^j in the Legendre moment i     (k x k)
        B[i][j]  = L^basix_int,i(phi^hier_j)                                            ((k^2-k) x nrt)"""
    leg = legendre_01(k)
    A = np.zeros((k, k))
    for i, (f, gg) in enumerate(leg):
        nrm = float(gg) ** -0.5
        for j, c in f.items():
            A[i, j] = float(c) * nrm
    ncd = k * k - k
    B = np.zeros((ncd, len(rt)))
    if k > 1:
        dub = dubiner_basis(k - 2)
        nsc = len(dub)
        for d in range(2):
            for i, (q, gg) in enumerate(dub):
                nrm = float(gg) ** -0.5
                for j, (px, py) in enumerate(rt):
                    B[d * nsc + i, j] = float(p_int_cell(p_mul(q, px if d == 0 else py))) * nrm
    return A, B


_CACHE: dict = {}


def make_tables(k: int, p: int | None = None) -> Tables:
    """Build all tables for flux degree k and data degree p (default k-1)."""
    if p is None:
        p = k - 1
    if k < 1 or p < 0 or p > k - 1:
        raise ValueError("need k >= 1 and 0 <= p <= k-1")
    key = (k, p)
    if key in _CACHE:
        return _CACHE[key]

    rt = hierarchic_rt_basis(k)
    dg = lagrange_basis(p)
    nrt, ndg = len(rt), len(dg)
    div_idx, add_idx = rt_moment_indices(k)
    ndiv, nadd = len(div_idx), len(add_idx)
    closure = lagrange_facet_closure_dofs(p)
    ndg_fct = len(closure[0])

    qdeg = 2 if k == 1 else 2 * k + 1
    qpts, qwts = cell_quadrature(qdeg)
    nq = len(qwts)
    # facet interpolation points = Gauss-Jacobi rule of degree (k if k==1 else 2k)
    # (`e_raviart_thomas.py:63-71`): m = (deg+2)//2 points
    fdeg = k if k == 1 else 2 * k
    nqf = (fdeg + 2) // 2
    fs, fw = gauss_legendre_01(nqf)

    T = Tables(k=k, p=p, nrt=nrt, ndg=ndg, ndg_fct=ndg_fct, nq=nq, nqf=nqf, ndiv=ndiv, nadd=nadd)
    T.qpts, T.qwts, T.fpts_s, T.fwts = qpts, qwts, fs, fw
    T.fct_closure = np.array(closure, dtype=np.int32)
    T.div_lm = np.array(div_idx, dtype=np.int32).reshape(ndiv, 2)

    fpts = np.array([facet_point(f, s) for f in range(3) for s in fs])

    # facet interpolation matrix  M[f][j][d][n] = n_ref[f][d] s_n^j w_n
    M = np.zeros((3, k, 2, nqf))
    for f in range(3):
        for j in range(k):
            for d in range(2):
                M[f, j, d, :] = REF_NORMALS[f][d] * fs**j * fw
    T.M = M

    def tab_vec(pts):
        out = np.zeros((len(pts), nrt, 2))
        for n, (x, y) in enumerate(pts):
            for i, (px, py) in enumerate(rt):
                out[n, i, 0] = p_eval(px, x, y)
                out[n, i, 1] = p_eval(py, x, y)
        return out

    T.rt_q = tab_vec(qpts)
    T.rt_f = tab_vec(fpts)

    dg_q = np.zeros((3, nq, ndg))
    for i, ph in enumerate(dg):
        phx, phy = p_dx(ph), p_dy(ph)
        for n, (x, y) in enumerate(qpts):
            dg_q[0, n, i] = p_eval(ph, x, y)
            dg_q[1, n, i] = p_eval(phx, x, y)
            dg_q[2, n, i] = p_eval(phy, x, y)
    T.dg_q = dg_q
    T.dg_f = np.array([[p_eval(ph, x, y) for ph in dg] for (x, y) in fpts])
    T.hat_q = np.array([[p_eval(h, x, y) for h in HAT] for (x, y) in qpts])
    T.hat_f = np.array([[p_eval(h, x, y) for h in HAT] for (x, y) in fpts])

    # reversed-facet transform (`se/KernelData.cpp:55-64`):
    # column `line` holds (-1)^(i+1) C(line, i), i <= line
    tr = np.zeros((k, k))
    for line in range(k):
        for i in range(line + 1):
            tr[i, line] = (-1.0) ** (i + 1) * comb(line, i)
    T.trafo = tr

    # ---- reference matrices (exact) ----
    m00 = np.zeros((nrt, nrt))
    m01 = np.zeros((nrt, nrt))
    m11 = np.zeros((nrt, nrt))
    for i in range(nrt):
        for j in range(nrt):
            m00[i, j] = float(p_int_cell(p_mul(rt[i][0], rt[j][0])))
            m01[i, j] = float(p_int_cell(p_mul(rt[i][0], rt[j][1])))
            m11[i, j] = float(p_int_cell(p_mul(rt[i][1], rt[j][1])))
    T.rt_mass = np.stack([m00, m01 + m01.T, m11])

    fct_mom = np.zeros((3, 3, k, ndg))
    for f in range(3):
        for v in range(3):
            if v == f:
                continue
            lam_s = p_on_facet(HAT[v], f)
            for i, ph in enumerate(dg):
                ph_s = u_mul(p_on_facet(ph, f), lam_s)
                for j in range(k):
                    fct_mom[f, v, j, i] = float(u_int01(u_mul(ph_s, {j: Fr(1)})))
    T.fct_mom = fct_mom

    lm_all = [(0, 0)] + div_idx
    cmf = np.zeros((3, 1 + ndiv, ndg))
    cmg = np.zeros((3, 1 + ndiv, ndg, 2))
    for v in range(3):
        for t, (l, m) in enumerate(lm_all):
            wgt = p_mul(HAT[v], {(l, m): Fr(1)})
            for i, ph in enumerate(dg):
                cmf[v, t, i] = float(p_int_cell(p_mul(ph, wgt)))
                cmg[v, t, i, 0] = float(p_int_cell(p_mul(p_dx(ph), wgt)))
                cmg[v, t, i, 1] = float(p_int_cell(p_mul(p_dy(ph), wgt)))
    T.cell_mom_f, T.cell_mom_g = cmf, cmg

    bc = np.zeros((3, 3, k, k))
    for f in range(3):
        nx, ny = REF_NORMALS[f]
        for v in range(3):
            if v == f:
                continue
            lam_s = p_on_facet(HAT[v], f)
            for i in range(k):
                px, py = rt[f * k + i]
                vn = p_on_facet(p_add(p_scale(px, Fr(nx)), p_scale(py, Fr(ny))), f)
                g = u_mul(vn, lam_s)
                for j in range(k):
                    bc[f, v, j, i] = float(u_int01(u_mul(g, {j: Fr(1)})))
    T.bc_mat = bc

    rp = np.zeros((nrt, 2, 3))
    for i in range(nrt):
        for d in range(2):
            for v in range(3):
                rp[i, d, v] = float(p_int_cell(p_mul(rt[i][d], HAT[v])))
    T.rt_p1 = rp

    dgm = np.zeros((1 + ndiv, ndg))
    mono = np.zeros(1 + ndiv)
    for t, (l, m) in enumerate(lm_all):
        mono[t] = float(p_int_cell({(l, m): Fr(1)}))
        for i, ph in enumerate(dg):
            dgm[t, i] = float(p_int_cell(p_mul(ph, {(l, m): Fr(1)})))
    T.dg_mono, T.mono_int = dgm, mono
    hdr = np.zeros((3, ndg, nrt, 2))
    for v in range(3):
        for m_, ph in enumerate(dg):
            w = p_mul(HAT[v], ph)
            for i in range(nrt):
                hdr[v, m_, i, 0] = float(p_int_cell(p_mul(w, rt[i][0])))
                hdr[v, m_, i, 1] = float(p_int_cell(p_mul(w, rt[i][1])))
    T.hat_dg_rt = hdr

    T.rt_basix_fct, T.rt_basix_int = basix_rt_legendre_maps(k, rt)

    # ---- primal space P_k: fused projection (`lsolver/projection.py:17-77` for G = Pi(-grad u_h), Pi f_h) and
    # error estimator (`demo/poisson/demo_error_estimation.py:52-122`) ----
    pk = lagrange_basis(k)
    npk = len(pk)
    T.npk = npk
    dgn = lagrange_nodes(p)
    gdg = np.zeros((ndg, npk, 2))
    for j, ph in enumerate(pk):
        phx, phy = p_dx(ph), p_dy(ph)
        for i, (xn, yn) in enumerate(dgn):
            gdg[i, j, 0] = p_eval(phx, xn, yn)
            gdg[i, j, 1] = p_eval(phy, xn, yn)
    T.pk_grad_dg = gdg
    # L2 projection P_k -> DG_p: M_dg^-1 (psi_i, phi_j), exact
    Mdg = [[p_int_cell(p_mul(a, b)) for b in dg] for a in dg]
    Cx = [[p_int_cell(p_mul(a, b)) for b in pk] for a in dg]
    Pj = solve_exact(Mdg, Cx)
    T.pk_to_dg = np.array([[float(v) for v in row] for row in Pj])
    pkq = np.zeros((nq, npk))
    pkg = np.zeros((nq, npk, 2))
    pkh = np.zeros((nq, npk, 3))
    for j, ph in enumerate(pk):
        phx, phy = p_dx(ph), p_dy(ph)
        hxx, hxy, hyy = p_dx(phx), p_dy(phx), p_dy(phy)
        for n_, (x, y) in enumerate(qpts):
            pkq[n_, j] = p_eval(ph, x, y)
            pkg[n_, j, 0] = p_eval(phx, x, y)
            pkg[n_, j, 1] = p_eval(phy, x, y)
            pkh[n_, j, 0] = p_eval(hxx, x, y)
            pkh[n_, j, 1] = p_eval(hxy, x, y)
            pkh[n_, j, 2] = p_eval(hyy, x, y)
    T.pk_q, T.pk_gq, T.pk_hq = pkq, pkg, pkh
    rdq = np.zeros((nq, nrt))
    for i, (px, py) in enumerate(rt):
        dv = p_add(p_dx(px), p_dy(py))
        for n_, (x, y) in enumerate(qpts):
            rdq[n_, i] = p_eval(dv, x, y)
    T.rt_div_q = rdq
    T.extra["rt_exact"] = rt
    T.extra["dg_exact"] = dg
    T.extra["fpts"] = fpts
    _CACHE[key] = T
    return T
