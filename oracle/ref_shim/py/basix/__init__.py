"""TEST INFRASTRUCTURE - NOT PRODUCT CODE.

Minimal stand-in for the `basix` Python package (Basix 0.6, absent from this image and
from /root/reference), just large enough to EXECUTE the reference's own element
definition `python/dolfinx_eqlb/elmtlib/e_raviart_thomas.py` unchanged.  That file only
needs: the RT span (`create_element(RT).wcoeffs`), `make_quadrature`, and
`create_custom_element`, which defines the basis as the dual of the functionals (x, M)
on span(wcoeffs).  The dual basis depends on the span and the functionals only - not on
Basix' orthonormal polyset or on which exact quadrature rule is used - so evaluating it
here reproduces the reference's hierarchic RT element up to rounding.

Polyset used by this stub (both producer and consumer of `wcoeffs` live here):
monomials x^a y^b, degree-major, `monomials(k)`.
"""

from __future__ import annotations

import enum
from fractions import Fraction as Fr

import numpy as np


class CellType(enum.Enum):
    point = 0
    interval = 1
    triangle = 2
    tetrahedron = 3


class ElementFamily(enum.Enum):
    custom = 0
    P = 1
    RT = 2


class LagrangeVariant(enum.Enum):
    unset = 0
    legendre = 1
    gll_warped = 2
    equispaced = 3


class MapType(enum.Enum):
    identity = 0
    contravariantPiola = 2


class SobolevSpace(enum.Enum):
    L2 = 0
    HDiv = 10


def monomials(k):
    """exponents (a, b) of the stub polyset of degree k"""
    return [(a, d - a) for d in range(k + 1) for a in range(d, -1, -1)]


class _RTSpan:
    """`basix.create_element(RT, triangle, k).wcoeffs`: a basis of
    RT_k = P_{k-1}^2 + x * Ptilde_{k-1} (Basix numbering: lowest order k = 1)."""

    def __init__(self, k):
        mon = monomials(k)
        idx = {m: i for i, m in enumerate(mon)}
        psize = len(mon)
        rows = []
        for d in range(2):
            for a, b in monomials(k - 1):
                r = np.zeros(2 * psize)
                r[d * psize + idx[(a, b)]] = 1.0
                rows.append(r)
        for a in range(k - 1, -1, -1):
            b = k - 1 - a
            r = np.zeros(2 * psize)
            r[idx[(a + 1, b)]] = 1.0
            r[psize + idx[(a, b + 1)]] = 1.0
            rows.append(r)
        self.wcoeffs = np.array(rows)
        self.degree = k


def create_element(family, cell, degree, lagrange_variant=LagrangeVariant.unset, *args):
    if family != ElementFamily.RT or cell != CellType.triangle:
        raise NotImplementedError("stub: only RT on triangles")
    return _RTSpan(degree)


def make_quadrature(cell, degree):
    """Gauss-Jacobi scheme with m = (degree+2)//2 points per direction (what Basix uses
    for intervals; for triangles Basix' default is a Xiao-Gimbutas table for degree <= 30 -
    any rule exact to `degree` defines the same functionals on polynomials)."""
    m = (degree + 2) // 2
    xg, wg = np.polynomial.legendre.leggauss(m)
    if cell == CellType.interval:
        return (0.5 + 0.5 * xg).reshape(m, 1), 0.5 * wg
    if cell == CellType.triangle:
        from scipy.special import roots_jacobi

        xa, wa = roots_jacobi(m, 1.0, 0.0)
        pts, wts = [], []
        for i in range(m):
            u = 0.5 * (1 + xa[i])
            for j in range(m):
                v = 0.5 * (1 + xg[j])
                pts.append([u, (1 - u) * v])
                wts.append(wa[i] * wg[j] * 0.125)
        return np.array(pts), np.array(wts)
    raise NotImplementedError


def _solve_exact(A, B):
    n = len(A)
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


class CustomElement:
    """Result of `create_custom_element`: dual basis of the functionals (x, M)."""

    def __init__(self, cell, value_shape, wcoeffs, x, M, nderivs, map_type, sobolev, discontinuous, hcd, hd):
        assert cell == CellType.triangle and list(value_shape) == [2]
        self.cell, self.map_type, self.sobolev = cell, map_type, sobolev
        self.discontinuous = bool(discontinuous)
        self.degree = int(hd)
        self.highest_complete_degree = int(hcd)
        self.nderivs = int(nderivs)
        self.x = [[np.array(a, dtype=np.float64) for a in lst] for lst in x]
        self.M = [[np.array(a, dtype=np.float64) for a in lst] for lst in M]
        wcoeffs = np.asarray(wcoeffs)
        n = wcoeffs.shape[0]
        mon = monomials(self.degree)
        psize = len(mon)
        assert wcoeffs.shape[1] == 2 * psize
        self.dim = n
        self.mon = mon
        nder = 1 + self.nderivs * 2

        # functionals in Basix order: entity dim, entity index, dof on entity
        rows = []  # rows[l][d*psize + m] = L_l applied to (monomial m in component d)
        self.entity_dofs = [[], [], []]
        for dim in range(3):
            for e in range(len(self.x[dim])):
                pts, mat = self.x[dim][e], self.M[dim][e]
                ids = []
                for i in range(mat.shape[0]):
                    ids.append(len(rows))
                    row = [Fr(0)] * (2 * psize)
                    for d in range(2):
                        for q in range(pts.shape[0]):
                            xq, yq = Fr(float(pts[q, 0])), Fr(float(pts[q, 1]))
                            for der in range(nder):
                                w = mat[i, d, q, der]
                                if w == 0.0:
                                    continue
                                w = Fr(float(w))
                                for m, (a, b) in enumerate(mon):
                                    if der == 0:
                                        v = xq**a * yq**b
                                    elif der == 1:
                                        v = a * xq ** (a - 1) * yq**b if a > 0 else Fr(0)
                                    else:
                                        v = b * xq**a * yq ** (b - 1) if b > 0 else Fr(0)
                                    row[d * psize + m] += w * v
                    rows.append(row)
                self.entity_dofs[dim].append(ids)
        assert len(rows) == n
        # D[l][j] = L_l(span_j);  basis_i = sum_j C[j][i] span_j with D C = I
        W = [[Fr(float(v)) for v in wcoeffs[j]] for j in range(n)]
        D = [[sum((rows[l][c] * W[j][c] for c in range(2 * psize) if W[j][c] != 0), Fr(0)) for j in range(n)] for l in range(n)]
        eye = [[Fr(int(i == j)) for j in range(n)] for i in range(n)]
        Cm = _solve_exact(D, eye)
        coef = np.zeros((n, 2, psize))
        for i in range(n):
            for c in range(2 * psize):
                s = sum((Cm[j][i] * W[j][c] for j in range(n) if W[j][c] != 0), Fr(0))
                coef[i, c // psize, c % psize] = float(s)
        self.coef = coef  # basis_i^d = sum_m coef[i, d, m] x^a_m y^b_m

    # what basix.finite_element.FiniteElement offers and the reference C++ reads
    @property
    def points(self):
        return np.vstack([a.reshape(-1, 2) for lst in self.x for a in lst])

    @property
    def interpolation_matrix(self):
        """(ndofs, value_size * npoints * nderivs) with index (d*npts + pt)*nder + der -
        the layout `base/KernelData.cpp:232-256` indexes."""
        npts = self.points.shape[0]
        nder = 1 + self.nderivs * 2
        out = np.zeros((self.dim, 2, npts, nder))
        row, col = 0, 0
        for dim in range(3):
            for e in range(len(self.x[dim])):
                mat = self.M[dim][e]
                nd, nq = mat.shape[0], mat.shape[2]
                out[row : row + nd, :, col : col + nq, :] = mat
                row += nd
                col += nq
        return out.reshape(self.dim, 2 * npts * nder)

    def tabulate(self, nd, pts):
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
        assert nd == 0
        out = np.zeros((1, pts.shape[0], self.dim, 2))
        for m, (a, b) in enumerate(self.mon):
            v = pts[:, 0] ** a * pts[:, 1] ** b
            out[0] += v[:, None, None] * self.coef[None, :, :, m]
        return out


def create_custom_element(cell, value_shape, wcoeffs, x, M, nderivs, map_type, sobolev, discontinuous, hcd, hd, *args):
    return CustomElement(cell, value_shape, wcoeffs, x, M, nderivs, map_type, sobolev, discontinuous, hcd, hd)
