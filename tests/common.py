"""Shared problem factory for the tests: meshes, Galerkin-orthogonal data, BCs."""

import numpy as np

import fem_mini as fm
from dolfinx_eqlb_b200 import eqlb, mesh as ms, tables as tb


def make_mesh(kind, n, scramble=None, perturb=0.0):
    if kind == "crossed":
        return ms.crossed_unit_square(n, scramble_seed=scramble, perturb=perturb)
    if kind == "randdiag":
        return ms.random_diagonal_square(n, seed=11, scramble_seed=scramble, perturb=perturb)
    if kind == "fan":  # interior hub with 4 n cells
        return ms.fan_unit_square(n, False, scramble_seed=scramble)
    if kind == "halffan":  # boundary hub with 3 n cells
        return ms.fan_unit_square(n, True, scramble_seed=scramble)
    if kind == "delaunay":  # unstructured, n segments per side
        return ms.delaunay_unit_square(n, seed=5, scramble_seed=scramble)
    raise ValueError(kind)


def neumann_coeffs(m, T, sides, rng, hom=False):
    out = {}
    for fct in m.boundary_facets(sides):
        out[int(fct)] = np.zeros(T.k) if hom else fm.random_dg(rng, T.k)
    return out


class PoissonCase:
    """One or several Poisson problems (different Neumann side sets) on one mesh with
    Galerkin-orthogonal projected fluxes."""

    def __init__(self, m, k, neumann_sets, seed=0, p=None, hom=False, galerkin=True):
        rng = np.random.default_rng(seed)
        self.mesh, self.k = m, k
        self.T = tb.make_tables(k, p)
        T = self.T
        self.G, self.F, self.neu, self.dsides = [], [], [], []
        for nsides in neumann_sets:
            dsides = [s for s in (1, 2, 3, 4) if s not in nsides]
            neu = neumann_coeffs(m, T, nsides, rng, hom)
            f = fm.random_dg(rng, m.ncell * T.ndg)
            if galerkin:
                G, _, _ = fm.solve_poisson(m, max(T.p + 1, 1), T, f, dsides, neu)
            else:
                G = fm.random_dg(rng, m.ncell * T.ndg * 2) * np.where(rng.random(m.ncell * T.ndg * 2) < 0.5, -1.0, 1.0)
            self.G.append(G)
            self.F.append(f)
            self.neu.append(neu)
            self.dsides.append(dsides)
        self.nrhs = len(neumann_sets)
        self.list_bfct_prime = [m.boundary_facets(d) for d in self.dsides]
        self.list_bcs = []
        for neu in self.neu:
            if len(neu):
                fcts = np.array(sorted(neu.keys()), dtype=np.int32)
                self.list_bcs.append([eqlb.fluxbc(fcts, np.array([neu[int(f)] for f in fcts]))])
            else:
                self.list_bcs.append([])
        self.bdata = eqlb.boundarydata(self.list_bcs, m, T, self.list_bfct_prime)

    def oracle_bc(self):
        from oracle import pyoracle as po

        return po.BCData(self.bdata.facet_type, self.bdata.bflux, self.bdata.local_fct_id, self.bdata.node_on_stress_bnd)
