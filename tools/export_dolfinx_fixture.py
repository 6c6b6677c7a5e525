#!/usr/bin/env python
"""Export a parity fixture from a LIVE dolfinx_eqlb installation (DOLFINx 0.6 + Basix 0.6).

This repository's oracle restates the reference algorithm but cannot be pinned against the
reference itself in the build image (no DOLFINx/Basix/Eigen there).  Run this
script where dolfinx_eqlb is installed; it writes `<name>.npz` files which
`tests/test_dolfinx_fixtures.py` compares with the oracle and with the CUDA path when they
are dropped into `tests/golden/dolfinx/`.

    python tools/export_dolfinx_fixture.py --n 5 --degree 2 --path se --out se_k2_n5.npz

NOT exercised in this repository's CI (needs DOLFINx); it only uses the public API of the
reference (`FluxEqlbSE.py:24-198`, `FluxEqlbEV.py:20-188`) and DOLFINx accessors.

Fixture layout (all arrays in DOLFINx's serial local numbering):
  x[nvert][2]          coordinates of the TOPOLOGY vertices
  cell_node[ncell][3]  cell -> vertex in the cell-local vertex order of DOLFINx
  G[nrhs][ncell][ndg][2], f[nrhs][ncell][ndg]   DG_(k-1) data by (cell, local dof)
  sigma_cells[nrhs][ncell][nrt]                 flux DOFs by (cell, local dof) through the flux dofmap
  bfct_prime / bfct_flux: boundary facets as sorted vertex pairs (pure Dirichlet: bfct_flux is empty)
  meta: degree, path, stress flag
"""

import argparse

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=5)
    ap.add_argument("--degree", type=int, default=2)
    ap.add_argument("--path", choices=["se", "ev"], default="se")
    ap.add_argument("--nrhs", type=int, default=1)
    ap.add_argument("--stress", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()

    from mpi4py import MPI
    from dolfinx import fem, mesh
    from dolfinx_eqlb.eqlb import FluxEqlbEV, FluxEqlbSE

    msh = mesh.create_unit_square(MPI.COMM_SELF, a.n, a.n, mesh.CellType.triangle, diagonal=mesh.DiagonalType.crossed)
    tdim = msh.topology.dim
    msh.topology.create_connectivity(tdim, 0)
    msh.topology.create_connectivity(tdim - 1, 0)
    msh.topology.create_connectivity(tdim - 1, tdim)
    ncell = msh.topology.index_map(tdim).size_local
    c2v = msh.topology.connectivity(tdim, 0).array.reshape(ncell, 3)
    gdm = msh.geometry.dofmap.array.reshape(ncell, -1)[:, :3]
    nvert = msh.topology.index_map(0).size_local
    xv = np.zeros((nvert, 2))
    xv[c2v.ravel()] = msh.geometry.x[gdm.ravel(), :2]

    k = a.degree
    V_dg = fem.FunctionSpace(msh, ("DG", k - 1))
    V_dgv = fem.VectorFunctionSpace(msh, ("DG", k - 1))
    rng = np.random.default_rng(a.seed)
    list_f, list_G = [], []
    for _ in range(a.nrhs):
        f, G = fem.Function(V_dg), fem.Function(V_dgv)
        f.x.array[:] = rng.standard_normal(f.x.array.shape[0])
        G.x.array[:] = rng.standard_normal(G.x.array.shape[0])
        list_f.append(f)
        list_G.append(G)

    if a.path == "se":
        eq = FluxEqlbSE(k, msh, list_f, list_G, equilibrate_stress=a.stress)
    else:
        eq = FluxEqlbEV(k, msh, list_f, list_G)

    # pure Dirichlet problem (the configuration of perftest.py:103,114-116): every boundary facet is "prime"
    f_prime = mesh.exterior_facet_indices(msh.topology)
    f_flux = np.zeros(0, dtype=np.int32)
    eq.set_boundary_conditions([f_prime] * a.nrhs, [[] for _ in range(a.nrhs)])
    eq.equilibrate_fluxes()

    ndg = V_dg.dofmap.list.array.reshape(ncell, -1).shape[1]
    dm = V_dg.dofmap.list.array.reshape(ncell, ndg)
    Gc = np.stack([G.x.array.reshape(-1, 2)[dm] for G in list_G])  # blocked bs = 2
    fc = np.stack([f.x.array[dm] for f in list_f])
    Vf = eq.list_flux[0].function_space
    fdm = Vf.dofmap.list.array.reshape(ncell, -1)
    sig = np.stack([s.x.array[fdm] for s in eq.list_flux])
    f2v = msh.topology.connectivity(tdim - 1, 0).array.reshape(-1, 2)
    np.savez_compressed(
        a.out, x=xv, cell_node=c2v.astype(np.int32), G=Gc, f=fc, sigma_cells=sig,
        bfct_prime=np.sort(f2v[f_prime], axis=1), bfct_flux=np.sort(f2v[f_flux], axis=1),
        meta=np.array([k, 0 if a.path == "se" else 1, int(a.stress), a.nrhs]),
    )
    print("wrote", a.out)


if __name__ == "__main__":
    main()
