"""Generates tests/golden/*.npz.  The reference ships no golden vectors.

SE and stress fixtures (`se_*`, `stress_*`) are OUTPUTS OF THE REFERENCE ITSELF: its own C++
sources compiled unchanged against stand-in headers (`oracle/_ref/libeqlb_ref.so`, recipe
`make -C oracle ref`, needs /root/reference) run on seeded inputs on the reference's own
fixture sizes (2x2 / 5x5 crossed unit squares, test_fluxeqlb_conditions.py:47-49); the
files carry `source = "reference"`.  EV fixtures (`ev_*`) come from the reference's
ev::reconstruction in the same library (its patch loop, assembly, lifting, LU and scatter;
the three fixed FFCx forms and the DOF transformations are restated in oracle/ref_driver.cpp).  Every
vector is accepted only if the reference's acceptance invariants hold.
Run from the repo root (build container):  python tests/golden/make_golden.py"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fem_mini as fm  # noqa: E402
from common import PoissonCase, make_mesh  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from oracle import pyref as pr  # noqa: E402

CASES = [
    # name, path, kind, n, scramble, perturb, k, neumann sets, seed
    ("se_k1_crossed2", "se", "crossed", 2, None, 0.0, 1, [[1, 4]], 1),
    ("se_k2_crossed5_scr", "se", "crossed", 5, 3, 0.2, 2, [[1, 4], [1, 3], [2], [1, 3, 4]], 2),
    ("se_k3_crossed2", "se", "crossed", 2, 7, 0.1, 3, [[1, 4]], 3),
    ("ev_k2_crossed5_scr", "ev", "crossed", 5, 3, 0.2, 2, [[], [1, 4]], 4),
    ("ev_k3_crossed2", "ev", "crossed", 2, None, 0.0, 3, [[]], 5),
]


# elasticity rows with weak symmetry + Korn constants: name, kind, n, scramble, perturb, k, traction sides, seed
# (traction layouts of test_stressqlb_bcond.py:167-190; [1, 2] needs the grouped corner patches)
STRESS_CASES = [
    ("stress_k2_crossed5_scr", "crossed", 5, 3, 0.2, 2, [1], 6),
    ("stress_k2_crossed4_corner", "crossed", 4, None, 0.2, 2, [1, 2], 7),
    ("stress_k3_crossed3", "crossed", 3, 5, 0.1, 3, [1], 8),
]


def build_stress(name, kind, n, scramble, perturb, k, nsides, seed):
    from test_gpu_stress import elasticity_case
    from dolfinx_eqlb_b200 import eqlb

    m = make_mesh(kind, n, scramble, perturb)
    T, G, f, bfp, bcs, neu = elasticity_case(m, k, nsides, seed=seed)
    bd = eqlb.boundarydata(bcs, m, T, bfp, True)
    return m, T, G, f, bfp, bcs, bd


def build(name, path, kind, n, scramble, perturb, k, nsets, seed):
    m = make_mesh(kind, n, scramble, perturb)
    case = PoissonCase(m, k, nsets, seed=seed, hom=(path == "ev"))
    return m, case


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    for spec in CASES:
        name, path = spec[0], spec[1]
        m, case = build(*spec)
        if path == "se":
            sig = pr.se_run(m, case.T, case.oracle_bc(), case.G, case.F)
            for r in range(case.nrhs):
                assert fm.check_divergence(m, case.T, sig[r], case.G[r], case.F[r]) < 1e-12
                assert fm.check_jump(m, case.T, sig[r], case.G[r]) < 1e-12
        else:
            sig = pr.ev_run(m, case.T, case.oracle_bc(), case.G, case.F)
            for r in range(case.nrhs):
                s = fm.conforming_to_drt(m, case.T, sig[r])
                z = np.zeros_like(case.G[r])
                assert fm.check_divergence(m, case.T, s, z, case.F[r]) < 1e-10
                assert fm.check_jump(m, case.T, s, z) < 1e-12
        maps = (pr if path == "se" else po).se_patch_maps(m, case.T, case.oracle_bc())
        np.savez_compressed(
            os.path.join(out, name + ".npz"), G=np.array(case.G), F=np.array(case.F), sigma=np.array(sig),
            source=np.array("reference"),
            cells=maps["cells"], fcts=maps["fcts"], type=maps["type"], fcts_local=maps["fcts_local"],
            reversed=maps["reversed"], cell_node=m.cell_node, x=m.x,
        )
        print("wrote", name)
    for spec in STRESS_CASES:
        m, T, G, f, bfp, bcs, bd = build_stress(*spec)
        bc = po.BCData(bd.facet_type, bd.bflux, bd.local_fct_id, bd.node_on_stress_bnd)
        sig, korn = pr.se_run(m, T, bc, G, f, stress=True, korn=True)
        for r in range(2):
            assert fm.check_divergence(m, T, sig[r], G[r], f[r]) < 1e-12
            assert fm.check_jump(m, T, sig[r], G[r]) < 1e-10
        ws = fm.check_weak_symmetry(m, T, sig[0], sig[1])
        assert ws < 1e-10, (spec[0], ws)
        np.savez_compressed(os.path.join(out, spec[0] + ".npz"), G=np.array(G), F=np.array(f), sigma=np.array(sig), korn=korn,
                            cell_node=m.cell_node, x=m.x, source=np.array("reference"))
        print("wrote", spec[0])


if __name__ == "__main__":
    main()
