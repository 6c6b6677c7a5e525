"""Wire format `.eqlb` (dolfinx_eqlb_b200/wire.py): round trip; a problem written from the
reference build's output replays through the oracle (CPU) and through the C ABI (GPU)."""

import numpy as np
import pytest

from common import PoissonCase, make_mesh
from dolfinx_eqlb_b200 import tables as tb, wire
from oracle import pyref as pr


def write_case(tmp_path, k=2):
    from oracle import pyoracle as po

    m = make_mesh("crossed", 4, 3, perturb=0.2)
    case = PoissonCase(m, k, [[1, 4], [2]], seed=3)
    bc = case.oracle_bc()
    run = pr.se_run if pr.available() else po.se_run
    maps = (pr if pr.available() else po).se_patch_maps(m, case.T, bc)
    sig = run(m, case.T, bc, case.G, case.F)
    path = str(tmp_path / "case.eqlb")
    wire.save_problem(path, m, k, case.T.p, case.bdata, case.G, case.F, sig, maps,
                      source="oracle/_ref (reference sources)" if pr.available() else "oracle port")
    return path, m, case, sig, maps


def test_roundtrip_and_oracle_replay(tmp_path):
    from oracle import pyoracle as po

    path, m, case, sig, maps = write_case(tmp_path)
    with open(path, "rb") as fh:
        assert fh.read(8) == b"EQLBWIRE"
    for mm in (False, True):
        d = wire.load_problem(path, mmap=mm)
        assert d["meta"]["k"] == 2 and d["meta"]["nrhs"] == 2 and d["meta"]["path"] == "se"
        for f in wire.MESH_FIELDS:
            assert np.array_equal(getattr(d["mesh"], f), getattr(m, f)), f
        assert np.array_equal(d["mesh"].bfct, m.bfct) and np.array_equal(d["mesh"].bfct_side, m.bfct_side)
        for key in ("cells", "fcts", "type"):
            assert np.array_equal(d["patch_maps"][key], maps[key])
    d = wire.load_problem(path)
    T = tb.make_tables(d["meta"]["k"], d["meta"]["p"])
    bd = d["bdata"]
    got = po.se_run(d["mesh"], T, po.BCData(bd.facet_type, bd.bflux, bd.local_fct_id, bd.node_on_stress_bnd), d["G"], d["F"])
    for r in range(2):
        assert np.abs(got[r] - d["sigma"][r]).max() < 1e-11 * np.abs(d["sigma"][r]).max()


def test_rejects_foreign_files(tmp_path):
    p = tmp_path / "x.eqlb"
    p.write_bytes(b"not a wire file at all")
    with pytest.raises(RuntimeError, match="not an EQLBWIRE file"):
        wire.load(str(p))


@pytest.mark.gpu
def test_gpu_replay(tmp_path):
    from dolfinx_eqlb_b200 import eqlb

    path, m, case, sig, maps = write_case(tmp_path)
    d = wire.load_problem(path)
    eq = eqlb.FluxEqlbSE(d["meta"]["k"], d["mesh"], d["F"], d["G"], degree_proj=d["meta"]["p"])
    eq.problem.set_bcs(d["bdata"])
    eq.equilibrate_fluxes()
    for r in range(2):
        assert np.abs(eq.list_flux[r] - d["sigma"][r]).max() < 1e-10 * np.abs(d["sigma"][r]).max()
    got = eq.problem.patch_maps()
    for key in ("ncells", "cells", "fcts", "inodes_local", "fcts_local", "type", "reversed", "reversion"):
        assert np.array_equal(got[key], d["patch_maps"][key]), key
