"""Host-pointer calls staged through the GPU (EQLB_FLAG_HOST_PIPELINE): copy-in, patch
kernels and copy-out of spatial stages overlap on three streams.  The staged result must
agree with the one-piece result (same patches, other summation order per DOF) and with
the oracle."""

import numpy as np
import pytest

from common import PoissonCase, make_mesh
from dolfinx_eqlb_b200 import eqlb

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def run(case, path, pipeline, twice=False):
    cls = eqlb.FluxEqlbSE if path == "se" else eqlb.FluxEqlbEV
    eq = cls(case.k, case.mesh, case.F, case.G, host_pipeline=pipeline)
    eq.set_boundary_conditions(case.list_bfct_prime, case.list_bcs)
    eq.equilibrate_fluxes()  # zero flux on entry: EQLB_HOST_ZEROED
    if twice:
        eq.equilibrate_fluxes()  # accumulates: EQLB_HOST
    return eq


@pytest.mark.parametrize("path", ["se", "ev"])
@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("scramble", [None, 5])
def test_staged_equals_one_piece(monkeypatch, path, k, scramble):
    monkeypatch.setenv("EQLB_PIPE_STAGES", "5")
    m = make_mesh("crossed", 36, scramble, perturb=0.2)  # 5184 cells: above the staging threshold
    case = PoissonCase(m, k, [[1, 4], [2]], seed=3, hom=False)
    a = run(case, path, False)
    b = run(case, path, True)
    _, _, ncol = b.problem.patch_dims()
    for r in range(case.nrhs):
        assert rel_err(b.list_flux[r], a.list_flux[r]) < 1e-12


@pytest.mark.parametrize("path", ["se", "ev"])
def test_staged_accumulates(monkeypatch, path):
    monkeypatch.setenv("EQLB_PIPE_STAGES", "4")
    m = make_mesh("crossed", 34, None, perturb=0.2)
    case = PoissonCase(m, 2, [[1, 3]], seed=5, hom=False)
    once = run(case, path, True)
    twice = run(case, path, True, twice=True)
    assert rel_err(twice.list_flux[0], 2.0 * once.list_flux[0]) < 1e-12


def test_staged_oracle_parity(monkeypatch):
    from oracle import pyoracle as po

    monkeypatch.setenv("EQLB_PIPE_STAGES", "7")
    m = make_mesh("crossed", 33, None, perturb=0.2)
    case = PoissonCase(m, 2, [[1, 4]], seed=9, hom=False)
    ref = po.ev_run(m, case.T, case.oracle_bc(), case.G, case.F)
    eq = run(case, "ev", True)
    assert rel_err(eq.list_flux[0], ref[0]) < 1e-10


@pytest.mark.parametrize("path", ["se", "ev"])
@pytest.mark.parametrize("pipeline", [False, True])
def test_host_inputs_device_result(monkeypatch, path, pipeline):
    """EQLB_HOST_IN: G, f from host memory, the flux accumulated into a device vector."""
    import ctypes as C

    import torch

    from dolfinx_eqlb_b200 import cabi

    monkeypatch.setenv("EQLB_PIPE_STAGES", "4")
    m = make_mesh("crossed", 34, None, perturb=0.2)
    case = PoissonCase(m, 2, [[1, 3]], seed=5, hom=False)
    ref = run(case, path, False)
    eq = run(case, path, pipeline)
    prob = eq.problem
    G = [np.ascontiguousarray(g) for g in case.G]
    F = [np.ascontiguousarray(f) for f in case.F]
    dS = [torch.ones(ref.list_flux[0].size, dtype=torch.float64, device="cuda")]
    pS = (cabi.c_double_p * 1)(C.cast(dS[0].data_ptr(), cabi.c_double_p))
    torch.cuda.synchronize()
    if path == "se":
        rc = prob.lib.eqlb_se_run(prob.h, cabi.ptr_array(G), cabi.ptr_array(F), pS, cabi.c_double_p(), 3)
    else:
        rc = prob.lib.eqlb_ev_run(prob.h, cabi.ptr_array(G), cabi.ptr_array(F), pS, 3)
    assert rc == 0, prob.lib.eqlb_last_error().decode()
    torch.cuda.synchronize()
    assert rel_err(dS[0].cpu().numpy() - 1.0, ref.list_flux[0]) < 1e-12


def test_staged_stress(monkeypatch):
    """The weak-symmetry stage is patch local, so the stress path is staged as well."""
    from test_gpu_stress import elasticity_case

    monkeypatch.setenv("EQLB_PIPE_STAGES", "5")
    m = make_mesh("crossed", 36, None, perturb=0.2)
    T, G, f, bfp, bcs, neu = elasticity_case(m, 2, [], seed=3, galerkin=False)
    res = []
    for pipeline in (False, True):
        eq = eqlb.FluxEqlbSE(2, m, f, G, equilibrate_stress=True, host_pipeline=pipeline)
        eq.set_boundary_conditions(bfp, bcs)
        eq.equilibrate_fluxes()
        res.append(eq.list_flux)
    for r in range(2):
        assert rel_err(res[1][r], res[0][r]) < 1e-11
