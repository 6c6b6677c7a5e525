"""Generate `tests/golden/ref_rt_element.npz` by EXECUTING the reference's own element
definition (`/root/reference/python/dolfinx_eqlb/elmtlib/e_raviart_thomas.py`, unchanged,
loaded from where it lies) on top of the basix stand-in `oracle/ref_shim/py/basix`.

    python tests/golden/make_ref_element.py          (needs /root/reference; build box only)

Per degree k = 1..4 the file holds, for `create_hierarchic_rt(triangle, k, True)`
(the discontinuous element FluxEqlbSE uses, `FluxEqlbSE.py:75-77`):
  coef_k   [nrt][2][nmono]  monomial coefficients of the basis (monomials degree-major,
                            x^a y^b with a descending inside a degree)
  X_k      [npts][2]        `element.points`
  M_k      [nrt][2*npts*nder] `element.interpolation_matrix` (flattened as Basix does)
  tab_k    [npts_s][nrt][2] tabulation at the fixed sample points `samples`
The committed file pins `dolfinx_eqlb_b200/tables.py` (tests/test_ref_element.py) and feeds
the shimmed reference build `oracle/_ref` with its flux element.
"""

import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/python/dolfinx_eqlb/elmtlib/e_raviart_thomas.py"


def load_reference_element_module():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim", "py"))
    spec = importlib.util.spec_from_file_location("ref_e_raviart_thomas", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import basix

    return mod, basix


def main():
    mod, basix = load_reference_element_module()
    rng = np.random.default_rng(7)
    s = rng.random((12, 2))
    samples = np.stack([s[:, 0] * (1 - s[:, 1]), s[:, 1] * (1 - s[:, 0]) * 0.9], axis=1)
    samples = samples[samples.sum(axis=1) < 1.0]
    out = {"samples": samples}
    for k in range(1, 5):
        for disc in (True, False):
            el = mod.create_hierarchic_rt(basix.CellType.triangle, k, disc)
            tag = f"{k}" if disc else f"{k}c"
            out[f"coef_{tag}"] = el.coef
            out[f"X_{tag}"] = el.points
            out[f"M_{tag}"] = el.interpolation_matrix
            out[f"tab_{tag}"] = el.tabulate(0, samples)[0]
            out[f"nder_{tag}"] = np.array(1 + 2 * el.nderivs)
    np.savez_compressed(os.path.join(HERE, "ref_rt_element.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_rt_element.npz"))


if __name__ == "__main__":
    main()
