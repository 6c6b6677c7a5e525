"""The C-ABI library loads and exports every symbol include/eqlb_b200.h declares
(no compute calls without a GPU)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "eqlb_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eqlb_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_all_declared_symbols():
    from dolfinx_eqlb_b200 import cabi

    lib = cabi.load_library()
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.eqlb_version().decode().startswith("eqlb_b200")


def test_struct_layout_matches_header():
    from dolfinx_eqlb_b200 import cabi

    # 9 int32 + pointers ; the C struct has the same field order
    text = open(os.path.join(ROOT, "include", "eqlb_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    body = text[text.index("typedef struct eqlb_tables") : text.index("} eqlb_tables;")]
    names = re.findall(r"\*\s*([a-zA-Z_0-9]+)", body)
    py_names = [n for n, t in cabi.EqlbTables._fields_ if t not in (ctypes.c_int32,)]
    assert names == py_names


def test_no_cpu_fallback_without_device():
    """Without a usable GPU the compute entry points fail loudly (EQLB_ERR_CUDA)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dolfinx_eqlb_b200 import eqlb, mesh as ms
    import numpy as np

    m = ms.crossed_unit_square(2)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        eqlb.FluxEqlbSE(1, m, [np.zeros(m.ncell)], [np.zeros(2 * m.ncell)])


def test_input_errors_precede_device_checks():
    """Reference error texts (se/Patch.cpp:353-359) surface as RuntimeError."""
    import numpy as np

    from dolfinx_eqlb_b200 import eqlb, mesh as ms

    x = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    m = ms.build_topology(x, np.array([[0, 1, 3], [0, 2, 3]]))
    with pytest.raises(RuntimeError, match="has only 1 cells"):
        eqlb.FluxEqlbSE(1, m, [np.zeros(2)], [np.zeros(4)])


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dolfinx_eqlb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|liboracle|pyoracle|#include\s+\".*oracle)", src), f
