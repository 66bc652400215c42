"""The C-ABI library loads and exports every symbol include/ms_b200.h declares (no GPU needed)."""

import ctypes
import os
import re

import pytest

from ms_test_helpers import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ms_b200.h")).read()
    return sorted(set(re.findall(r"MS_API\s+[\w\s\*]+?\b(ms_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    from membrane_solver_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 40
    assert sorted(_lib.SIGNATURES) == declared


def test_library_exports_every_symbol():
    from membrane_solver_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(handle, name), name
    assert handle.ms_version() >= 100


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device every compute entry point fails loudly."""
    from membrane_solver_b200 import _lib

    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = ctypes.c_void_p()
    rc = _lib.lib().ms_ctx_create(0, ctypes.byref(h))
    assert rc != 0 and not h.value
    with pytest.raises(_lib.B200Error):
        _lib.check(rc)
    import numpy as np

    e = ctypes.c_double(0.0)
    pos = np.zeros((3, 3))
    tri = np.array([[0, 1, 2]], dtype=np.int32)
    rc = _lib.lib().ms_surface_energy_and_gradient(3, 1, _lib.dptr(pos), _lib.iptr(tri), _lib.dptr(np.ones(1)),
                                                   _lib.dptr(np.zeros((3, 3))), ctypes.byref(e), 1)
    assert rc != 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "membrane_solver_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text, f
