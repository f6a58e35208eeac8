"""The C-ABI library builds, loads and exports every symbol include/mrmp_b200.h declares.
No compute calls here (no GPU); the product path must fail loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

import sssp_b200
from sssp_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mrmp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrmp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = L.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mrmp_b200.h but not exported"
        assert n in L._SIGS, f"{n} has no ctypes signature in sssp_b200/lib.py"
    assert lib.mrmp_version().startswith(b"mrmp_b200")


def test_no_cpu_fallback_without_device():
    if L.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(L.MRMPError) as ei:
        sssp_b200.Context(L.MODEL_POINT2D, [0.1], np.array([[0.5, 0.5, 0.1]]))
    assert ei.value.code == L.ERR_CUDA
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(L.MRMPError):
        sssp_b200.gen_connect(sssp_b200.StatePoint2D(0, 0), [sssp_b200.CircleObstacle2D(.5, .5, .1)],
                              [0.1])


def test_argument_validation_is_reported_not_thrown():
    lib = L.load()
    h = ctypes.c_void_p()
    rads = np.array([0.1])
    rc = lib.mrmp_ctx_create(ctypes.byref(h), 7, 1, L.ptr(rads, L._dp), None, None, 0, None, 0.01,
                             0.01, float("nan"), 0)
    assert rc == L.ERR_INVALID and b"unknown model" in lib.mrmp_last_error()
    rc = lib.mrmp_ctx_create(ctypes.byref(h), L.MODEL_ARM22, 1, L.ptr(rads, L._dp), None, None, 0,
                             None, 0.01, 0.01, float("nan"), 0)
    assert rc == L.ERR_INVALID


def test_product_never_imports_oracle():
    """The shipped package must not reference the oracle (test infrastructure only)."""
    pkg = os.path.join(ROOT, "sssp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "README.md", f"{f} mentions the oracle"
