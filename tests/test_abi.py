"""The C-ABI library loads without a GPU and exports exactly what include/vbg.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

from vibertgrid_pytorch_b200 import _lib


def _header_symbols():
    with open(os.path.join(ROOT, "include", "vbg.h")) as f:
        return re.findall(r"^VBG_API (?:int|long long) (vbg_\w+)\(", f.read(), flags=re.M)


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 25 and len(set(syms)) == len(syms)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vbg.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes binding and header drifted apart"
    assert lib.vbg_version() == 100


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    rc = lib.vbg_gemm(None, 0, None, 0, 0, None, 0, None, 0, None, 0, 1, 1, 1, None, 0, None)
    assert rc == _lib.VBG_EINVAL and "null pointer" in _lib.last_error()
    rc = lib.vbg_roi_align_fwd(None, 1, 1, 1, 4, None, None, 0, 0.25, 7, None, None, None)
    assert rc == _lib.VBG_EINVAL
    with pytest.raises(_lib.VbgError):
        _lib.check(rc, "vbg_roi_align_fwd")


def test_ops_refuse_cpu_tensors():
    import torch
    from vibertgrid_pytorch_b200 import ops
    with pytest.raises(TypeError, match="CUDA"):
        ops.softmax_rows(torch.zeros(2, 3))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libvbg_sm100a.so")
    with pytest.raises(ImportError, match="no fallback"):
        _lib.load()
