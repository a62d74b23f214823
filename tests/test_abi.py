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


def _header_prototypes():
    """name -> list of C parameter declarations, parsed from include/vbg.h (comments stripped)."""
    with open(os.path.join(ROOT, "include", "vbg.h")) as f:
        text = re.sub(r"/\*.*?\*/", " ", f.read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"VBG_API\s+(?:int|long long)\s+(vbg_\w+)\s*\((.*?)\)\s*;", text, flags=re.S):
        params = [p.strip() for p in m.group(2).split(",")]
        protos[m.group(1)] = [] if params == ["void"] else params
    return protos


def _class_of_c(decl):
    d = " ".join(decl.split())
    if "*" in d or "vbg_stream_t" in d:
        return "ptr"
    if re.match(r"(const )?(float|double)\b", d):
        return "float"
    return "int"


def _class_of_ctypes(t):
    if t in (ctypes.c_float, ctypes.c_double):
        return "float"
    if t in (ctypes.c_int, ctypes.c_longlong, ctypes.c_size_t, ctypes.c_uint, ctypes.c_ulonglong):
        return "int"
    return "ptr"


def test_ctypes_signatures_match_the_header_prototypes():
    """Same number of arguments, and the same class (pointer / integer / floating point) in every position, between the
    prototypes of include/vbg.h and the ctypes table the Python wrappers call through: an ABI drift (an argument added on one
    side only) would otherwise surface as a crash or as silent garbage on the GPU box."""
    protos = _header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    bad = []
    for name, params in protos.items():
        sig = _lib.SIGNATURES[name]
        if len(sig) != len(params):
            bad.append((name, f"{len(params)} parameters in the header, {len(sig)} in _lib.SIGNATURES"))
            continue
        for i, (p, t) in enumerate(zip(params, sig)):
            if _class_of_c(p) != _class_of_ctypes(t):
                bad.append((name, f"argument {i}: `{p}` vs {getattr(t, '__name__', t)}"))
    assert not bad, bad
