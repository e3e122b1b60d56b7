"""The C-ABI library loads and exports every symbol include/dgprf.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dgprf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dgprf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from dgprf import build, _ffi
    lib = build.build()
    L = ctypes.CDLL(lib)
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/dgprf.h but not exported"
        assert n in _ffi.SIGNATURES, f"{n} has no ctypes signature in dgprf/_ffi.py"
    assert set(_ffi.SIGNATURES) == set(names)


def test_struct_sizes_match_header():
    """ctypes mirrors of the POD structs must have the C layout (LP64)."""
    from dgprf import _ffi
    assert ctypes.sizeof(_ffi.Segment) == 24
    assert ctypes.sizeof(_ffi.Layer) == 6 * 4 + 4 * 8 + 8 + 8
    assert ctypes.sizeof(_ffi.Model) == 6 * 4 + 5 * 8 + _ffi.MAX_LAYERS * ctypes.sizeof(_ffi.Layer)


def test_argument_validation_needs_no_gpu():
    from dgprf import _ffi
    L = _ffi.lib()
    assert L.dgprf_version() >= 100
    n = ctypes.c_size_t(0)
    m = _ffi.Model()                      # all zeros: n_layers = 0 -> EINVAL, with a message
    assert L.dgprf_workspace_bytes(ctypes.byref(m), 10, 0, ctypes.byref(n)) == -1
    assert b"n_layers" in L.dgprf_last_error()
