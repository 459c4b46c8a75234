"""The C-ABI library loads on a machine WITHOUT a GPU and exports every symbol include/sg4d.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sg4d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sg4d_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_reference_launchers():
    names = _declared()
    for n in ("sg4d_furthest_point_sampling", "sg4d_gather_points", "sg4d_gather_points_grad", "sg4d_ball_query",
              "sg4d_group_points", "sg4d_group_points_grad"):
        assert n in names


def test_library_exports_every_declared_symbol():
    import sg4d
    lib = ctypes.CDLL(sg4d.library_path())
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/sg4d.h but not exported"
    assert lib.sg4d_abi_version() == 1


def test_python_binding_covers_the_header():
    from sg4d import _lib
    assert sorted(list(_lib.SIGNATURES) + _lib.OTHER_SYMBOLS) == _declared()
    lib = _lib.load()
    lib.sg4d_error_string.restype = ctypes.c_char_p
    assert b"invalid argument" in lib.sg4d_error_string(10001)


def test_missing_library_fails_loudly(monkeypatch):
    from sg4d import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", "/nonexistent/libsg4d.so")
    with pytest.raises(RuntimeError, match="no fallback"):
        _lib.load()
