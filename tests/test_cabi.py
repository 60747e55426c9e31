"""The C-ABI library loads on a CPU-only host and exports every symbol include/slak_b200.h
declares (no compute calls here)."""
import os
import re

from slak_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "slak_b200.h")).read()
    return sorted(set(re.findall(r"SLAK_API\s+[\w\s\*]+?\b(slak_\w+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in slak_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string_without_gpu():
    lib = _lib.load()
    assert lib.slak_version() >= 100
    assert isinstance(lib.slak_last_error(), bytes)
    assert lib.slak_device_ok() in (0, 1)


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    rc = lib.slak_dwconv2d_fwd(None, None, None, 1, 1, 1, 1, 3, 3, 0, 0, None)
    assert rc == -1 and b"null" in lib.slak_last_error()
    rc = lib.slak_dwconv2d_fwd(1, 1, 1, 1, 1, 4, 4, 4, 3, 0, 0, None)   # even kernel side
    assert rc == -1 and b"odd" in lib.slak_last_error()
    rc = lib.slak_dwconv2d_fwd(1, 1, 1, 1, 1, 4, 4, 3, 3, 7, 0, None)   # unknown dtype
    assert rc == -1 and b"Only support" in lib.slak_last_error()
